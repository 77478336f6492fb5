// Host-side walk of the kernel's tile maps (decode_tile / TileIter of igemm_tc.cuh are __host__ __device__): every
// launch geometry the planner can produce must enumerate each (problem, frame group, tile row, tile column, N tile)
// exactly once, the incremental walk must agree with the full decode for every grid size, and under `pair_order` the
// tiles 2k, 2k+1 of a CTA pair must be two different M tiles of ONE N tile of ONE problem.
// Built and run by tests/test_host_cpu.py (no GPU, no CUDA runtime call).
#include "../att-aspp-unet_b200/csrc/igemm_tc.cuh"
#include <cstdio>
#include <cstring>
#include <set>
#include <tuple>
using namespace aau;

static int failures = 0;
#define CHECK(c, ...) do { if (!(c)) { if (failures < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++failures; } } while (0)

static void fill(IgemmParams& P, int nprob, int B, int H, int W, int TH, int TW, int VW, int MT, int TB, int n_tiles, int n_out, int amode,
                 int pair_order, int tile_iter) {
    memset(&P, 0, sizeof(P));
    P.nprob = nprob; P.TH = TH; P.TW = TW; P.VW = VW; P.MT = MT; P.TB = TB; P.n_out = n_out; P.BN = n_out; P.amode = amode;
    P.pair_order = pair_order; P.tile_iter = tile_iter;
    int begin = 0;
    for (int i = 0; i < nprob; ++i) {
        IgemmProblem& q = P.prob[i];
        q.H = H; q.W = W;
        q.tiles_x = (W + VW - 1) / VW;
        q.tiles_per_img = q.tiles_x * ((H + TH * MT - 1) / (TH * MT));
        q.m_tiles = B / TB * q.tiles_per_img;
        q.n_tiles = n_tiles;
        q.tile_begin = begin;
        q.fd_n_tiles = make_fastdiv((uint32_t)q.n_tiles);
        q.fd_tiles_per_img = make_fastdiv((uint32_t)q.tiles_per_img);
        q.fd_tiles_x = make_fastdiv((uint32_t)q.tiles_x);
        begin += q.m_tiles * q.n_tiles;
    }
    P.total_tiles = begin;
}

static void check_geometry(const char* name, int nprob, int B, int H, int W, int TH, int TW, int VW, int MT, int TB, int n_tiles, int n_out,
                           int amode, int pair_order) {
    IgemmParams P;
    const int halo = amode >= AMODE_DXN ? 1 : 0;
    for (int titer = 0; titer < 2; ++titer) {
        fill(P, nprob, B, H, W, TH, TW, VW, MT, TB, n_tiles, n_out, amode, pair_order, titer);
        std::set<std::tuple<int, int, int, int, int>> seen;
        for (int t = 0; t < P.total_tiles; ++t) {
            const TileCoord tc = decode_tile(P, t);
            const IgemmProblem& q = P.prob[tc.pi];
            CHECK(tc.pi >= 0 && tc.pi < nprob && t >= q.tile_begin && t < q.tile_begin + q.m_tiles * q.n_tiles, "%s t=%d problem %d", name, t, tc.pi);
            CHECK(tc.b >= 0 && tc.b + TB <= B && tc.b % TB == 0, "%s t=%d frame %d", name, t, tc.b);
            CHECK(tc.y0 >= 0 && tc.y0 < H && tc.y0 % (TH * MT) == 0, "%s t=%d y0 %d", name, t, tc.y0);
            CHECK(tc.x0 + halo >= 0 && tc.x0 + halo < W && (tc.x0 + halo) % VW == 0, "%s t=%d x0 %d", name, t, tc.x0);
            CHECK(tc.n0 >= 0 && tc.n0 < n_tiles * n_out && tc.n0 % n_out == 0, "%s t=%d n0 %d", name, t, tc.n0);
            CHECK(seen.insert({tc.pi, tc.b, tc.y0, tc.x0, tc.n0}).second, "%s t=%d visited twice", name, t);
        }
        CHECK((int)seen.size() == P.total_tiles, "%s: %d distinct tiles of %d", name, (int)seen.size(), P.total_tiles);
        if (pair_order)
            for (int t = 0; t + 1 < P.total_tiles; t += 2) {
                const TileCoord a = decode_tile(P, t), b = decode_tile(P, t + 1);
                CHECK(a.pi == b.pi && a.n0 == b.n0, "%s pair %d: problems %d/%d n0 %d/%d", name, t, a.pi, b.pi, a.n0, b.n0);
                CHECK(a.b != b.b || a.y0 != b.y0 || a.x0 != b.x0, "%s pair %d: same M tile twice", name, t);
            }
        // every role walks t = first, first + stride, ...; the incremental form must reproduce the full decode
        const int grids[] = {1, 2, 3, 7, 74, 148, 296, 592};
        for (int grid : grids)
            for (int first = 0; first < grid && first < 5; ++first) {
                TileIter it;
                int steps = 0;
                for (it.init(P, first, grid); it.valid(); it.next(), ++steps) {
                    const TileCoord a = it.coord(P), b = decode_tile(P, it.t);
                    CHECK(a.pi == b.pi && a.b == b.b && a.y0 == b.y0 && a.x0 == b.x0 && a.n0 == b.n0,
                          "%s titer=%d grid=%d first=%d t=%d: (%d,%d,%d,%d) vs (%d,%d,%d,%d)", name, titer, grid, first, it.t, a.b, a.y0, a.x0, a.n0,
                          b.b, b.y0, b.x0, b.n0);
                }
                CHECK(steps == (P.total_tiles - first + grid - 1) / grid || first >= P.total_tiles, "%s grid=%d first=%d: %d steps", name, grid, first, steps);
            }
    }
}

int main() {
    // (name, problems, B, H, W, TH, TW, VW, MT, TB, N tiles, n_out, staging mode, pair order)
    check_geometry("d1.1 rs 562x744", 1, 28, 562, 744, 4, 32, 30, 1, 1, 1, 32, AMODE_RS, 0);
    check_geometry("u2.conv.1 rs MT2", 1, 28, 281, 372, 4, 32, 30, 2, 1, 1, 64, AMODE_RS, 0);
    check_geometry("u2.conv.0 dxn split", 1, 28, 281, 372, 4, 32, 30, 1, 1, 2, 32, AMODE_DXN, 0);
    check_geometry("u3.conv slab MT2", 1, 28, 140, 186, 8, 16, 16, 2, 1, 1, 128, AMODE_SLAB, 0);
    check_geometry("u4.conv slab", 1, 56, 70, 93, 8, 16, 16, 1, 1, 1, 256, AMODE_SLAB, 0);
    check_geometry("aspp 3 problems, 2 N tiles, pairs, 2 frames", 3, 28, 35, 46, 4, 16, 16, 1, 2, 2, 256, AMODE_TAP, 1);
    check_geometry("project pairs", 1, 28, 35, 46, 8, 16, 16, 1, 1, 2, 256, AMODE_TAP, 1);
    check_geometry("u4.up 2 frames", 1, 8, 35, 46, 4, 16, 16, 1, 2, 2, 256, AMODE_TAP, 0);
    check_geometry("u1.up 1x128", 1, 3, 281, 372, 1, 128, 128, 1, 1, 1, 128, AMODE_TAP, 0);
    check_geometry("odd tiny", 1, 3, 17, 33, 4, 32, 30, 1, 1, 1, 16, AMODE_DXN, 0);
    check_geometry("one tile", 1, 1, 4, 30, 4, 32, 30, 1, 1, 1, 32, AMODE_RS, 0);
    check_geometry("4 problems", 4, 2, 16, 16, 8, 16, 16, 1, 1, 3, 64, AMODE_TAP, 0);
    // the multiply-high division itself, against the hardware-free definition
    for (uint32_t d : {1u, 2u, 3u, 5u, 7u, 12u, 13u, 15u, 25u, 71u, 141u, 843u, 3525u, 65535u, 1000003u}) {
        const FastDiv f = make_fastdiv(d);
        for (uint32_t n : {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 1000u * d + 999u, 123456789u, (1u << 31) - 1})
            CHECK(fdiv(n, f) == n / d, "fdiv %u / %u = %u", n, d, fdiv(n, f));
    }
    printf("%s (%d failures)\n", failures ? "FAILED" : "ok", failures);
    return failures ? 1 : 0;
}
