// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands staged in
// shared memory by TMA, results leaving through swizzled shared memory + TMA stores).  One persistent kernel
// serves every GEMM-shaped layer of AttentionASPPUNet:
//
//   rows    (M) : 128 output pixels = a TH x TW patch of one frame (NHWC activations, TH*TW == 128)
//   columns (N) : BN output channels (<= 256)
//   depth   (K) : taps x Cin, walked in sub-blocks of KC channels (KC*2 bytes == the TMA/UMMA swizzle width)
//
// A operand (activations), two staging modes, own ring of `nA` slots:
//   AMODE_TAP  : one TMA box (KC, TW, TH) per (tap, channel chunk), shifted by the tap offset * dilation; image
//                borders and dilation overhang come back as zeros from the TMA out-of-bounds fill.  Works for
//                1x1, 3x3 and dilated 3x3.
//   AMODE_SLAB : 3x3 / dilation 1 only.  One TMA box (KC, TW, TH+2) per (dx, channel chunk) -- a column-shifted
//                slab with a one-row halo above and below.  The three vertical taps are then three MMA groups whose
//                A descriptors start TW rows apart inside the same slab (TW % 8 == 0 keeps them on the swizzle
//                period), so every activation byte is fetched from L2 3.75x (TH=8) instead of 9x.
//   AMODE_DXN  : 3x3 / dilation 1 / Cout <= 64.  The three horizontal taps are stacked along N instead of being
//                three shifted A reads: ONE slab (KC, 32, TH+2) per channel chunk, B = [W(dx=-1); W(0); W(+1)]
//                (3*Cout rows), so an MMA of N' = 3*Cout produces E_dx = A . W_dx for all three dx at once and the
//                epilogue forms out[x] = E_-1[x-1] + E_0[x] + E_+1[x+1] with two warp shuffles per channel (a warp
//                is one 32-pixel row of the tile; its 30 inner pixels are valid outputs, tiles overlap by two
//                columns).  3x fewer MMAs and A loads -- what the small-N layers, which are limited by the
//                shared-memory read of the A operand (64 cycles per MMA whatever N), need.
//   AMODE_RS   : 3x3 / dilation 1 / small N with smem-resident weights.  ONE slab (KC, 32, TH+2) per channel chunk as in
//                DXN, but the nine taps are nine MMAs that accumulate into the SAME columns: tap (dy, dx) reads the
//                slab through an A descriptor whose start address is advanced by (dy*32 + dx) pixel rows.  The UMMA
//                swizzle is a function of the absolute shared-memory address, so a start address that is not a
//                multiple of the 8-row swizzle period reads exactly what TMA wrote (tools/mma_probe.cu, probe 2).
//                Row r = ty*32 + j of the accumulator is output pixel (y0+ty, x0+j) for j < 30; j = 30, 31 wrap into
//                the next slab row and are discarded.  Costs 3x the MMAs of DXN (each 32 + N/4 cycles: the shared-
//                memory operand read, profiles/r01_mma_probe.txt) but needs no dx combination in the epilogue, which
//                is what bounds the DXN layers whose K is small.
//   With MT = 2 (SLAB, BN <= 128) a tile is two vertically adjacent 128-pixel blocks fed from one taller slab: every
//   weight sub-block is used for two MMA groups, which halves the L2->SM weight traffic of the mid-size layers.
//   Per-tap staged tiles of small images may span TB = 2 frames (the TMA boxes get a batch extent): the 35 x 46 bridge
//   pads to 36 instead of 40 rows.
// B operand (weights [N][K], K contiguous), own ring of `nB` slots of one (KC x BN) sub-block each -- or, when the
//   whole weight matrix of the layer fits (`b_resident`), loaded ONCE per CTA and kept for every tile.
//
// CTA pairs (PAIR, cluster (2,1,1), tcgen05.*.cta_group::2): one M = 256 MMA spans two SMs -- each CTA stages its own 128
// pixel rows of A and HALF of the rows of every weight sub-block (streamed or resident), the leader's elected lane issues
// the instruction for both, accumulators land in each CTA's own TMEM, one multicast commit frees both CTAs' ring slots.
// It halves the weight bytes every SM pulls from L2 (what bounded the N = 128 slab layers), halves the B operand's
// shared-memory read per MMA (32 + N/8 instead of 32 + N/4 cycles: the small-N layers) and lets one issuer warp pace
// two SMs.  Launches with several N tiles walk the tiles pairwise (`pair_order`) so that a pair shares its N tile.
//
// Instantiations: the template parameters AM / EP / KKT / PL fix the staging mode, the epilogue, the MMAs per sub-block
// and the fused MaxPool at compile time for the hot path, both storage types (a third of the generic kernel's 110 KB of SASS: the
// single-warp roles miss the instruction cache less and skip the uniform mode branches); -1 = decided at run time
// from IgemmParams, which every plan and the forced-plan parity tests can use (igemm_inst.cuh lists the instantiations).
//
// Warp roles (64 + 128*NG threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (both run warp-uniform,
// the asynchronous instructions under elect.sync), then NG in {2, 4} epilogue groups of four warps (one warp per TMEM
// lane quarter); group g drains accumulator stage g, so the epilogue of a tile overlaps the MMAs of the next ones.
// Small-N layers run two CTAs per SM (NG = 2 each) or, when shared memory allows one CTA only, NG = 4: what paces them
// is the per-tile epilogue latency chain and the shared-memory operand feed, not the tensor pipe (profiles/r01_*).
//
// Epilogues (fp32 math on the accumulator, single rounding to the 16-bit activation type):
//   EPI_STORE   : + bias (per channel or per image), optional ReLU -> swizzled smem -> TMA store into the NHWC
//                 destination at a channel offset / pixel stride (this is how ASPP branches and skip tensors land
//                 directly inside concatenated buffers; ragged tiles are clipped by the TMA unit)
//   EPI_CONVT   : ConvTranspose2d(2,2): column block (a,b) is TMA-stored through a strided view of the output whose
//                 pixel (x, y) is output pixel (2x+b, 2y+a)
//   EPI_CONVTFIX: ConvTranspose2d(2,2) fused with the decoder's one-row / one-column bilinear fix-up (F.interpolate to
//                 the skip's size when the pooled size was odd: 2n -> 2n+1 along ONE axis).  The GEMM has two taps
//                 along that axis (offset -1 and 0) and three column blocks per tile, up = X[j-1].W1, mid0 = X[j].W0,
//                 mid1 = X[j].W1 (W0 / W1 = the transposed-conv phases along the fixed axis), i.e. the transposed-conv
//                 rows 2j-1, 2j, 2j+1.  Output row o = 2j+p is a two-term blend of two ADJACENT transposed-conv rows,
//                 so the epilogue forms both of its outputs as  c[o][0].up + c[o][1].mid0 + c[o][2].mid1 + bias  in
//                 fp32 (c = ATen's bilinear weights, tabulated per output row on the host) and stores them through
//                 the same strided (a,b) views as EPI_CONVT -- straight into the concatenated decoder buffer, with no
//                 intermediate tensor, no separate resize kernel and a single rounding.
//   EPI_GATE    : attention gate: psi = sigmoid(w_psi . relu(acc + bias) + b_psi); the skip tensor row is scaled
//                 by psi (or 1+psi for the ablation flavour) in place; psi optionally written out
//   EPI_OUTCONV : last decoder conv fused with out_conv: logit = w_out . relu(acc + bias) + b_out (fp32 out)
//
// Reference semantics being implemented: attention_aspp_unet_pipeline_stage.py:59-65 (ConvBNReLU), :67-83 (ASPP),
// :85-92 (AttentionGate), :98-109 (UpBlock), :122 (out_conv); test_ablation.py:128-143 (ablation gate).
#pragma once
#include "ptx_sm100.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

// Optional phase timing (build with -DAAU_EPI_TIMING; tools only): cycle counters per role phase, accumulated
// into the 24 uint64 slots behind the error flag and dumped per launch by the engine in profile mode.
#ifdef AAU_EPI_TIMING
#define TM_DECL() long long tm_mark_ = clock64(); unsigned long long tm_acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TM_MARK(i) do { const long long n_ = clock64(); tm_acc_[i] += (unsigned long long)(n_ - tm_mark_); tm_mark_ = n_; } while (0)
#define TM_FLUSH(base, cond) do { if (cond) { for (int i_ = 0; i_ < 8; ++i_) atomicAdd(reinterpret_cast<unsigned long long*>(P.err + 16) + (base) + i_, tm_acc_[i_]); } } while (0)
#else
#define TM_DECL() do {} while (0)
#define TM_MARK(i) do {} while (0)
#define TM_FLUSH(base, cond) do {} while (0)
#endif

namespace aau {

enum { EPI_STORE = 0, EPI_CONVT = 1, EPI_GATE = 2, EPI_OUTCONV = 3, EPI_CONVTFIX = 4 };
enum { AMODE_TAP = 0, AMODE_SLAB = 1, AMODE_DXN = 2, AMODE_RS = 3 };
enum { IGEMM_MAX_PROBLEMS = 4, IGEMM_MAX_SLOTS = 12, IGEMM_MAX_GROUPS = 4 };
__host__ __device__ constexpr int igemm_threads(int ng) { return 64 + 128 * ng; }   // TMA warp + MMA warp + ng epilogue groups of 4 warps
enum { ERR_PRODUCER_WAIT = 101, ERR_MMA_WAIT_FULL = 102, ERR_MMA_WAIT_TMEM = 103, ERR_EPI_WAIT = 104 };

// Division by a launch-time constant as multiply-high + shift (the tile decode runs once per tile per role and four
// hardware integer divisions were ~25 % of the epilogue's instructions).  Exact for n < 2^31.
struct FastDiv {
    uint32_t mul, shr;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.shr = s;
    f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
    return f;
}
__host__ __device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv f) {
#ifdef __CUDA_ARCH__
    return (__umulhi(n, f.mul) + n) >> f.shr;
#else
    return ((uint32_t)(((uint64_t)n * f.mul) >> 32) + n) >> f.shr;   // host twin: tests/decode_check.cu walks the tile maps on the CPU
#endif
}

struct alignas(64) IgemmProblem {
    CUtensorMap tmA;        // activations, 4-D (C, W, H, B)
    CUtensorMap tmB;        // weights, 2-D (K, N)
    const float* bias;      // [N] or [B][bias_img_stride]
    void* out;              // GATE: skip tensor, scaled in place
    const float* vec;       // GATE: w_psi[BN]; OUTCONV: w_out[BN]
    float* aux;             // GATE: psi map (may be null); OUTCONV: logits
    int H, W;               // pixel grid of the GEMM rows
    int tiles_x, tiles_per_img, m_tiles, n_tiles, tile_begin;
    FastDiv fd_n_tiles, fd_tiles_per_img, fd_tiles_x;
    int taps, dil, nchunk;  // taps in {1, 9}; nchunk = Cin / KC
    int epi, relu, bias_img_stride;
    int outH, outW, out_ld, out_choff;
    int convt_cout;
    int gate_C, gate_plus_x;
    float scalar;           // GATE: b_psi; OUTCONV: b_out
    int fix_axis;           // CONVTFIX: 0 = rows (H), 1 = columns (W)
    int fix_out;            // CONVTFIX: output size along the fixed axis (2n + 1)
    const float* fix_coef;  // CONVTFIX: [fix_out][3] blend weights on (up, mid0, mid1)
};

struct alignas(64) IgemmParams {
    IgemmProblem prob[IGEMM_MAX_PROBLEMS];
    CUtensorMap tmC[4];     // output maps: STORE -> one per problem; CONVT -> one per (a,b) of the single problem
    CUtensorMap tmP;        // fused MaxPool2d(2) output (pool != 0, single-problem STORE launches)
    int nprob, total_tiles;
    int amode;              // AMODE_*
    int KC;                 // channels per sub-block (16 / 32 / 64)
    int TW, TH, tw_shift;
    int MT;                 // M-blocks (128 pixels each, stacked vertically) per tile sharing every B sub-block: 1, or 2 (SLAB, BN <= 128)
    int BN, CB;             // MMA N (DXN: 3*Cout); channels per TMA store (CB*2 bytes == the store swizzle width)
    int n_out;              // output channels per tile (== BN except DXN: BN / 3)
    int VW;                 // valid output columns per tile (== TW except DXN: TW - 2)
    int nA, nB, b_resident; // ring depths; b_resident: nB == number of k-steps and B is loaded once
    int a_slot_bytes, b_slot_bytes, c_slot_bytes;
    int cslots;             // staging tiles per epilogue group (2: a TMA store drains while the next tile is being staged)
    int cbatch;             // cslots == chunks per tile: stage the whole tile, then one fence / barrier / store group
    int pool, p_slot_bytes; // fused 2x2 max-pool of the stored tile (floor semantics), its staging slot size
    int b_region_bytes;     // nB * b_slot_bytes rounded up to 1024 (the staging tiles behind it need that alignment)
    int tmem_cols;
    int acc_stages;         // TMEM accumulator stages (2, or 1 when 2*BN columns would leave no room for a second CTA)
    int is_fp16;
    int lean_sync;          // 1: drop the per-tile top barrier where a static bias and alternating staging tiles allow it
    int tile_iter;          // 1: incremental tile coordinates (single-problem launches), 0: full decode per tile
    int pair_order;         // CTA pairs over several N tiles: consecutive tiles are two M tiles of one N tile
    int TB;                 // images per tile (per-tap staging of small images: the TMA boxes span TB frames of TH rows each)
    int fault_inject;       // test hook: the TMA producer of CTA 0 never loads anything, so its MMA warp runs into the bounded wait
    int fix_compact;        // CONVTFIX with resident weights: only the non-zero blocks are kept / multiplied (tap -1: `up`, tap 0: `mid0 | mid1`)
    int skip_oob;           // per-tap staging, 3x3: taps whose whole box lies outside the image (dilated ASPP branches) are not issued
    int* err;
};

// ---- 16-bit activation helpers (bf16 by default, fp16 as the higher-precision storage option) -------------
// fp16 storage saturates to +-65504 (cvt.satfinite) instead of overflowing to inf: one instruction either way
__device__ __forceinline__ uint32_t pack2(float a, float b, int is_fp16) {
    if (is_fp16) {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        return r;
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// round-to-nearest pack with the ReLU clamp folded into the conversion instruction (cvt.rn.relu.*x2.f32)
template <bool F16>
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t r;
    if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else     asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack2(uint32_t v, int is_fp16) {
    if (is_fp16) return __half22float2(*reinterpret_cast<__half2*>(&v));
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b, int is_fp16) {
    if (is_fp16) {
        __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b, int f) {
    return make_uint4(max2(a.x, b.x, f), max2(a.y, b.y, f), max2(a.z, b.z, f), max2(a.w, b.w, f));
}
// 16-byte shared-memory accesses through 32-bit shared addresses (a generic pointer costs 64-bit address arithmetic
// and a generic ST/LD per access in the epilogue's innermost loops)
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
struct TileCoord { int pi, b, y0, x0, n0; };

__host__ __device__ __forceinline__ TileCoord decode_tile(const IgemmParams& P, int t) {
    TileCoord tc;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < IGEMM_MAX_PROBLEMS; ++i)
        if (i < P.nprob && t >= P.prob[i].tile_begin) pi = i;
    const IgemmProblem& q = P.prob[pi];
    const int local = t - q.tile_begin;
    int mt = (int)fdiv((uint32_t)local, q.fd_n_tiles);
    int nt = local - mt * q.n_tiles;
    if (P.pair_order) {                                           // CTA pairs with several N tiles: tiles 2k, 2k+1 = M tiles 2j, 2j+1 of ONE N tile
        const int k = local >> 1;
        const int j = (int)fdiv((uint32_t)k, q.fd_n_tiles);
        nt = k - j * q.n_tiles;
        mt = 2 * j + (local & 1);
    }
    tc.pi = pi;
    const int bt = (int)fdiv((uint32_t)mt, q.fd_tiles_per_img);
    tc.b = bt * P.TB;
    const int r = mt - bt * q.tiles_per_img;
    const int tyi = (int)fdiv((uint32_t)r, q.fd_tiles_x);
    tc.y0 = tyi * P.TH * P.MT;
    tc.x0 = (r - tyi * q.tiles_x) * P.VW - (P.amode >= AMODE_DXN ? 1 : 0);   // DXN / RS: slab column 0 is the left halo
    tc.n0 = nt * P.n_out;
    return tc;
}


// Tile walk of one role: t = first, first + stride, ...  For single-problem launches the (n tile, x, y, frame)
// coordinates advance by the stride's own mixed-radix digits with carries (a handful of adds per tile); the three
// divisions are done once.  The roles are single warps running serial instruction streams, and for the small-K
// layers a full decode per tile was a visible share of the MMA warp's time per tile.
struct TileIter {
    int t, stride, total;
    int nt, x, y, b;          // current digits
    int s_nt, s_x, s_y, s_b;  // digits of the stride
    int tiles_x, tiles_y, n_tiles;
    bool incremental;
    __host__ __device__ __forceinline__ void init(const IgemmParams& P, int first, int stride_) {
        t = first; stride = stride_; total = P.total_tiles;
        incremental = P.nprob == 1 && P.tile_iter != 0 && P.pair_order == 0;
        if (incremental) {
            const IgemmProblem& q = P.prob[0];
            tiles_x = q.tiles_x; n_tiles = q.n_tiles;
            tiles_y = (int)fdiv((uint32_t)q.tiles_per_img, q.fd_tiles_x);
            auto digits = [&](int v, int& dn, int& dx, int& dy, int& db) {
                const int m = (int)fdiv((uint32_t)v, q.fd_n_tiles);
                dn = v - m * q.n_tiles;
                db = (int)fdiv((uint32_t)m, q.fd_tiles_per_img);
                const int r = m - db * q.tiles_per_img;
                dy = (int)fdiv((uint32_t)r, q.fd_tiles_x);
                dx = r - dy * q.tiles_x;
            };
            digits(first, nt, x, y, b);
            digits(stride_, s_nt, s_x, s_y, s_b);
        }
    }
    __host__ __device__ __forceinline__ bool valid() const { return t < total; }
    __host__ __device__ __forceinline__ void next() {
        t += stride;
        if (incremental) {
            nt += s_nt;
            int c = nt >= n_tiles ? 1 : 0;
            nt -= c ? n_tiles : 0;
            x += s_x + c;
            c = x >= tiles_x ? 1 : 0;
            x -= c ? tiles_x : 0;
            y += s_y + c;
            c = y >= tiles_y ? 1 : 0;
            y -= c ? tiles_y : 0;
            b += s_b + c;
        }
    }
    __host__ __device__ __forceinline__ TileCoord coord(const IgemmParams& P) const {
        if (!incremental) return decode_tile(P, t);
        TileCoord tc;
        tc.pi = 0;
        tc.b = b * P.TB;
        tc.y0 = y * P.TH * P.MT;
        tc.x0 = x * P.VW - (P.amode >= AMODE_DXN ? 1 : 0);
        tc.n0 = nt * P.n_out;
        return tc;
    }
};

// MaxPool2d(2) of a staged output tile (rows = th x vw pixels, c_pitch bytes each, TMA-swizzled) into a staged
// pooled tile ((th/2) x (vw/2) pixels, same pitch / swizzle).  One 16-byte vector (8 channels) per thread-iteration;
// c_pitch / 16 is a power of two, so a thread keeps its vector index and walks pooled pixels.
__device__ __forceinline__ void pool_staged_tile(uint32_t cs, uint32_t ps, int th, int vw, int c_pitch, uint32_t swz_mask,
                                                 int etid, int f16) {
    const int PW = vw >> 1, PH = th >> 1;
    const int lcv = 31 - __clz(c_pitch >> 4);                     // log2(vectors per pixel): 1, 2 or 3
    const int v = etid & ((1 << lcv) - 1), i0 = etid >> lcv, istep = 128 >> lcv;    // 128 threads per epilogue group
    // pooled pixels are dealt out in one flat sequence (i = py * PW + px), so a 2 x 15 pooled tile occupies 30 of the 32
    // thread slots once instead of 15 slots twice; i / PW by multiply-shift (exact for i * PW < 2^16)
    const uint32_t inv_pw = (65536u + (uint32_t)PW - 1u) / (uint32_t)PW;
    for (int i = i0; i < PH * PW; i += istep) {
        {
            const int py = (int)(((uint32_t)i * inv_pw) >> 16), px = i - py * PW;
            const int r00 = (2 * py) * vw + 2 * px;
            uint32_t o0 = (uint32_t)(r00 * c_pitch + v * 16), o1 = o0 + (uint32_t)c_pitch;
            uint32_t o2 = o0 + (uint32_t)(vw * c_pitch), o3 = o2 + (uint32_t)c_pitch;
            o0 ^= ((o0 >> 7) & swz_mask) << 4; o1 ^= ((o1 >> 7) & swz_mask) << 4;
            o2 ^= ((o2 >> 7) & swz_mask) << 4; o3 ^= ((o3 >> 7) & swz_mask) << 4;
            if (c_pitch == 64 && (i & 1)) {
                // 64-byte pixels: the left pixel of every 2x2 window sits in banks 0-15, the right one in banks 16-31, and the
                // 8 threads of a quarter-warp cover two windows -- odd windows read right-then-left so that each LDS.128 spreads
                // over all 32 banks instead of queueing two-deep on 16 (ncu: 2x excessive wavefronts on d1.1's four pool loads)
                uint32_t t = o0; o0 = o1; o1 = t;
                t = o2; o2 = o3; o3 = t;
            }
            const uint4 m = max8(max8(lds128(cs + o0), lds128(cs + o1), f16), max8(lds128(cs + o2), lds128(cs + o3), f16), f16);
            uint32_t po = (uint32_t)((py * PW + px) * c_pitch + v * 16);
            po ^= ((po >> 7) & swz_mask) << 4;
            sts128(ps + po, m);
        }
    }
}

__device__ __forceinline__ void tap_offsets(const IgemmProblem& q, int tap, int& dy, int& dx) {
    dy = 0; dx = 0;
    if (q.taps == 9) { dy = (tap / 3 - 1) * q.dil; dx = (tap % 3 - 1) * q.dil; }
    else if (q.taps == 2) { dy = q.fix_axis == 0 ? tap - 1 : 0; dx = q.fix_axis == 1 ? tap - 1 : 0; }
}

// Per-tap staging of a (dilated) 3x3 convolution: a tap whose whole TH x TW box lies outside the image would read nothing but
// TMA zero fill and add exact zeros -- neither its boxes nor its weight sub-blocks are fetched, none of its MMAs issued.  At
// 35 x 46 this is 58 % of the taps of the dilation-18 branch and 22 % of the dilation-12 one; a 1x1 convolution embedded as
// the centre tap of a 3x3 with a dilation beyond the image (how ASPP's blocks.0 joins the dilated branches' launch) keeps
// exactly its one tap.  Producer and MMA issuer evaluate the same predicate on the same tile, so the rings stay in step; the
// two CTAs of a pair share every MMA and every barrier, so there a tap goes only when it lies outside BOTH their tiles
// (`tcp` = the peer's tile, tile index ^ 1).
__device__ __forceinline__ bool box_outside(const IgemmParams& P, const IgemmProblem& q, const TileCoord& tc, int dy, int dx) {
    const int ylo = tc.y0 + dy, xlo = tc.x0 + dx;
    return ylo >= q.H || ylo + P.TH <= 0 || xlo >= q.W || xlo + P.TW <= 0;
}
__device__ __forceinline__ bool tap_outside(const IgemmParams& P, const IgemmProblem& q, const TileCoord& tc, const TileCoord& tcp, int dy, int dx) {
    if (!P.skip_oob || q.taps != 9) return false;
    return box_outside(P, q, tc, dy, dx) && box_outside(P, q, tcp, dy, dx);
}

// KK tcgen05.mma (K = 16 each) over one KC-wide sub-block.  Descriptors only differ in their 14-bit start-address
// field, so stepping K by 32 bytes is "+2" on the low word.
template <int KK, bool PAIR = false>
__device__ __forceinline__ void mma_subblock(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
#pragma unroll
    for (int k = 0; k < KK; ++k) {
        const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
        const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
        if (PAIR) ptx::umma_f16_pair(d_tmem, da, db, idesc, accumulate);
        else      ptx::umma_f16(d_tmem, da, db, idesc, accumulate);
        accumulate = 1;
    }
}

// The MMA role is executed by ALL 32 lanes of warp 1 with warp-uniform control flow (the warp index is broadcast
// with a shuffle so that the compiler can prove it): every mbarrier wait is done by the whole warp and only the
// tcgen05.mma / tcgen05.commit instructions sit under elect.sync.  Written as `if (threadIdx.x == 32)` the compiler
// must assume divergence, wraps every UTCHMMA in an elect/branch "waterfall" loop and moves each descriptor from
// vector to uniform registers first (R2UR): ~70-100 cycles per MMA instead of the 40-48 cycle hardware floor of
// the small-N layers (profiles/r01_mma_probe.txt).
template <int KK, bool MULTI, bool PAIR, int AM>
__device__ __forceinline__ void mma_role(const IgemmParams& P, uint8_t* smem_a, uint8_t* smem_b, uint64_t* full_a_p, uint64_t* empty_a_p,
                                         uint64_t* full_b_p, uint64_t* empty_b_p, uint64_t* b_res_bar, uint64_t* tmem_full_p,
                                         uint64_t* tmem_empty_p, uint32_t tmem_base) {
    // barrier addresses once, opaque, as 32-bit shared addresses (slot s = base + 8 s): formed at each use they cost an
    // S2UR SR_CgaCtaId + three dependent uniform ops in front of every wait / commit of this serial instruction stream
    const uint32_t full_a = ptx::keep_u32(ptx::smem_u32(full_a_p)), empty_a = ptx::keep_u32(ptx::smem_u32(empty_a_p));
    const uint32_t full_b = ptx::keep_u32(ptx::smem_u32(full_b_p)), empty_b = ptx::keep_u32(ptx::smem_u32(empty_b_p));
    const uint32_t tmem_full_bar = ptx::keep_u32(ptx::smem_u32(tmem_full_p)), tmem_empty_bar = ptx::keep_u32(ptx::smem_u32(tmem_empty_p));
    const int amode = AM >= 0 ? AM : P.amode;                              // compile-time in the specialised instantiations
    const int swz = P.KC * 2;
    const uint32_t idesc = ptx::make_idesc_f16(PAIR ? 256 : 128, P.BN, P.is_fp16 != 0);   // pair: M = 256 over the two CTAs
    const uint64_t proto = ptx::make_kmajor_desc(0, swz);
    const uint32_t desc_hi = (uint32_t)(proto >> 32);
    const uint32_t lo_flags = (uint32_t)proto;                               // LBO field; start address = 0
    const uint32_t a_base = lo_flags | ((ptx::smem_u32(smem_a) & 0x3FFFFu) >> 4);
    const uint32_t b_base = lo_flags | ((ptx::smem_u32(smem_b) & 0x3FFFFu) >> 4);
    const uint32_t a_slot16 = (uint32_t)P.a_slot_bytes >> 4, b_slot16 = (uint32_t)P.b_slot_bytes >> 4;
    const uint32_t a_dy16 = (uint32_t)(P.TW * swz) >> 4;                      // slab: next vertical tap = TW rows further
    const uint32_t a_mb16 = (uint32_t)(P.TH * P.TW * swz) >> 4;               // slab: second M-block = TH image rows further
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool res = P.b_resident != 0;
    if (res) ptx::mbar_wait(b_res_bar, 0, P.err, ERR_MMA_WAIT_FULL);
    TileIter it;
    TM_DECL();
    for (it.init(P, blockIdx.x, gridDim.x); it.valid(); it.next()) {
        const TileCoord tc = it.coord(P);
        const IgemmProblem& q = MULTI ? P.prob[tc.pi] : P.prob[0];   // single-problem launches: fixed parameter offsets
        TM_MARK(2);
        ptx::mbar_wait(tmem_empty_bar + 8u * (uint32_t)acc, acc_phase ^ 1, P.err, ERR_MMA_WAIT_TMEM);
        ptx::tc_fence_after();
        TM_MARK(0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P.BN * P.MT);
        uint32_t accumulate = 0;
        if (amode == AMODE_TAP) {
            const int steps = q.taps * q.nchunk;
            const TileCoord tcp = (PAIR && P.skip_oob) ? decode_tile(P, it.t ^ 1) : tc;
            for (int s = 0; s < steps; ++s) {
                if (P.skip_oob) {                                             // (same predicate as the producers')
                    int dy, dx;
                    tap_offsets(q, s / q.nchunk, dy, dx);
                    if (tap_outside(P, q, tc, tcp, dy, dx)) continue;
                }
                ptx::mbar_wait(full_a + 8u * (uint32_t)ia, pa, P.err, ERR_MMA_WAIT_FULL);
                const int bslot = res ? s : ib;
                if (!res) ptx::mbar_wait(full_b + 8u * (uint32_t)ib, pb, P.err, ERR_MMA_WAIT_FULL);
                ptx::tc_fence_after();
                if (P.fix_compact) {                                          // CONVTFIX: tap -1 feeds `up` only, tap 0 feeds `mid0 | mid1`
                    if (ptx::elect_one()) {
                        const int CC = P.BN / 3, tap = s >= q.nchunk ? 1 : 0, ch = s - tap * q.nchunk;
                        const uint32_t up16 = (uint32_t)(CC * swz) >> 4;
                        const uint32_t b_lo = b_base + (tap ? (uint32_t)q.nchunk * up16 + (uint32_t)ch * 2u * up16 : (uint32_t)ch * up16);
                        mma_subblock<KK, PAIR>(d_tmem + (tap ? (uint32_t)CC : 0u), a_base + ia * a_slot16, b_lo, desc_hi,
                                               ptx::make_idesc_f16(128, tap ? 2 * CC : CC, P.is_fp16 != 0), ch == 0 ? 0u : 1u);
                        ptx::umma_commit(empty_a + 8u * (uint32_t)ia);
                    }
                    __syncwarp();
                    if (++ia == P.nA) { ia = 0; pa ^= 1; }
                    continue;
                }
                if (ptx::elect_one()) {
                    mma_subblock<KK, PAIR>(d_tmem, a_base + ia * a_slot16, b_base + bslot * b_slot16, desc_hi, idesc, accumulate);
                    if (PAIR) { ptx::umma_commit_pair(empty_a + 8u * (uint32_t)ia); if (!res) ptx::umma_commit_pair(empty_b + 8u * (uint32_t)ib); }
                    else      { ptx::umma_commit(empty_a + 8u * (uint32_t)ia); if (!res) ptx::umma_commit(empty_b + 8u * (uint32_t)ib); }
                }
                __syncwarp();
                accumulate = 1;
                if (++ia == P.nA) { ia = 0; pa ^= 1; }
                if (!res && ++ib == P.nB) { ib = 0; pb ^= 1; }
            }
        } else if (amode == AMODE_RS) {
            const uint32_t row16 = (uint32_t)swz >> 4;                        // one pixel row of the slab, in 16-byte units
            for (int ch = 0; ch < q.nchunk; ++ch) {
                TM_MARK(2);
                ptx::mbar_wait(full_a + 8u * (uint32_t)ia, pa, P.err, ERR_MMA_WAIT_FULL);
                ptx::tc_fence_after();
                TM_MARK(1);
                if (ptx::elect_one()) {
                    const uint32_t a_lo = a_base + ia * a_slot16;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {                       // B resident, TAP order: slot tap*nchunk + ch
                        const uint32_t shift = (uint32_t)((tap / 3) * P.TW + (tap % 3)) * row16;
                        const uint32_t b_lo = b_base + (uint32_t)(tap * q.nchunk + ch) * b_slot16;
                        if (P.MT == 2)                                        // second M-block: TH rows further down the slab
                            mma_subblock<KK, PAIR>(d_tmem + P.BN, a_lo + shift + a_mb16, b_lo, desc_hi, idesc, tap == 0 ? accumulate : 1u);
                        mma_subblock<KK, PAIR>(d_tmem, a_lo + shift, b_lo, desc_hi, idesc, tap == 0 ? accumulate : 1u);
                    }
                    if (PAIR) ptx::umma_commit_pair(empty_a + 8u * (uint32_t)ia); else ptx::umma_commit(empty_a + 8u * (uint32_t)ia);
                }
                __syncwarp();
                TM_MARK(3);
                accumulate = 1;
                if (++ia == P.nA) { ia = 0; pa ^= 1; }
            }
        } else {
            int step = 0;
            const int ndc = (amode == AMODE_SLAB ? 3 : 1) * q.nchunk;
            for (int dc = 0; dc < ndc; ++dc) {                                // SLAB: (dx, channel chunk); DXN: channel chunk
                TM_MARK(2);
                ptx::mbar_wait(full_a + 8u * (uint32_t)ia, pa, P.err, ERR_MMA_WAIT_FULL);
                TM_MARK(1);
                const uint32_t a_lo = a_base + ia * a_slot16;
                if (res) {
                    // resident weights: nothing to wait for between the vertical taps, so all of them go out under one
                    // election (a fence + elect + warp sync per four MMAs kept the issue rate of the one-CTA-per-SM
                    // dx-stacked layer at 87 cycles per MMA against the 56-cycle operand-read floor)
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            const uint32_t b_lo = b_base + (uint32_t)(step + dyi) * b_slot16;
                            if (P.MT == 2)
                                mma_subblock<KK, PAIR>(d_tmem + P.BN, a_lo + dyi * a_dy16 + a_mb16, b_lo, desc_hi, idesc, dyi == 0 ? accumulate : 1u);
                            mma_subblock<KK, PAIR>(d_tmem, a_lo + dyi * a_dy16, b_lo, desc_hi, idesc, dyi == 0 ? accumulate : 1u);
                        }
                        if (PAIR) ptx::umma_commit_pair(empty_a + 8u * (uint32_t)ia); else ptx::umma_commit(empty_a + 8u * (uint32_t)ia);
                    }
                    __syncwarp();
                    step += 3;
                    accumulate = 1;
                    TM_MARK(3);
                    if (++ia == P.nA) { ia = 0; pa ^= 1; }
                    continue;
                }
#pragma unroll
                for (int dyi = 0; dyi < 3; ++dyi, ++step) {
                    const int bslot = res ? step : ib;
                    if (!res) ptx::mbar_wait(full_b + 8u * (uint32_t)ib, pb, P.err, ERR_MMA_WAIT_FULL);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        if (P.MT == 2) {                                      // second M-block: TH rows further down the slab
                            uint32_t acc2 = accumulate;
                            mma_subblock<KK, PAIR>(d_tmem + P.BN, a_lo + dyi * a_dy16 + a_mb16, b_base + bslot * b_slot16, desc_hi, idesc, acc2);
                        }
                        mma_subblock<KK, PAIR>(d_tmem, a_lo + dyi * a_dy16, b_base + bslot * b_slot16, desc_hi, idesc, accumulate);
                        if (!res) { if (PAIR) ptx::umma_commit_pair(empty_b + 8u * (uint32_t)ib); else ptx::umma_commit(empty_b + 8u * (uint32_t)ib); }
                        if (dyi == 2) { if (PAIR) ptx::umma_commit_pair(empty_a + 8u * (uint32_t)ia); else ptx::umma_commit(empty_a + 8u * (uint32_t)ia); }
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (!res && ++ib == P.nB) { ib = 0; pb ^= 1; }
                }
                TM_MARK(3);
                if (++ia == P.nA) { ia = 0; pa ^= 1; }
            }
        }
        if (ptx::elect_one()) {                                               // accumulator complete -> epilogue (of both CTAs of a pair)
            if (PAIR) ptx::umma_commit_pair(tmem_full_bar + 8u * (uint32_t)acc);
            else      ptx::umma_commit(tmem_full_bar + 8u * (uint32_t)acc);
        }
        __syncwarp();
        if (++acc == P.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    TM_MARK(2);
    TM_FLUSH(0, (threadIdx.x & 31) == 0);
}

// NG = epilogue groups = TMEM accumulator stages (2, or 4 for single-CTA-per-SM layers whose 4 accumulators fit in
// the 512 TMEM columns: there the per-tile epilogue latency chain, not the tensor pipe, sets the pace).
// MULTI = the launch holds several problems (the three dilated ASPP branches); single-problem launches read their
// problem record at fixed parameter offsets (uniform constant loads the compiler can hoist out of the tile loops).
// PAIR = launched as clusters of two CTAs that share every MMA (cta_group::2, M = 256): each CTA stages its own A slabs
// and HALF of each weight sub-block, so the L2 -> SM weight traffic and the shared-memory operand reads per SM drop.
// Used for the slab-staged layers whose weights stream (N = 128 / 256); the leader (cluster rank 0) issues the MMAs,
// both CTAs run their own producer and epilogue groups on their own 128 rows.
template <int NG, bool F16, bool MULTI, bool PAIR = false, int AM = -1, int EP = -1, int KKT = -1, int PL = -1>
__global__ void __launch_bounds__(igemm_threads(NG), NG == 2 ? 2 : 1) igemm_tc_kernel(const __grid_constant__ IgemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[IGEMM_MAX_SLOTS], empty_a[IGEMM_MAX_SLOTS];
    __shared__ __align__(8) uint64_t full_b[IGEMM_MAX_SLOTS], empty_b[IGEMM_MAX_SLOTS];
    __shared__ __align__(8) uint64_t b_res_bar, c_load_bar[NG];
    __shared__ __align__(8) uint64_t tmem_full_bar[NG], tmem_empty_bar[NG];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[1024];            // [epilogue group][tile parity][512 / NG channels]
    __shared__ __align__(16) float s_vec[256];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform, and provably so for the compiler
    const int lane = threadIdx.x & 31;
    const int amode = AM >= 0 ? AM : P.amode;               // AM / EP >= 0: staging mode / epilogue fixed at compile time
#define EPI_OF(q_) (EP >= 0 ? EP : (q_).epi)
    const bool fused_pool = PL >= 0 ? (PL != 0) : (P.pool != 0);   // KKT / PL >= 0: MMAs per sub-block / fused MaxPool fixed at compile time
    // operand tiles need 1024-byte alignment for the 128-byte swizzle
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + (size_t)P.nA * P.a_slot_bytes;
    uint8_t* smem_c = smem_b + (size_t)P.b_region_bytes;
    uint8_t* smem_p = smem_c + (size_t)(NG * P.cslots) * P.c_slot_bytes;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.nprob; ++i) {
            ptx::prefetch_tmap(&P.prob[i].tmA);
            ptx::prefetch_tmap(&P.prob[i].tmB);
        }
        for (int s = 0; s < P.nA; ++s) { ptx::mbar_init(&full_a[s], 1); ptx::mbar_init(&empty_a[s], 1); }
        if (!P.b_resident)
            for (int s = 0; s < P.nB; ++s) { ptx::mbar_init(&full_b[s], 1); ptx::mbar_init(&empty_b[s], 1); }
        ptx::mbar_init(&b_res_bar, 1);
        for (int a = 0; a < NG; ++a) {
            ptx::mbar_init(&c_load_bar[a], 1);
            ptx::mbar_init(&tmem_full_bar[a], 1);
            ptx::mbar_init(&tmem_empty_bar[a], PAIR ? 8 : 4);   // one arrive per epilogue warp of the owning group (of both CTAs)
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        if (PAIR) { ptx::tmem_alloc_pair(&tmem_base_smem, (uint32_t)P.tmem_cols); ptx::tmem_relinquish_pair(); }
        else      { ptx::tmem_alloc(&tmem_base_smem, (uint32_t)P.tmem_cols); ptx::tmem_relinquish(); }
    }
    const uint32_t pair_rank = PAIR ? ptx::cluster_ctarank() : 0u;
    if (threadIdx.x >= 64 && P.prob[0].vec != nullptr)                        // GATE / OUTCONV vector, constant per launch
        for (int i = threadIdx.x - 64; i < P.n_out; i += 128 * NG) s_vec[i] = P.prob[0].vec[i];
    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR) ptx::cluster_sync_all();                      // the peer's barriers exist before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // Programmatic dependent launch: let the next kernel of the stream be scheduled onto SMs as our CTAs retire (its
    // barrier / TMEM set-up and resident-weight load then overlap our tail); every role that touches activations
    // first executes griddepcontrol.wait, which returns once the PREVIOUS kernel has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // =========================== TMA producer ===========================
        // whole warp, warp-uniform control flow; the TMA instructions themselves are issued under elect.sync
        if (!(P.fault_inject && blockIdx.x == 0)) {   // (fault injection, tests only: a silent producer must end in a reported trap, not a hang)
            int ia = 0, ib = 0;
            uint32_t pa = 0, pb = 0;
            const bool res = P.b_resident != 0;
            if (res) {                                                        // whole weight matrix once per CTA
                const IgemmProblem& q = P.prob[0];
                const int cin = q.nchunk * P.KC;
                const int steps = (amode == AMODE_DXN ? 3 : q.taps) * q.nchunk;
                if (P.fix_compact) {
                    // transposed conv + fix-up: of the [up | mid0 | mid1] x (tap -1, tap 0) weight tile only `up` x tap -1 and
                    // `mid0 | mid1` x tap 0 are non-zero (prep_upfix): the tap -1 sub-blocks keep CC rows, the tap 0 ones 2 CC
                    // (half the shared memory of the padded tile -- it goes to the A ring -- and 42 % fewer MMA cycles)
                    if (ptx::elect_one()) {
                        const int CC = P.BN / 3, n0 = (int)(blockIdx.x % q.n_tiles) * P.BN;
                        const uint32_t up_bytes = (uint32_t)(CC * P.KC * 2);
                        ptx::mbar_expect_tx(&b_res_bar, (uint32_t)q.nchunk * 3u * up_bytes);
                        for (int ch = 0; ch < q.nchunk; ++ch) {
                            ptx::tma_load_2d(smem_b + (size_t)ch * up_bytes, &q.tmB, &b_res_bar, ch * P.KC, n0);
                            uint8_t* mid = smem_b + (size_t)q.nchunk * up_bytes + (size_t)ch * 2 * up_bytes;
                            ptx::tma_load_2d(mid, &q.tmB, &b_res_bar, (q.nchunk + ch) * P.KC, n0 + CC);
                            ptx::tma_load_2d(mid + up_bytes, &q.tmB, &b_res_bar, (q.nchunk + ch) * P.KC, n0 + 2 * CC);
                        }
                    }
                } else if (ptx::elect_one()) {
                    // pair: each CTA keeps HALF of the weight rows of every sub-block; both halves complete the leader's barrier
                    if (!PAIR || pair_rank == 0) ptx::mbar_expect_tx(&b_res_bar, (uint32_t)((PAIR ? 2 : 1) * steps * P.b_slot_bytes));
                    for (int s = 0; s < steps; ++s) {
                        int kcoord = s * P.KC;                                // TAP / RS order: (tap, chunk)
                        if (amode == AMODE_SLAB) {                          // SLAB order: (dx, chunk, dy)
                            const int dyi = s % 3, dc = s / 3;
                            const int dxi = dc / q.nchunk, ch = dc - dxi * q.nchunk;
                            kcoord = (dyi * 3 + dxi) * cin + ch * P.KC;
                        } else if (amode == AMODE_DXN) {                    // DXN order: (chunk, dy); K' = (dy, Cin)
                            const int dyi = s % 3, ch = s / 3;
                            kcoord = dyi * cin + ch * P.KC;
                        }
                        if (PAIR) ptx::tma_load_2d_pair(smem_b + (size_t)s * P.b_slot_bytes, &q.tmB, &b_res_bar, kcoord, (int)pair_rank * (P.BN / 2));
                        else      ptx::tma_load_2d(smem_b + (size_t)s * P.b_slot_bytes, &q.tmB, &b_res_bar, kcoord, (int)(blockIdx.x % q.n_tiles) * P.BN);
                    }
                }
                __syncwarp();
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");                // weights are constants; activations are not
            TileIter it;
            for (it.init(P, blockIdx.x, gridDim.x); it.valid(); it.next()) {
                const TileCoord tc = it.coord(P);
                const IgemmProblem& q = MULTI ? P.prob[tc.pi] : P.prob[0];   // single-problem launches: fixed parameter offsets
                if (amode == AMODE_TAP) {
                    const int steps = q.taps * q.nchunk;
                    const TileCoord tcp = (PAIR && P.skip_oob) ? decode_tile(P, it.t ^ 1) : tc;
                    for (int s = 0; s < steps; ++s) {
                        const int tap = s / q.nchunk;
                        const int ch = s - tap * q.nchunk;
                        int dy, dx;
                        tap_offsets(q, tap, dy, dx);
                        if (tap_outside(P, q, tc, tcp, dy, dx)) continue;     // all zero fill: the MMA issuer skips it too
                        ptx::mbar_wait(&empty_a[ia], pa ^ 1, P.err, ERR_PRODUCER_WAIT);
                        if (!res) ptx::mbar_wait(&empty_b[ib], pb ^ 1, P.err, ERR_PRODUCER_WAIT);
                        if (ptx::elect_one()) {
                            if (PAIR) {                                       // both CTAs' boxes complete the LEADER's barriers
                                if (pair_rank == 0) {
                                    ptx::mbar_expect_tx(&full_a[ia], 2 * (uint32_t)P.a_slot_bytes);
                                    if (!res) ptx::mbar_expect_tx(&full_b[ib], 2 * (uint32_t)P.b_slot_bytes);
                                }
                                ptx::tma_load_4d_pair(smem_a + (size_t)ia * P.a_slot_bytes, &q.tmA, &full_a[ia], ch * P.KC, tc.x0 + dx, tc.y0 + dy, tc.b);
                                if (!res) ptx::tma_load_2d_pair(smem_b + (size_t)ib * P.b_slot_bytes, &q.tmB, &full_b[ib], s * P.KC,
                                                                tc.n0 + (int)pair_rank * (P.BN / 2));   // this CTA's half of the weight rows
                            } else {
                                ptx::mbar_expect_tx(&full_a[ia], (uint32_t)P.a_slot_bytes);
                                ptx::tma_load_4d(smem_a + (size_t)ia * P.a_slot_bytes, &q.tmA, &full_a[ia], ch * P.KC, tc.x0 + dx, tc.y0 + dy, tc.b);
                                if (!res) {
                                    ptx::mbar_expect_tx(&full_b[ib], (uint32_t)P.b_slot_bytes);
                                    ptx::tma_load_2d(smem_b + (size_t)ib * P.b_slot_bytes, &q.tmB, &full_b[ib], s * P.KC, tc.n0);
                                }
                            }
                        }
                        __syncwarp();
                        if (++ia == P.nA) { ia = 0; pa ^= 1; }
                        if (!res && ++ib == P.nB) { ib = 0; pb ^= 1; }
                    }
                } else {
                    const int cin = q.nchunk * P.KC;
                    const uint32_t slab_bytes = (uint32_t)((P.TH * P.MT + 2) * P.TW * P.KC * 2);
                    const bool dxn = amode == AMODE_DXN;
                    const bool one_slab = amode != AMODE_SLAB;              // DXN / RS: a single slab per channel chunk
                    for (int dxi = 0; dxi < (one_slab ? 1 : 3); ++dxi) {
                        for (int ch = 0; ch < q.nchunk; ++ch) {
                            ptx::mbar_wait(&empty_a[ia], pa ^ 1, P.err, ERR_PRODUCER_WAIT);
                            if (ptx::elect_one()) {
                                if (PAIR) {                                   // both CTAs' slabs complete the LEADER's barrier
                                    if (pair_rank == 0) ptx::mbar_expect_tx(&full_a[ia], 2 * slab_bytes);
                                    ptx::tma_load_4d_pair(smem_a + (size_t)ia * P.a_slot_bytes, &q.tmA, &full_a[ia], ch * P.KC,
                                                          one_slab ? tc.x0 : tc.x0 + dxi - 1, tc.y0 - 1, tc.b);
                                } else {
                                    ptx::mbar_expect_tx(&full_a[ia], slab_bytes);
                                    ptx::tma_load_4d(smem_a + (size_t)ia * P.a_slot_bytes, &q.tmA, &full_a[ia], ch * P.KC,
                                                     one_slab ? tc.x0 : tc.x0 + dxi - 1, tc.y0 - 1, tc.b);
                                }
                            }
                            __syncwarp();
                            if (++ia == P.nA) { ia = 0; pa ^= 1; }
                            if (!res) {
                                for (int dyi = 0; dyi < 3; ++dyi) {
                                    ptx::mbar_wait(&empty_b[ib], pb ^ 1, P.err, ERR_PRODUCER_WAIT);
                                    if (ptx::elect_one()) {
                                        if (PAIR) {                           // this CTA's half of the weight rows
                                            if (pair_rank == 0) ptx::mbar_expect_tx(&full_b[ib], 2 * (uint32_t)P.b_slot_bytes);
                                            ptx::tma_load_2d_pair(smem_b + (size_t)ib * P.b_slot_bytes, &q.tmB, &full_b[ib],
                                                                  (dyi * 3 + dxi) * cin + ch * P.KC, tc.n0 + (int)pair_rank * (P.BN / 2));
                                        } else {
                                            ptx::mbar_expect_tx(&full_b[ib], (uint32_t)P.b_slot_bytes);
                                            ptx::tma_load_2d(smem_b + (size_t)ib * P.b_slot_bytes, &q.tmB, &full_b[ib],
                                                             dxn ? dyi * cin + ch * P.KC : (dyi * 3 + dxi) * cin + ch * P.KC,
                                                             dxn ? (tc.n0 / P.n_out) * P.BN : tc.n0);
                                        }
                                    }
                                    __syncwarp();
                                    if (++ib == P.nB) { ib = 0; pb ^= 1; }
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (!PAIR || pair_rank == 0) {                      // pair: only the leader issues MMAs
            if constexpr (KKT > 0) {
                mma_role<KKT, MULTI, PAIR, AM>(P, smem_a, smem_b, full_a, empty_a, full_b, empty_b, &b_res_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
            } else {
                if (P.KC == 64)      mma_role<4, MULTI, PAIR, AM>(P, smem_a, smem_b, full_a, empty_a, full_b, empty_b, &b_res_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
                else if (P.KC == 32) mma_role<2, MULTI, PAIR, AM>(P, smem_a, smem_b, full_a, empty_a, full_b, empty_b, &b_res_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
                else                 mma_role<1, MULTI, PAIR, AM>(P, smem_a, smem_b, full_a, empty_a, full_b, empty_b, &b_res_bar, tmem_full_bar, tmem_empty_bar, tmem_base);
            }
        }
    } else {
        // =========================== epilogue (warps 2..9) ===========================
        // Two independent epilogue groups of four warps (one warp per TMEM lane quarter, warp % 4): group g drains
        // accumulator stage g, i.e. every second tile of this CTA.  Per tile the epilogue is a latency chain
        // (tmem wait -> tcgen05.ld -> math -> smem -> proxy fence -> barrier -> TMA store), not a throughput problem
        // (ncu: profiles/r01_ncu_dxn_summary.txt), so two chains in flight per CTA (x 2 CTAs per SM) is what pays.
        // per-thread constants of the tile loop are made opaque (keep_*): left transparent, ptxas re-derives them from
        // SR_TID.X / SR_CgaCtaId / SR_SWINHI inside every tile (three S2R and five S2UR round trips per tile, ~8 % of the
        // epilogue's samples on d1.1 in profiles/r01_ncu_v14_igemm_all_launches.txt) instead of keeping seven registers
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int lane = ptx::keep_i32((int)(threadIdx.x & 31));
        const int row = ptx::keep_i32(quarter * 32 + lane);
        const int etid = ptx::keep_i32((int)((threadIdx.x - 64) & 127));   // 0..127 inside the group
        const int ty = ptx::keep_i32(row >> P.tw_shift);
        const int tx = ptx::keep_i32(row & (P.TW - 1));
        constexpr int f16 = F16 ? 1 : 0;
        const int c_pitch = P.CB * 2;                           // bytes per staged row == store swizzle width
        const uint32_t swz_mask = (uint32_t)(c_pitch >> 4) - 1; // 7 / 3 / 1 for 128 / 64 / 32-byte swizzle
        const int acc = grp;
        uint32_t acc_phase = 0;
        uint32_t c_phase = 0;
        int cslot = 0;                                          // this group's two staging tiles alternate per TMA store
        const uint32_t smem_c_u32 = ptx::keep_u32(ptx::smem_u32(smem_c)), smem_p_u32 = ptx::keep_u32(ptx::smem_u32(smem_p));
        const uint32_t c_grp = smem_c_u32 + (uint32_t)(grp * P.cslots * P.c_slot_bytes);   // this group's staging tiles
        const uint32_t p_grp = smem_p_u32 + (uint32_t)(grp * P.cslots * P.p_slot_bytes);
        uint32_t cs = c_grp, ps = p_grp;                        // current staging tile / pooled staging tile (shared addresses)
#define NEXT_CSLOT()                                                                   \
    do {                                                                               \
        cslot = (cslot + 1 == P.cslots) ? 0 : cslot + 1;                               \
        cs = c_grp + (uint32_t)(cslot * P.c_slot_bytes);                               \
        ps = p_grp + (uint32_t)(cslot * P.p_slot_bytes);                               \
    } while (0)
#define WAIT_STORE_READS()                                                             \
    do {                                                                               \
        if (P.cslots == 2 && !P.cbatch) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); \
        else                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); \
    } while (0)
        int par = 1;                                            // parity of this group's tile iteration
        const uint32_t cbar = ptx::keep_u32(ptx::smem_u32(&c_load_bar[grp]));
        const uint32_t full_bar = ptx::keep_u32(ptx::smem_u32(&tmem_full_bar[acc]));
        const uint32_t empty_bar = ptx::keep_u32(PAIR ? (ptx::smem_u32(&tmem_empty_bar[acc]) & ptx::kPeerBitMask) : ptx::smem_u32(&tmem_empty_bar[acc]));
        // the bias slice of a CTA never changes when there is one problem, no per-image bias and one N tile per CTA:
        // it is staged once instead of once per tile
        const bool bias_static = P.nprob == 1 && P.prob[0].bias_img_stride == 0 && (P.prob[0].n_tiles == 1 || P.b_resident != 0);
        if (bias_static) {
            const IgemmProblem& q0 = P.prob[0];
            const int n0 = (int)(blockIdx.x % q0.n_tiles) * P.n_out;
            float* sb0 = s_bias + (grp * 2) * (512 / NG);
            for (int i = etid; i < P.n_out; i += 128)
                sb0[i] = __ldg(q0.bias + (EPI_OF(q0) == EPI_CONVT ? (n0 + i) % q0.convt_cout
                                          : EPI_OF(q0) == EPI_CONVTFIX ? ((n0 / P.BN) * (P.BN / 3) + i % (P.BN / 3)) % q0.convt_cout : n0 + i));
        }
        // 32-channel STORE layers with a static bias (d1.1 -- the largest layer of the forward): the 32 bias values live in
        // registers for the whole kernel instead of eight LDS.128 per tile in front of the first FADD of the conversion chain
        // (ncu source view of d1.1: 13 % of the kernel's samples were FADDs waiting on those loads, short_sb)
        // (only in the instantiation that layer uses -- row-shifted taps, STORE, KC = 32: elsewhere the 32 registers would be
        // reserved for nothing and push the conversion loop into spills)
        constexpr bool BIAS_REGS = (AM == AMODE_RS && EP == EPI_STORE && KKT == 2);
        const bool bias_regs = BIAS_REGS && bias_static && P.BN == 32 && P.CB == 32 && P.n_out == 32;
        float4 bq[BIAS_REGS ? 8 : 1];
        if constexpr (BIAS_REGS) {
#pragma unroll
            for (int k = 0; k < 8; ++k) bq[k] = bias_regs ? __ldg(reinterpret_cast<const float4*>(P.prob[0].bias) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // The barrier at the top of a tile publishes the tile's bias and the "staging tile is free again" news.  With a
        // static bias it can go when nothing is staged (OUTCONV), or when tiles are single-chunk STOREs alternating
        // between two staging tiles: there thread 0 waits for the PREVIOUS tile's store to have read its tile right
        // before the mid-tile barrier every thread passes anyway (that store was issued a whole tile ago).
        const int epi0 = EPI_OF(P.prob[0]);
        const int chunks_per_tile = amode == AMODE_DXN ? 1 : P.MT * P.BN / P.CB;
        const bool early_wait = bias_static && !P.cbatch && P.cslots == 2 && chunks_per_tile == 1 && epi0 == EPI_STORE && P.lean_sync != 0;
        const bool skip_top_bar = bias_static && P.lean_sync != 0 && (epi0 == EPI_OUTCONV || early_wait);
        asm volatile("griddepcontrol.wait;" ::: "memory");          // before the first store / activation read of this role
        // immediate barrier ids: a register operand would make ptxas reserve all 16 hardware barriers per CTA
#define EPI_BAR()                                                           \
    do {                                                                    \
        if (NG == 2) {                          /* two predicated barriers instead of a four-way jump table */ \
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");    \
            else          asm volatile("bar.sync 2, 128;" ::: "memory");    \
        } else if (grp < 2) {                                               \
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");    \
            else          asm volatile("bar.sync 2, 128;" ::: "memory");    \
        } else {                                                            \
            if (grp == 2) asm volatile("bar.sync 3, 128;" ::: "memory");    \
            else          asm volatile("bar.sync 4, 128;" ::: "memory");    \
        }                                                                   \
    } while (0)
        if (skip_top_bar) EPI_BAR();                                 // the static bias becomes visible once
        TileIter it;
        TM_DECL();
        for (it.init(P, blockIdx.x + grp * gridDim.x, NG * gridDim.x); it.valid(); it.next()) {
            const TileCoord tc = it.coord(P);
            const IgemmProblem& q = MULTI ? P.prob[tc.pi] : P.prob[0];   // single-problem launches: fixed parameter offsets
            const int xoff = amode == AMODE_RS ? 1 : 0;           // RS: output column j sits at slab column j + 1
            const int y = tc.y0 + ty, x = tc.x0 + xoff + tx;
            const bool colok = amode != AMODE_RS || tx < P.VW;    // RS: the last two columns of a row are wrap-around garbage
            const bool valid = (y < q.H) && (x < q.W) && colok;
            const int srow = ty * P.VW + tx;                        // row inside the (CB, VW, TH) store box
            // bias of this tile, double-buffered by iteration parity (a slow thread may still read the previous one)
            if (!bias_static) par ^= 1;
            float* sb = s_bias + (grp * 2 + (bias_static ? 0 : par)) * (512 / NG);
            if (!bias_static) {
                const float* bsrc = q.bias + (q.bias_img_stride ? (size_t)tc.b * q.bias_img_stride : 0);
                for (int i = etid; i < P.n_out; i += 128)
                    sb[i] = __ldg(bsrc + (EPI_OF(q) == EPI_CONVT ? (tc.n0 + i) % q.convt_cout
                                          : EPI_OF(q) == EPI_CONVTFIX ? ((tc.n0 / P.BN) * (P.BN / 3) + i % (P.BN / 3)) % q.convt_cout : tc.n0 + i));
            }
            TM_MARK(0);                                         // 0: tile bookkeeping
            if (etid == 0 && !skip_top_bar) WAIT_STORE_READS(); // the store that last used the next staging tile has left smem
            if (EPI_OF(q) == EPI_GATE && etid == 0) {
                // attention gate: the first chunk of the skip tile that psi will scale does not depend on the MMAs -- fetch it
                // now, under the wait for the accumulator, instead of after the sigmoid
                ptx::mbar_expect_tx(cbar, (uint32_t)P.c_slot_bytes);
                ptx::tma_load_4d(cs, &P.tmC[0], cbar, 0, tc.x0, tc.y0, tc.b);
            }
            TM_MARK(1);                                         // 1: wait for the previous TMA store to have read its tile
            ptx::mbar_wait(full_bar, acc_phase, P.err, ERR_EPI_WAIT);
            acc_phase ^= 1;
            ptx::tc_fence_after();
            TM_MARK(2);                                         // 2: wait for the accumulator
            if (!skip_top_bar) EPI_BAR();                       // bias visible, staging tiles free
            TM_MARK(3);                                         // 3: group barrier at the top
            const uint32_t taddr0 = tmem_base + (uint32_t)(acc * P.BN * P.MT) + ((uint32_t)(quarter * 32) << 16);
            uint32_t taddr = taddr0;

            if (amode == AMODE_DXN) {
                // ---- combine the three dx column groups: out[p] = E0[p-1] + E1[p] + E2[p+1]  (p = lane = slab column)
                const int N = P.n_out;
                const bool inner = lane >= 1 && lane <= P.VW;                    // the 30 valid output columns
                float dot = 0.f;
                const int srow = quarter * P.VW + lane - 1;                      // DXN: output column = slab column - 1
                for (int c0 = 0; c0 < N; c0 += 16) {
                    uint32_t e0[32], e1[32], e2[32];
                    ptx::tmem_ld_32x16(taddr + c0, e0);
                    ptx::tmem_ld_32x16(taddr + N + c0, e1);
                    ptx::tmem_ld_32x16(taddr + 2 * N + c0, e2);
                    ptx::tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + v * 4);   // (hoisting these above the wait measured +1.7 % on u1.conv.1)
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = v * 4 + u;
                            const float l = __shfl_up_sync(0xffffffffu, __uint_as_float(e0[i]), 1);
                            const float r = __shfl_down_sync(0xffffffffu, __uint_as_float(e2[i]), 1);
                            f[i] = fmaxf((l + __uint_as_float(e1[i])) + (r + bb[u]), 0.f);   // every DXN layer ends in ReLU
                        }
                    }
                    if (EPI_OF(q) == EPI_STORE) {
                        if (inner) {
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                const uint4 o = make_uint4(pack2(f[v * 8 + 0], f[v * 8 + 1], f16), pack2(f[v * 8 + 2], f[v * 8 + 3], f16),
                                                           pack2(f[v * 8 + 4], f[v * 8 + 5], f16), pack2(f[v * 8 + 6], f[v * 8 + 7], f16));
                                uint32_t off = (uint32_t)(srow * c_pitch + (c0 + v * 8) * 2);
                                off ^= ((off >> 7) & swz_mask) << 4;
                                sts128(cs + off, o);
                            }
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 w4 = *reinterpret_cast<const float4*>(s_vec + c0 + v * 4);
                            dot = fmaf(f[v * 4 + 0], w4.x, dot);
                            dot = fmaf(f[v * 4 + 1], w4.y, dot);
                            dot = fmaf(f[v * 4 + 2], w4.z, dot);
                            dot = fmaf(f[v * 4 + 3], w4.w, dot);
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) ptx::mbar_arrive_leader(empty_bar); else ptx::mbar_arrive(empty_bar); }
                if (EPI_OF(q) == EPI_STORE) {
                    if (early_wait && etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    EPI_BAR();
                    if (fused_pool) {
                        pool_staged_tile(cs, ps, P.TH, P.VW, c_pitch, swz_mask, etid, f16);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        EPI_BAR();
                    }
                    if (etid == 0) {
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"((uint64_t)&P.tmC[0]), "r"(cs), "r"(tc.n0), "r"(tc.x0 + 1), "r"(tc.y0), "r"(tc.b) : "memory");
                        if (fused_pool)
                            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                         ::"l"((uint64_t)&P.tmP), "r"(ps), "r"(tc.n0), "r"((tc.x0 + 1) >> 1), "r"(tc.y0 >> 1), "r"(tc.b) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    NEXT_CSLOT();
                } else if (inner && valid) {
                    q.aux[((size_t)tc.b * q.H + y) * q.W + x] = dot + q.scalar;
                }
            } else if (EPI_OF(q) == EPI_STORE || EPI_OF(q) == EPI_CONVT) {
              // One staging tile per CB-channel chunk.  cbatch: the tile's chunks all have their own staging tile, so the
              // whole accumulator is converted first and fence / barrier / TMA stores happen ONCE per tile; otherwise the
              // chunks rotate through `cslots` tiles with a barrier pair per chunk.
              const bool batch = P.cbatch != 0;
              auto issue_store = [&](uint32_t c_tile, uint32_t p_tile, int mb, int c0) {
                  const int n = tc.n0 + c0, yb = tc.y0 + mb * P.TH;
                  const void* tm;
                  int cch;
                  if (EPI_OF(q) == EPI_STORE) { tm = &P.tmC[tc.pi]; cch = n; }
                  else { const int ab = n / q.convt_cout; tm = &P.tmC[ab]; cch = n - ab * q.convt_cout; }
                  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                               ::"l"((uint64_t)tm), "r"(c_tile), "r"(cch), "r"(tc.x0 + xoff), "r"(yb), "r"(tc.b) : "memory");
                  if (fused_pool)
                      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                   ::"l"((uint64_t)&P.tmP), "r"(p_tile), "r"(n), "r"((tc.x0 + xoff) >> 1), "r"(yb >> 1), "r"(tc.b) : "memory");
              };
              for (int mb = 0; mb < P.MT; ++mb) {                               // M-blocks of the tile (vertically stacked)
                taddr = taddr0 + (uint32_t)(mb * P.BN);
                for (int c0 = 0; c0 < P.BN; c0 += P.CB) {
                    if (!batch && (c0 > 0 || mb > 0)) {                         // next staging tile: the store that last used it must be out
                        if (etid == 0) WAIT_STORE_READS();
                        EPI_BAR();
                    }
                    {
                        // staged row of this thread: row_base + (column byte ^ xr); the TMA swizzle XORs address bits
                        // 4..6 with bits 7..9, and a staged row never crosses a 128-byte line, so xr is per thread
                        const uint32_t row_base = (uint32_t)(srow * c_pitch);
                        const uint32_t xr = ((row_base >> 7) & swz_mask) << 4;
                        const uint32_t cs_row = cs + row_base;
                        const bool relu = q.relu != 0;
                        auto convert8b = [&](const uint32_t* r8, int col, const float4& b0, const float4& b1) {   // 8 accumulator columns -> one 16-byte vector
                            const float f0 = __uint_as_float(r8[0]) + b0.x, f1 = __uint_as_float(r8[1]) + b0.y;
                            const float f2 = __uint_as_float(r8[2]) + b0.z, f3 = __uint_as_float(r8[3]) + b0.w;
                            const float f4 = __uint_as_float(r8[4]) + b1.x, f5 = __uint_as_float(r8[5]) + b1.y;
                            const float f6 = __uint_as_float(r8[6]) + b1.z, f7 = __uint_as_float(r8[7]) + b1.w;
                            const uint4 o = relu ? make_uint4(pack2_relu<F16>(f0, f1), pack2_relu<F16>(f2, f3), pack2_relu<F16>(f4, f5), pack2_relu<F16>(f6, f7))
                                                 : make_uint4(pack2(f0, f1, f16), pack2(f2, f3, f16), pack2(f4, f5, f16), pack2(f6, f7, f16));
                            if (colok) sts128(cs_row + ((uint32_t)(col * 2) ^ xr), o);
                        };
                        auto convert8 = [&](const uint32_t* r8, int col) {
                            convert8b(r8, col, *reinterpret_cast<const float4*>(sb + c0 + col), *reinterpret_cast<const float4*>(sb + c0 + col + 4));
                        };
                        bool done = false;
                        if constexpr (BIAS_REGS) {
                            if (bias_regs) {                                         // BN == CB == 32: one pass, bias from registers
                                uint32_t ra[32], rb[32];
                                ptx::tmem_ld_32x16(taddr, ra);
                                ptx::tmem_ld_wait();
                                ptx::tmem_ld_32x16(taddr + 16, rb);
                                convert8b(ra, 0, bq[0], bq[1]);
                                convert8b(ra + 8, 8, bq[2], bq[3]);
                                ptx::tmem_ld_wait();
                                convert8b(rb, 16, bq[4], bq[5]);
                                convert8b(rb + 8, 24, bq[6], bq[7]);
                                done = true;
                            }
                        }
                        if (done) {
                        } else if ((P.CB & 31) == 0) {
                            // 16 accumulator columns at a time, the next 16 in flight while these are converted (one x32
                            // load per 32 columns left the whole TMEM read latency in front of the first FADD: 21 % of
                            // d1.1's epilogue samples)
                            // (the bias vectors of each 16-column step are read BEFORE the wait on that step's accumulator load,
                            // so their shared-memory latency hides under it instead of stalling the first FADDs)
                            uint32_t ra[32], rb[32];
                            ptx::tmem_ld_32x16(taddr + c0, ra);
                            for (int cc = 0; cc < P.CB; cc += 32) {
                                const float4* bp = reinterpret_cast<const float4*>(sb + c0 + cc);
                                const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
                                ptx::tmem_ld_wait();
                                ptx::tmem_ld_32x16(taddr + c0 + cc + 16, rb);
                                const float4 b4 = bp[4], b5 = bp[5], b6 = bp[6], b7 = bp[7];
                                convert8b(ra, cc, b0, b1);
                                convert8b(ra + 8, cc + 8, b2, b3);
                                ptx::tmem_ld_wait();
                                if (cc + 32 < P.CB) ptx::tmem_ld_32x16(taddr + c0 + cc + 32, ra);
                                convert8b(rb, cc + 16, b4, b5);
                                convert8b(rb + 8, cc + 24, b6, b7);
                            }
                        } else {                                                     // CB == 16
                            uint32_t r[32];
                            ptx::tmem_ld_32x16(taddr + c0, r);
                            ptx::tmem_ld_wait();
                            convert8(r, 0);
                            convert8(r + 8, 8);
                        }
                    }
                    if (c0 + P.CB >= P.BN && mb == P.MT - 1) {                // accumulator fully read: free the TMEM stage
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR) ptx::mbar_arrive_leader(empty_bar); else ptx::mbar_arrive(empty_bar); }
                    }
                    TM_MARK(4);                                                 // 4: TMEM -> registers -> staged tile
                    if (!batch) {
                        if (early_wait && etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        EPI_BAR();
                        TM_MARK(5);                                             // 5: proxy fence + barrier
                        if (fused_pool) {
                            pool_staged_tile(cs, ps, P.TH, P.VW, c_pitch, swz_mask, etid, f16);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            EPI_BAR();
                            TM_MARK(6);                                         // 6: pooling + fence + barrier
                        }
                        if (etid == 0) {
                            issue_store(cs, ps, mb, c0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        TM_MARK(7);                                             // 7: TMA store issue
                    }
                    NEXT_CSLOT();
                }
              }
              if (batch) {                                                      // cslot is back at 0: chunk j sits in staging tile j
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  EPI_BAR();
                  if (fused_pool) {
                      for (int j = 0; j < P.cslots; ++j)
                          pool_staged_tile(c_grp + (uint32_t)(j * P.c_slot_bytes), p_grp + (uint32_t)(j * P.p_slot_bytes), P.TH, P.VW, c_pitch, swz_mask, etid, f16);
                      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                      EPI_BAR();
                  }
                  if (etid == 0) {
                      int j = 0;
                      for (int mb = 0; mb < P.MT; ++mb)
                          for (int c0 = 0; c0 < P.BN; c0 += P.CB, ++j)
                              issue_store(c_grp + (uint32_t)(j * P.c_slot_bytes), p_grp + (uint32_t)(j * P.p_slot_bytes), mb, c0);
                      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                  }
              }
            } else if (EPI_OF(q) == EPI_CONVTFIX) {
                // n tile = (phase t along the free axis, channel block); columns [up | mid0 | mid1], CC channels each
                const int CC = P.BN / 3;
                const int nt = tc.n0 / P.BN, nsub = q.convt_cout / CC;
                const int t = nt / nsub, ch0 = (nt - t * nsub) * CC;
                const int j = q.fix_axis == 0 ? y : x;                          // index along the fixed axis
                float cf[2][3];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int o = 2 * j + p;
#pragma unroll
                    for (int k = 0; k < 3; ++k) cf[p][k] = o < q.fix_out ? __ldg(q.fix_coef + o * 3 + k) : 0.f;
                }
                const uint32_t row_base = (uint32_t)(srow * c_pitch);
                const uint32_t xr = ((row_base >> 7) & swz_mask) << 4;
                const uint32_t slot0 = c_grp;
                for (int c0 = 0; c0 < CC; c0 += P.CB) {
                    if (c0 > 0) {                                               // both staging tiles are reused per chunk
                        if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        EPI_BAR();
                    }
                    for (int cc = 0; cc < P.CB; cc += 16) {
                        uint32_t u[32], m0[32], m1[32];
                        ptx::tmem_ld_32x16(taddr + c0 + cc, u);
                        ptx::tmem_ld_32x16(taddr + CC + c0 + cc, m0);
                        ptx::tmem_ld_32x16(taddr + 2 * CC + c0 + cc, m1);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                float f[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int c = v * 8 + i;
                                    f[i] = fmaf(cf[p][0], __uint_as_float(u[c]), fmaf(cf[p][1], __uint_as_float(m0[c]),
                                                fmaf(cf[p][2], __uint_as_float(m1[c]), sb[c0 + cc + c])));
                                }
                                const uint4 o4 = make_uint4(pack2(f[0], f[1], f16), pack2(f[2], f[3], f16), pack2(f[4], f[5], f16), pack2(f[6], f[7], f16));
                                sts128(slot0 + (uint32_t)(p * P.c_slot_bytes) + row_base + ((uint32_t)((cc + v * 8) * 2) ^ xr), o4);
                            }
                        }
                    }
                    if (c0 + P.CB >= CC) {                                      // accumulator fully read: free the TMEM stage
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR) ptx::mbar_arrive_leader(empty_bar); else ptx::mbar_arrive(empty_bar); }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    EPI_BAR();
                    if (etid == 0) {
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            const int ab = q.fix_axis == 0 ? p * 2 + t : t * 2 + p;   // (a, b) phase of the output view
                            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                         ::"l"((uint64_t)&P.tmC[ab]), "r"(slot0 + (uint32_t)(p * P.c_slot_bytes)), "r"(ch0 + c0), "r"(tc.x0), "r"(tc.y0), "r"(tc.b) : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else {
                // GATE / OUTCONV: dot = sum_n relu(acc_n + bias_n) * vec_n over all BN channels of the pixel
                float dot = 0.f;
                for (int c0 = 0; c0 < P.BN; c0 += 16) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x16(taddr + c0, r);
                    float4 bb[4], ww[4];                                        // read under the accumulator load, not after it
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        bb[v] = *reinterpret_cast<const float4*>(sb + c0 + v * 4);
                        ww[v] = *reinterpret_cast<const float4*>(s_vec + c0 + v * 4);
                    }
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 b = bb[v], w = ww[v];
                        dot = fmaf(fmaxf(__uint_as_float(r[v * 4 + 0]) + b.x, 0.f), w.x, dot);
                        dot = fmaf(fmaxf(__uint_as_float(r[v * 4 + 1]) + b.y, 0.f), w.y, dot);
                        dot = fmaf(fmaxf(__uint_as_float(r[v * 4 + 2]) + b.z, 0.f), w.z, dot);
                        dot = fmaf(fmaxf(__uint_as_float(r[v * 4 + 3]) + b.w, 0.f), w.w, dot);
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) ptx::mbar_arrive_leader(empty_bar); else ptx::mbar_arrive(empty_bar); }
                if (EPI_OF(q) == EPI_OUTCONV) {
                    if (valid) q.aux[((size_t)tc.b * q.H + y) * q.W + x] = dot + q.scalar;
                } else {
                    const float a = 1.f / (1.f + expf(-(dot + q.scalar)));
                    const float scale = q.gate_plus_x ? (1.f + a) : a;
                    if (valid && q.aux) q.aux[((size_t)tc.b * q.H + y) * q.W + x] = a;
                    // scale the skip tile in place: TMA load (L2 hit: the same bytes were just streamed in as the A
                    // operand) -> multiply this thread's pixel row in shared memory -> TMA store, CB channels at a time
                    for (int c0 = 0; c0 < q.gate_C; c0 += P.CB) {
                        if (etid == 0 && c0 > 0) {                          // (chunk 0 was requested at the top of the tile)
                            WAIT_STORE_READS();
                            ptx::mbar_expect_tx(cbar, (uint32_t)P.c_slot_bytes);
                            ptx::tma_load_4d(cs, &P.tmC[0], cbar, c0, tc.x0, tc.y0, tc.b);
                        }
                        ptx::mbar_wait(cbar, c_phase, P.err, ERR_EPI_WAIT);
                        c_phase ^= 1;
                        for (int v = 0; v < (c_pitch >> 4); ++v) {
                            uint32_t off = (uint32_t)(row * c_pitch + v * 16);
                            off ^= ((off >> 7) & swz_mask) << 4;
                            uint4 val = lds128(cs + off);
                            const float2 p0 = unpack2(val.x, f16), p1 = unpack2(val.y, f16), p2 = unpack2(val.z, f16), p3 = unpack2(val.w, f16);
                            val.x = pack2(p0.x * scale, p0.y * scale, f16);
                            val.y = pack2(p1.x * scale, p1.y * scale, f16);
                            val.z = pack2(p2.x * scale, p2.y * scale, f16);
                            val.w = pack2(p3.x * scale, p3.y * scale, f16);
                            sts128(cs + off, val);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        EPI_BAR();
                        if (etid == 0) {
                            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                         ::"l"((uint64_t)&P.tmC[0]), "r"(cs), "r"(c0), "r"(tc.x0), "r"(tc.y0), "r"(tc.b) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        NEXT_CSLOT();
                    }
                }
            }
        }
        TM_FLUSH(8 + 8 * (grp & 1), lane == 0 && quarter == 0);
#undef EPI_BAR
#undef EPI_OF
#undef NEXT_CSLOT
#undef WAIT_STORE_READS
        if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all TMA stores landed
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR) ptx::cluster_sync_all();                      // both CTAs are done with each other's barriers and TMEM
    if (warp == 1) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_pair(tmem_base, (uint32_t)P.tmem_cols);
        else      ptx::tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

}  // namespace aau
