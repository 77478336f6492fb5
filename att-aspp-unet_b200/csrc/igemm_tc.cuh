// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands staged in
// shared memory by TMA).  One persistent kernel serves every GEMM-shaped layer of AttentionASPPUNet:
//
//   rows    (M) : 128 output pixels = a TH x TW patch of one frame (NHWC activations, TH*TW == 128)
//   columns (N) : BN output channels (<= 256)
//   depth   (K) : taps x Cin, walked in sub-blocks of KC channels (KC*2 bytes == the TMA/UMMA swizzle width)
//
// A operand, two staging modes:
//   AMODE_TAP  : one TMA box (KC, TW, TH) per (tap, channel chunk), shifted by the tap offset * dilation; image
//                borders and dilation overhang come back as zeros from the TMA out-of-bounds fill.  Works for
//                1x1, 3x3 and dilated 3x3.
//   AMODE_SLAB : 3x3 / dilation 1 only.  One TMA box (KC, TW, TH+2) per (dx, channel chunk) -- a column-shifted
//                slab with a one-row halo above and below.  The three vertical taps are then three MMAs whose A
//                descriptors start TW rows apart inside the same slab (TW % 8 == 0 keeps them on the swizzle
//                period), so every activation byte is fetched from L2 3.75x instead of 9x.
// B operand: weights [N][K] (K contiguous), one TMA box (KC, BN) per sub-block.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (each owns the TMEM lane quarter warp_id % 4).  Two accumulator stages in TMEM let the
// epilogue of tile i overlap the MMAs of tile i+1.
//
// Epilogues (fp32 math on the accumulator, single rounding to the 16-bit activation type):
//   EPI_STORE   : + bias (per channel or per image), optional ReLU, NHWC store at a channel offset / pixel stride
//                 (this is how ASPP branches and skip tensors land directly inside concatenated buffers)
//   EPI_CONVT   : ConvTranspose2d(2,2): column n = (a*2+b)*Cout + co is scattered to pixel (2y+a, 2x+b)
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

namespace aau {

enum { EPI_STORE = 0, EPI_CONVT = 1, EPI_GATE = 2, EPI_OUTCONV = 3 };
enum { AMODE_TAP = 0, AMODE_SLAB = 1 };
enum { IGEMM_THREADS = 192, IGEMM_MAX_PROBLEMS = 4, IGEMM_MAX_STAGES = 8 };
enum { ERR_PRODUCER_WAIT = 101, ERR_MMA_WAIT_FULL = 102, ERR_MMA_WAIT_TMEM = 103, ERR_EPI_WAIT = 104 };

struct alignas(64) IgemmProblem {
    CUtensorMap tmA;        // activations, 4-D (C, W, H, B)
    CUtensorMap tmB;        // weights, 2-D (K, N)
    const float* bias;      // [N] or [B][bias_img_stride]
    void* out;              // STORE / CONVT: NHWC destination; GATE: skip tensor, scaled in place
    const float* vec;       // GATE: w_psi[BN]; OUTCONV: w_out[BN]
    float* aux;             // GATE: psi map (may be null); OUTCONV: logits
    int H, W;               // pixel grid of the GEMM rows
    int tiles_x, tiles_per_img, m_tiles, n_tiles, tile_begin;
    int taps, dil, nchunk;  // taps in {1, 9}; nchunk = Cin / KC
    int epi, relu, bias_img_stride;
    int outH, outW, out_ld, out_choff;
    int convt_cout;
    int gate_C, gate_plus_x;
    float scalar;           // GATE: b_psi; OUTCONV: b_out
    int pad_;
};

struct alignas(64) IgemmParams {
    IgemmProblem prob[IGEMM_MAX_PROBLEMS];
    int nprob, total_tiles;
    int amode;              // AMODE_*
    int KC, G, nstages;     // channels per sub-block, sub-blocks per stage (TAP), smem ring depth
    int TW, TH, tw_shift;
    int BN;
    int a_stage_bytes, b_sub_bytes, stage_bytes;
    int tmem_cols;
    int is_fp16;
    int* err;
};

// ---- 16-bit activation helpers (bf16 by default, fp16 as the higher-precision storage option) -------------
__device__ __forceinline__ uint32_t pack2(float a, float b, int is_fp16) {
    if (is_fp16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t v, int is_fp16) {
    if (is_fp16) return __half22float2(*reinterpret_cast<__half2*>(&v));
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

struct TileCoord { int pi, b, y0, x0, n0; };

__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& P, int t) {
    TileCoord tc;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < IGEMM_MAX_PROBLEMS; ++i)
        if (i < P.nprob && t >= P.prob[i].tile_begin) pi = i;
    const IgemmProblem& q = P.prob[pi];
    const int local = t - q.tile_begin;
    const int nt = local % q.n_tiles;
    const int mt = local / q.n_tiles;
    tc.pi = pi;
    tc.b = mt / q.tiles_per_img;
    const int r = mt - tc.b * q.tiles_per_img;
    const int tyi = r / q.tiles_x;
    tc.y0 = tyi * P.TH;
    tc.x0 = (r - tyi * q.tiles_x) * P.TW;
    tc.n0 = nt * P.BN;
    return tc;
}

__global__ void __launch_bounds__(IGEMM_THREADS, 1) igemm_tc_kernel(const __grid_constant__ IgemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[IGEMM_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[IGEMM_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // operand tiles need 1024-byte alignment for the 128-byte swizzle
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.nprob; ++i) {
            ptx::prefetch_tmap(&P.prob[i].tmA);
            ptx::prefetch_tmap(&P.prob[i].tmB);
        }
        for (int s = 0; s < P.nstages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full_bar[a], 1);
            ptx::mbar_init(&tmem_empty_bar[a], 4);     // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_smem, (uint32_t)P.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int swz_bytes = P.KC * 2;
    const int kk_per_sub = P.KC >> 4;                  // UMMA K = 16 elements = 32 bytes

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(P, t);
                const IgemmProblem& q = P.prob[tc.pi];
                if (P.amode == AMODE_TAP) {
                    const int sub_total = q.taps * q.nchunk;
                    for (int s0 = 0; s0 < sub_total; s0 += P.G) {
                        const int nsub = min(P.G, sub_total - s0);
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1, P.err, ERR_PRODUCER_WAIT);
                        uint8_t* sa = smem + (size_t)stage * P.stage_bytes;
                        uint8_t* sb = sa + P.a_stage_bytes;
                        const int a_sub_bytes = 128 * swz_bytes;
                        ptx::mbar_expect_tx(&full_bar[stage], (uint32_t)(nsub * (a_sub_bytes + P.b_sub_bytes)));
                        for (int j = 0; j < nsub; ++j) {
                            const int sub = s0 + j;
                            const int tap = sub / q.nchunk;
                            const int ch = sub - tap * q.nchunk;
                            int dy = 0, dx = 0;
                            if (q.taps == 9) { dy = (tap / 3 - 1) * q.dil; dx = (tap % 3 - 1) * q.dil; }
                            ptx::tma_load_4d(sa + j * a_sub_bytes, &q.tmA, &full_bar[stage], ch * P.KC, tc.x0 + dx, tc.y0 + dy, tc.b);
                            ptx::tma_load_2d(sb + j * P.b_sub_bytes, &q.tmB, &full_bar[stage], sub * P.KC, tc.n0);
                        }
                        if (++stage == P.nstages) { stage = 0; phase ^= 1; }
                    }
                } else {
                    const int cin = q.nchunk * P.KC;
                    const int slab_bytes = (P.TH + 2) * P.TW * swz_bytes;
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        for (int ch = 0; ch < q.nchunk; ++ch) {
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1, P.err, ERR_PRODUCER_WAIT);
                            uint8_t* sa = smem + (size_t)stage * P.stage_bytes;
                            uint8_t* sb = sa + P.a_stage_bytes;
                            ptx::mbar_expect_tx(&full_bar[stage], (uint32_t)(slab_bytes + 3 * P.b_sub_bytes));
                            ptx::tma_load_4d(sa, &q.tmA, &full_bar[stage], ch * P.KC, tc.x0 + dxi - 1, tc.y0 - 1, tc.b);
                            for (int dyi = 0; dyi < 3; ++dyi)
                                ptx::tma_load_2d(sb + dyi * P.b_sub_bytes, &q.tmB, &full_bar[stage],
                                                 (dyi * 3 + dxi) * cin + ch * P.KC, tc.n0);
                            if (++stage == P.nstages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_f16(128, P.BN, P.is_fp16 != 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(P, t);
                const IgemmProblem& q = P.prob[tc.pi];
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, P.err, ERR_MMA_WAIT_TMEM);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P.BN);
                uint32_t accumulate = 0;
                int nstage_k, nsub_full, sub_total;
                if (P.amode == AMODE_TAP) {
                    sub_total = q.taps * q.nchunk;
                    nstage_k = (sub_total + P.G - 1) / P.G;
                    nsub_full = P.G;
                } else {
                    sub_total = 9 * q.nchunk;
                    nstage_k = 3 * q.nchunk;
                    nsub_full = 3;
                }
                const int a_step = (P.amode == AMODE_TAP) ? 128 * swz_bytes : P.TW * swz_bytes;
                for (int s = 0; s < nstage_k; ++s) {
                    const int nsub = min(nsub_full, sub_total - s * nsub_full);
                    ptx::mbar_wait(&full_bar[stage], phase, P.err, ERR_MMA_WAIT_FULL);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * P.stage_bytes);
                    const uint32_t sb = sa + (uint32_t)P.a_stage_bytes;
                    for (int j = 0; j < nsub; ++j) {
                        for (int k = 0; k < kk_per_sub; ++k) {
                            const uint64_t da = ptx::make_kmajor_desc(sa + j * a_step + k * 32, swz_bytes);
                            const uint64_t db = ptx::make_kmajor_desc(sb + j * P.b_sub_bytes + k * 32, swz_bytes);
                            ptx::umma_f16(d_tmem, da, db, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    ptx::umma_commit(&empty_bar[stage]);        // frees the smem slot once these MMAs retire
                    if (++stage == P.nstages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full_bar[acc]);          // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // =========================== epilogue (warps 2..5) ===========================
        const int quarter = warp & 3;                           // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;
        const int ty = row >> P.tw_shift;
        const int tx = row & (P.TW - 1);
        const int f16 = P.is_fp16;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(P, t);
            const IgemmProblem& q = P.prob[tc.pi];
            const int y = tc.y0 + ty, x = tc.x0 + tx;
            const bool valid = (y < q.H) && (x < q.W);
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase, P.err, ERR_EPI_WAIT);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(acc * P.BN) + ((uint32_t)(quarter * 32) << 16);
            const float* bias = q.bias + (q.bias_img_stride ? (size_t)tc.b * q.bias_img_stride : 0) + tc.n0;
            float dot = 0.f;                                    // GATE / OUTCONV reduction over channels
            uint8_t* out_row = nullptr;
            if (q.epi == EPI_STORE)
                out_row = (uint8_t*)q.out + ((((size_t)tc.b * q.outH + y) * q.outW + x) * q.out_ld + q.out_choff + tc.n0) * 2;

            for (int c0 = 0; c0 < P.BN; c0 += 32) {
                uint32_t r[32];
                const int ncol = min(32, P.BN - c0);
                if (ncol == 32) ptx::tmem_ld_32x32(taddr + c0, r);
                else            ptx::tmem_ld_32x16(taddr + c0, r);
                ptx::tmem_ld_wait();
                if (q.epi == EPI_STORE || q.epi == EPI_CONVT) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (v * 8 < ncol) {
                            float f[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[v * 8 + i]);
                            if (q.epi == EPI_STORE) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    f[i] += __ldg(bias + c0 + v * 8 + i);
                                    if (q.relu) f[i] = fmaxf(f[i], 0.f);
                                }
                                if (valid) {
                                    uint4 o = make_uint4(pack2(f[0], f[1], f16), pack2(f[2], f[3], f16), pack2(f[4], f[5], f16), pack2(f[6], f[7], f16));
                                    *reinterpret_cast<uint4*>(out_row + (c0 + v * 8) * 2) = o;
                                }
                            } else {
                                const int n = tc.n0 + c0 + v * 8;
                                const int ab = n / q.convt_cout;
                                const int co = n - ab * q.convt_cout;
                                const int oy = 2 * y + (ab >> 1), ox = 2 * x + (ab & 1);
#pragma unroll
                                for (int i = 0; i < 8; ++i) f[i] += __ldg(q.bias + co + i);
                                if (valid && oy < q.outH && ox < q.outW) {
                                    uint4 o = make_uint4(pack2(f[0], f[1], f16), pack2(f[2], f[3], f16), pack2(f[4], f[5], f16), pack2(f[6], f[7], f16));
                                    uint8_t* dst = (uint8_t*)q.out + ((((size_t)tc.b * q.outH + oy) * q.outW + ox) * q.out_ld + q.out_choff + co) * 2;
                                    *reinterpret_cast<uint4*>(dst) = o;
                                }
                            }
                        }
                    }
                } else {
                    // GATE / OUTCONV: dot += sum_n relu(acc_n + bias_n) * vec_n
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i < ncol) {
                            float f = __uint_as_float(r[i]) + __ldg(bias + c0 + i);
                            f = fmaxf(f, 0.f);
                            dot = fmaf(f, __ldg(q.vec + tc.n0 + c0 + i), dot);
                        }
                    }
                }
            }
            // the accumulator has been read into registers: hand the TMEM stage back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);

            if (q.epi == EPI_OUTCONV) {
                if (valid) q.aux[((size_t)tc.b * q.H + y) * q.W + x] = dot + q.scalar;
            } else if (q.epi == EPI_GATE) {
                const float a = 1.f / (1.f + expf(-(dot + q.scalar)));
                const float scale = q.gate_plus_x ? (1.f + a) : a;
                if (valid) {
                    if (q.aux) q.aux[((size_t)tc.b * q.H + y) * q.W + x] = a;
                    uint4* xr = reinterpret_cast<uint4*>((uint8_t*)q.out + ((((size_t)tc.b * q.outH + y) * q.outW + x) * q.out_ld + q.out_choff) * 2);
                    for (int j = 0; j < (q.gate_C >> 3); ++j) {
                        uint4 v = xr[j];
                        float2 p0 = unpack2(v.x, f16), p1 = unpack2(v.y, f16), p2 = unpack2(v.z, f16), p3 = unpack2(v.w, f16);
                        v.x = pack2(p0.x * scale, p0.y * scale, f16);
                        v.y = pack2(p1.x * scale, p1.y * scale, f16);
                        v.z = pack2(p2.x * scale, p2.y * scale, f16);
                        v.w = pack2(p3.x * scale, p3.y * scale, f16);
                        xr[j] = v;
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

}  // namespace aau
