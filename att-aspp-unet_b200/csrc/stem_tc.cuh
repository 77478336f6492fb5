// d1.0 on the tensor cores: Conv2d(1 -> C, 3x3, pad 1) + BN + ReLU of a uint8 sweep as an implicit GEMM whose K = 9
// taps are padded to one K = 16 tcgen05.mma per 128 pixels (reference: attention_aspp_unet_pipeline_stage.py:24-33,
// d1 = DoubleConv(1, c); the frame is divided by 255 before it -- here 1/255 is folded into the weights, so the A
// operand holds the raw pixel values 0..255, which bf16 / fp16 represent exactly; the weights go in as two 16-bit terms,
// hi + lo, one MMA each into the same accumulator, so the layer computes with ~22-bit weights: as the FIRST layer its
// weight rounding was the largest single source of logit error of the whole 16-bit network).
//
// Why not FMAs: 288 FMAs per pixel put the packed-fp32 stem at 2.9 TB/s of output (profiles/r01_ncu_v10_stem.txt, FMA
// pipe bound); on the tensor pipe the same work is one 40-cycle MMA per 128 pixels and the layer is bounded by writing
// its 2*C bytes per pixel.  What is left for the CUDA cores is the im2col itself: 9 byte loads, 9 conversions and two
// 16-byte shared stores per pixel.
//
// Pixels are taken in linear order over the whole batch (p = (b*H + y)*W + x), 512 per macro-tile, so there is no
// tile padding at row / image ends; taps outside the image are masked to zero per pixel.  Roles of a CTA (448 threads,
// two CTAs per SM): warp 0 = TMEM owner + MMA issuer, warps 1-4 = im2col builders (one pixel row of the A tile per
// thread and 128-pixel sub-tile), warps 5-12 = two epilogue groups (group g drains accumulator stage g: tcgen05.ld ->
// bias, ReLU, 16-bit pack -> swizzled staging -> one TMA store per sub-tile), warp 13 = input producer: the three
// 514-byte row segments a macro-tile reads (rows y-1, y, y+1 in linear pixel order) arrive in a four-slot shared-memory
// ring by cp.async.bulk, several macro-tiles ahead -- read straight from global memory the builders spent 57 % of their
// time in load latency (ncu source view, profiles/r01_ncu_stem_tc.txt).  Macro-tiles whose segments would start
// before the buffer or end after it (the first and last few) keep the masked global loads.
#pragma once
#include "igemm_tc.cuh"

namespace aau {

constexpr int STEM_TC_THREADS = 32 + 128 + 256 + 32;
constexpr int STEM_IN_SLOTS = 4, STEM_SEG_BYTES = 544, STEM_IN_SLOT_BYTES = 3 * STEM_SEG_BYTES;   // 512 + 2 pixels + 16-byte alignment slack
constexpr int STEM_TC_SUB = 4;                       // 128-pixel sub-tiles (MMAs) per macro-tile
enum { ERR_STEM_BUILD_WAIT = 111, ERR_STEM_MMA_WAIT = 112, ERR_STEM_EPI_WAIT = 113 };

struct StemTcParams {
    CUtensorMap tmC;        // (C, P) 16-bit output, box (CB, 128), swizzle CB*2 bytes
    const uint8_t* x;       // [P] pixels, frames contiguous
    const uint16_t* wB;     // [2][C][16] K-major weights w*s/255 as hi + lo terms in the activation type (taps 0..8, then zeros)
    const float* bias;      // [C]
    int* err;
    uint32_t P;             // B*H*W
    int H, W, C, CB;
    FastDiv fdW, fdH;
    int n_macro;            // ceil(P / 512)
    int tmem_cols;          // power of two >= 2 * STEM_TC_SUB * C
    int is_fp16;
    int x_aligned;          // x is 16-byte aligned (cp.async.bulk source)
};

static inline size_t stem_tc_smem_bytes(int C) {
    return 1024 /*align*/ + 2 * STEM_TC_SUB * 4096 /*A*/ + 4096 /*B hi, lo*/ + 2 * STEM_TC_SUB * 128 * C * 2 /*staging*/ + STEM_IN_SLOTS * STEM_IN_SLOT_BYTES /*input ring*/;
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <bool F16>
__global__ void __launch_bounds__(STEM_TC_THREADS, 2) stem_tc_kernel(const __grid_constant__ StemTcParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2], in_full[STEM_IN_SLOTS], in_empty[STEM_IN_SLOTS];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[64];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem0;                                      // [stage][sub][128 rows][32 B], 32-byte swizzle
    const uint32_t smem_b = smem_a + 2 * STEM_TC_SUB * 4096;            // [C rows][32 B], same swizzle
    const uint32_t smem_c = smem_b + 4096;                              // [group][sub][chunk][128 rows][CB*2 B]
    const int C = P.C;
    const uint32_t smem_in = smem_c + (uint32_t)(2 * STEM_TC_SUB * 128 * C * 2);   // [slot][row -1, 0, +1][544 B]
    // a macro-tile is "fast" when its three row segments, widened to 16-byte alignment, lie inside the pixel buffer
    auto fast_tile = [&](int m) -> bool {
        const uint32_t p0 = (uint32_t)m * 512u;
        return P.x_aligned != 0 && p0 >= (uint32_t)P.W + 16u && (unsigned long long)p0 + (unsigned)P.W + STEM_SEG_BYTES <= (unsigned long long)P.P;
    };

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&P.tmC);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&a_full[s], 128);
            ptx::mbar_init(&a_empty[s], 1);
            ptx::mbar_init(&t_full[s], 1);
            ptx::mbar_init(&t_empty[s], 4);
        }
        for (int s = 0; s < STEM_IN_SLOTS; ++s) { ptx::mbar_init(&in_full[s], 1); ptx::mbar_init(&in_empty[s], 128); }
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&tmem_base_smem, (uint32_t)P.tmem_cols);
        ptx::tmem_relinquish();
        // weights: C rows of 16 K-values (32 bytes), rows swizzled exactly as TMA would have written them
        // (two matrices: the high-order terms at smem_b, the low-order terms 2 KB behind)
        for (int i = lane; i < C * 4; i += 32) {
            const int part = i >= C * 2 ? 1 : 0, k = i - part * C * 2;
            const int n = k >> 1, c = k & 1;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.wB) + i);
            sts128(smem_b + (uint32_t)(part * 2048 + n * 32 + ((c ^ ((n >> 2) & 1)) << 4)), v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x >= 160 && (int)threadIdx.x - 160 < C) s_bias[threadIdx.x - 160] = __ldg(P.bias + (threadIdx.x - 160));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // =========================== MMA issuer ===========================
        const uint32_t idesc = ptx::make_idesc_f16(128, C, F16);
        const uint64_t b_desc = ptx::make_kmajor_desc(smem_b, 32), b_lo_desc = ptx::make_kmajor_desc(smem_b + 2048, 32);
        int i = 0;
        for (int m = blockIdx.x; m < P.n_macro; m += gridDim.x, ++i) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            ptx::mbar_wait(&t_empty[s], ph ^ 1u, P.err, ERR_STEM_MMA_WAIT);
            ptx::mbar_wait(&a_full[s], ph, P.err, ERR_STEM_MMA_WAIT);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < STEM_TC_SUB; ++j) {
                    const uint64_t a_desc = ptx::make_kmajor_desc(smem_a + (uint32_t)((s * STEM_TC_SUB + j) * 4096), 32);
                    ptx::umma_f16(tmem_base + (uint32_t)((s * STEM_TC_SUB + j) * C), a_desc, b_desc, idesc, 0u);
                    ptx::umma_f16(tmem_base + (uint32_t)((s * STEM_TC_SUB + j) * C), a_desc, b_lo_desc, idesc, 1u);   // + pixels x low-order weight terms
                }
                ptx::umma_commit(&a_empty[s]);
                ptx::umma_commit(&t_full[s]);
            }
            __syncwarp();
        }
    } else if (warp <= 4) {
        // =========================== im2col builders ===========================
        const int t = (int)threadIdx.x - 32;                            // A row inside a sub-tile
        const uint32_t row_off = (uint32_t)(t * 32);
        const uint32_t sw = (uint32_t)((t >> 2) & 1) << 4;
        const int W = P.W, H = P.H;
        int i = 0;
        uint32_t nf = 0;                                                // fast macro-tiles so far = uses of the input ring
        for (int m = blockIdx.x; m < P.n_macro; m += gridDim.x, ++i) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            ptx::mbar_wait(&a_empty[s], ph ^ 1u, P.err, ERR_STEM_BUILD_WAIT);
            const bool fast = fast_tile(m);
            uint32_t seg[3] = {0u, 0u, 0u};                             // shared address of this thread's centre byte per row
            const uint32_t slot = nf & (STEM_IN_SLOTS - 1);
            if (fast) {
                ptx::mbar_wait(&in_full[slot], (nf >> 2) & 1u, P.err, ERR_STEM_BUILD_WAIT);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t s0 = (uint32_t)m * 512u + (uint32_t)((k - 1) * W) - 1u;      // first byte the segment needs
                    seg[k] = smem_in + slot * STEM_IN_SLOT_BYTES + (uint32_t)(k * STEM_SEG_BYTES) + (s0 & 15u) + 1u + (uint32_t)t;
                }
            }
#pragma unroll
            for (int j = 0; j < STEM_TC_SUB; ++j) {
                const uint32_t p = (uint32_t)m * 512u + (uint32_t)(j * 128 + t);
                uint32_t v[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) v[k] = 0;
                if (fast) {
                    const uint32_t q = fdiv(p, P.fdW);                  // b*H + y
                    const int x = (int)(p - q * (uint32_t)W);
                    const int y = (int)(q - fdiv(q, P.fdH) * (uint32_t)H);
                    const uint32_t mrow[3] = {y > 0 ? 0xffu : 0u, 0xffu, y < H - 1 ? 0xffu : 0u};
                    const uint32_t ml = x > 0 ? 0xffu : 0u, mr = x < W - 1 ? 0xffu : 0u;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t a = seg[k] + (uint32_t)(j * 128);
                        v[k * 3 + 0] = lds_u8(a - 1u) & mrow[k] & ml;
                        v[k * 3 + 1] = lds_u8(a) & mrow[k];
                        v[k * 3 + 2] = lds_u8(a + 1u) & mrow[k] & mr;
                    }
                } else if (p < P.P) {
                    const uint32_t q = fdiv(p, P.fdW);                  // b*H + y
                    const int x = (int)(p - q * (uint32_t)W);
                    const int y = (int)(q - fdiv(q, P.fdH) * (uint32_t)H);
                    const uint8_t* c = P.x + p;
                    const bool up = y > 0, dn = y < H - 1, lf = x > 0, rt = x < W - 1;
                    if (up) { if (lf) v[0] = __ldg(c - W - 1); v[1] = __ldg(c - W); if (rt) v[2] = __ldg(c - W + 1); }
                    if (lf) v[3] = __ldg(c - 1);
                    v[4] = __ldg(c);
                    if (rt) v[5] = __ldg(c + 1);
                    if (dn) { if (lf) v[6] = __ldg(c + W - 1); v[7] = __ldg(c + W); if (rt) v[8] = __ldg(c + W + 1); }
                }
                uint32_t h[9];                                           // 16-bit patterns of the (exact) pixel values
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (F16) h[k] = (uint32_t)__half_as_ushort(__float2half_rn((float)v[k]));
                    else     h[k] = __float_as_uint((float)v[k]) >> 16;
                }
                const uint4 lo = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                const uint4 hi = make_uint4(h[8], 0u, 0u, 0u);
                const uint32_t base = smem_a + (uint32_t)((s * STEM_TC_SUB + j) * 4096) + row_off;
                sts128(base + sw, lo);
                sts128(base + (sw ^ 16u), hi);
            }
            if (fast) { ptx::mbar_arrive(&in_empty[slot]); ++nf; }        // this thread is done with the input slot
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async reads
            ptx::mbar_arrive(&a_full[s]);
        }
    } else if (warp == 13) {
        // =========================== input producer ===========================
        uint32_t nf = 0;
        for (int m = blockIdx.x; m < P.n_macro; m += gridDim.x) {
            if (!fast_tile(m)) continue;
            const uint32_t slot = nf & (STEM_IN_SLOTS - 1);
            ptx::mbar_wait(&in_empty[slot], ((nf >> 2) & 1u) ^ 1u, P.err, ERR_STEM_BUILD_WAIT);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(&in_full[slot], (uint32_t)STEM_IN_SLOT_BYTES);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t s0 = (uint32_t)m * 512u + (uint32_t)((k - 1) * P.W) - 1u;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_in + slot * STEM_IN_SLOT_BYTES + (uint32_t)(k * STEM_SEG_BYTES)), "l"(P.x + (s0 & ~15u)),
                                   "r"((uint32_t)STEM_SEG_BYTES), "r"(ptx::smem_u32(&in_full[slot])) : "memory");
                }
            }
            __syncwarp();
            ++nf;
        }
    } else {
        // =========================== epilogue groups ===========================
        const int g = (warp - 5) >> 2;
        const int quarter = warp & 3;                                   // TMEM lane quarter this warp may read
        const int r = quarter * 32 + lane;                              // pixel row inside a sub-tile
        const int etid = ((int)threadIdx.x - 160) & 127;
        const int CB = P.CB, nchunk = C / CB;
        const int pitch = CB * 2;
        const uint32_t swz_mask = (uint32_t)(pitch >> 4) - 1u;
        const uint32_t chunk_bytes = (uint32_t)(128 * pitch);
        const uint32_t cg = smem_c + (uint32_t)g * (uint32_t)(STEM_TC_SUB * 128 * C * 2);
        const uint32_t row_base = (uint32_t)(r * pitch);
        const uint32_t xr = ((row_base >> 7) & swz_mask) << 4;
        int k = 0;
        for (int m = blockIdx.x + g * gridDim.x; m < P.n_macro; m += 2 * gridDim.x, ++k) {
            if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous stores have left the staging tiles
            ptx::mbar_wait(&t_full[g], (uint32_t)k & 1u, P.err, ERR_STEM_EPI_WAIT);
            ptx::tc_fence_after();
            if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else        asm volatile("bar.sync 2, 128;" ::: "memory");
            for (int j = 0; j < STEM_TC_SUB; ++j) {
                for (int ch = 0; ch < nchunk; ++ch) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((g * STEM_TC_SUB + j) * C + ch * CB);
                    const uint32_t dst = cg + (uint32_t)(j * nchunk + ch) * chunk_bytes + row_base;
                    uint32_t a[32];
                    if (CB == 32) ptx::tmem_ld_32x32(taddr, a);
                    else          ptx::tmem_ld_32x16(taddr, a);
                    ptx::tmem_ld_wait();
                    const float* sb = s_bias + ch * CB;
#pragma unroll
                    for (int v8 = 0; v8 < 4; ++v8) {
                        if (v8 * 8 < CB) {
                            const float4 b0 = *reinterpret_cast<const float4*>(sb + v8 * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(sb + v8 * 8 + 4);
                            const uint32_t* q8 = a + v8 * 8;
                            const uint4 o = make_uint4(pack2_relu<F16>(__uint_as_float(q8[0]) + b0.x, __uint_as_float(q8[1]) + b0.y),
                                                       pack2_relu<F16>(__uint_as_float(q8[2]) + b0.z, __uint_as_float(q8[3]) + b0.w),
                                                       pack2_relu<F16>(__uint_as_float(q8[4]) + b1.x, __uint_as_float(q8[5]) + b1.y),
                                                       pack2_relu<F16>(__uint_as_float(q8[6]) + b1.z, __uint_as_float(q8[7]) + b1.w));
                            sts128(dst + ((uint32_t)(v8 * 16) ^ xr), o);
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&t_empty[g]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else        asm volatile("bar.sync 2, 128;" ::: "memory");
            if (etid == 0) {
                for (int j = 0; j < STEM_TC_SUB; ++j)
                    for (int ch = 0; ch < nchunk; ++ch)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"((uint64_t)&P.tmC), "r"(cg + (uint32_t)(j * nchunk + ch) * chunk_bytes), "r"(ch * CB),
                                       "r"((int)((uint32_t)m * 512u + (uint32_t)(j * 128))) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

}  // namespace aau
