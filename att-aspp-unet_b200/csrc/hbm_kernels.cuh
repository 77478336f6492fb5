// HBM-bound kernels of the AttentionASPPUNet forward: everything that is not a tensor-core GEMM.
// All activations are NHWC with 16-bit elements (bf16 default, fp16 optional); a tensor inside a concatenated
// buffer is addressed by (pixel stride `ld` in channels, channel offset `choff`).  Every thread moves 16-byte
// vectors (8 channels) and neighbouring threads touch neighbouring addresses.
#pragma once
#include "igemm_tc.cuh"

namespace aau {

// ---------------------------------------------------------------------------------------------------------
// d1.0: Conv2d(1 -> C, 3x3, pad 1, no bias) + BN(folded) + ReLU straight from the input frame.
// (attention_aspp_unet_pipeline_stage.py:114 first ConvBNReLU; input normalisation `u8/255` as
//  model_attention_aspp.py:17.)  K = 9 is far too small for the tensor cores: one thread per pixel, fp32 FMAs,
// weights broadcast from shared memory, the C output channels of a pixel written as 16-byte vectors.
// x_dtype: 0 = float32 [B,1,H,W] (already in [0,1]), 1 = uint8 [B,H,W] (normalised here as float(u8)/255.0f).
__device__ __forceinline__ float load_px(const void* __restrict__ x, int x_dtype, long long idx) {
    return x_dtype == 0 ? __ldg((const float*)x + idx) : (float)__ldg((const uint8_t*)x + idx) / 255.0f;
}
// Two horizontally adjacent pixels per thread: the 9x4 weights of a 4-channel group are read from shared memory
// as nine broadcast 16-byte loads and used for 72 FMAs, and the thread writes 2*C*2 contiguous bytes.
__global__ void __launch_bounds__(256) stem_conv3x3_kernel(const void* __restrict__ x, int x_dtype, int B, int H, int W,
                                                           const float* __restrict__ w9c,   // [9][C], BN scale folded in
                                                           const float* __restrict__ bias,  // [C]
                                                           uint8_t* __restrict__ out, int out_ld, int out_choff, int C, int is_fp16) {
    extern __shared__ __align__(16) float s_w[];      // [9][C] weights then [C] bias
    for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) s_w[i] = i < 9 * C ? w9c[i] : bias[i - 9 * C];
    __syncthreads();
    const int Wp = (W + 1) >> 1;                                  // pixel pairs per row
    const long long npair = (long long)B * H * Wp;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npair; p += (long long)gridDim.x * blockDim.x) {
        const int xp = (int)(p % Wp);
        const long long t = p / Wp;
        const int y = (int)(t % H);
        const long long b = t / H;
        const int x0 = xp * 2;
        const bool two = x0 + 1 < W;
        float v[3][4];                                            // rows y-1..y+1, columns x0-1..x0+2
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int yy = y + ky - 1, xx = x0 + kx - 1;
                v[ky][kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? load_px(x, x_dtype, (b * H + yy) * W + xx) : 0.f;
            }
        uint8_t* dst0 = out + ((((size_t)b * H + y) * W + x0) * out_ld + out_choff) * 2;
        uint8_t* dst1 = dst0 + (size_t)out_ld * 2;
        for (int c0 = 0; c0 < C; c0 += 8) {
            float a0[8], a1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a0[i] = a1[i] = s_w[9 * C + c0 + i];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 wa = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * C + c0);
                    const float4 wb = *reinterpret_cast<const float4*>(s_w + (ky * 3 + kx) * C + c0 + 4);
                    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        a0[i] = fmaf(v[ky][kx], w[i], a0[i]);
                        a1[i] = fmaf(v[ky][kx + 1], w[i], a1[i]);
                    }
                }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a0[i] = fmaxf(a0[i], 0.f); a1[i] = fmaxf(a1[i], 0.f); }
            *reinterpret_cast<uint4*>(dst0 + c0 * 2) = make_uint4(pack2(a0[0], a0[1], is_fp16), pack2(a0[2], a0[3], is_fp16),
                                                                  pack2(a0[4], a0[5], is_fp16), pack2(a0[6], a0[7], is_fp16));
            if (two)
                *reinterpret_cast<uint4*>(dst1 + c0 * 2) = make_uint4(pack2(a1[0], a1[1], is_fp16), pack2(a1[2], a1[3], is_fp16),
                                                                      pack2(a1[4], a1[5], is_fp16), pack2(a1[6], a1[7], is_fp16));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// MaxPool2d(2) with floor semantics (attention_aspp_unet_pipeline_stage.py:115-118): the last odd row/column is
// never read.  One thread = one output pixel x 8 channels.
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const uint8_t* __restrict__ in, int in_ld, int in_choff, int B, int H, int W, int C,
                                                         uint8_t* __restrict__ out, int is_fp16) {
    const int OH = H >> 1, OW = W >> 1, CV = C >> 3;
    const long long total = (long long)B * OH * OW * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long t = i / CV;
        const int ox = (int)(t % OW);
        t /= OW;
        const int oy = (int)(t % OH);
        const long long b = t / OH;
        const uint8_t* p00 = in + ((((size_t)b * H + 2 * oy) * W + 2 * ox) * in_ld + in_choff + cv * 8) * 2;
        const size_t px = (size_t)in_ld * 2, row = (size_t)W * in_ld * 2;
        const uint4 a = __ldg((const uint4*)p00), bq = __ldg((const uint4*)(p00 + px));
        const uint4 c = __ldg((const uint4*)(p00 + row)), d = __ldg((const uint4*)(p00 + row + px));
        *reinterpret_cast<uint4*>(out + (size_t)i * 16) = max8(max8(a, bq, is_fp16), max8(c, d, is_fp16), is_fp16);
    }
}

// ---------------------------------------------------------------------------------------------------------
// ASPP image-pooling branch folded into a per-image bias of the project conv (SURVEY.md identity i2;
// attention_aspp_unet_pipeline_stage.py:75-77,80-83): bilinear up-sampling of a 1x1 map is a broadcast, so
//   bias_img[b][o] = sum_j Wproj[o][4*Co + j] * relu(sum_i Wpool[j][i] * mean_hw(x[b,:,:,i]) + bpool[j]) + bproj[o].
// Kernel 1 (grid = splits x frames): per-channel partial sums over a slice of the frame's pixels, coalesced
// 4-byte (2-channel) loads, fixed summation order (no atomics -> deterministic).
// Kernel 2 finishes the mean and runs the two small mat-vecs (fp32 weights).
__global__ void __launch_bounds__(256) gap_partial_kernel(const uint8_t* __restrict__ x, int HW, int Cin, int splits,
                                                          float* __restrict__ partial /* [B][splits][Cin] */, int is_fp16) {
    extern __shared__ float s_buf[];                 // [lanes][Cin]
    const int b = blockIdx.y, sp = blockIdx.x;
    const int pairs = Cin >> 1;
    const int lanes = blockDim.x / pairs;
    const int pl = threadIdx.x / pairs, cp = threadIdx.x % pairs;
    const int per = (HW + splits - 1) / splits;
    const int p0 = sp * per, p1 = min(HW, p0 + per);
    if (pl < lanes) {
        float s0 = 0.f, s1 = 0.f;
        const uint32_t* base = reinterpret_cast<const uint32_t*>(x + (size_t)b * HW * Cin * 2) + cp;
        for (int p = p0 + pl; p < p1; p += lanes) {
            const float2 f = unpack2(__ldg(base + (size_t)p * pairs), is_fp16);
            s0 += f.x;
            s1 += f.y;
        }
        s_buf[pl * Cin + 2 * cp] = s0;
        s_buf[pl * Cin + 2 * cp + 1] = s1;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += s_buf[l * Cin + c];
        partial[((size_t)b * splits + sp) * Cin + c] = s;
    }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// grid = (frames, output slices).  Every block recomputes the (cheap) pooled 1x1 conv `v`, then produces its slice
// of the per-image bias.  One warp per output channel, lanes split the reduction (coalesced weight rows).
__global__ void __launch_bounds__(512) aspp_pool_bias_kernel(const float* __restrict__ partial, int splits, int HW, int Cin, int Cout,
                                                             const float* __restrict__ wpool,    // [Cout][Cin], BN folded
                                                             const float* __restrict__ bpool,    // [Cout]
                                                             const float* __restrict__ wproj,    // [Cout][Cout] pool slice of project, BN folded
                                                             const float* __restrict__ bproj,    // [Cout]
                                                             float* __restrict__ bias_img) {
    extern __shared__ float s_buf[];                 // mean[Cin] | v[Cout]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float* mean = s_buf;
    float* v = mean + Cin;
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)b * splits + sp) * Cin + c];
        mean[c] = s / (float)HW;
    }
    __syncthreads();
    for (int o = warp; o < Cout; o += nwarp) {
        float s = 0.f;
        for (int i = lane; i < Cin; i += 32) s = fmaf(__ldg(wpool + (size_t)o * Cin + i), mean[i], s);
        s = warp_sum_f(s);
        if (lane == 0) v[o] = fmaxf(s + bpool[o], 0.f);
    }
    __syncthreads();
    const int per = (Cout + gridDim.y - 1) / gridDim.y;
    const int o0 = blockIdx.y * per, o1 = min(Cout, o0 + per);
    for (int o = o0 + warp; o < o1; o += nwarp) {
        float s = 0.f;
        for (int j = lane; j < Cout; j += 32) s = fmaf(__ldg(wproj + (size_t)o * Cout + j), v[j], s);
        s = warp_sum_f(s);
        if (lane == 0) bias_img[(size_t)b * Cout + o] = s + bproj[o];
    }
}

// ---------------------------------------------------------------------------------------------------------
// F.interpolate(mode="bilinear", align_corners=False) from (IH, IW) to (OH, OW) -- the decoder's fix-up when the
// floor-pooled skip is one row/column larger than the transposed-conv output
// (attention_aspp_unet_pipeline_stage.py:106-107).  Index arithmetic follows ATen's upsample_bilinear2d:
// scale = in/out, src = max(0, scale*(dst+0.5)-0.5), i0 = floor(src), i1 = min(i0+1, in-1), w1 = src-i0.
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const uint8_t* __restrict__ in, int IH, int IW, int C,
                                                              uint8_t* __restrict__ out, int OH, int OW, int out_ld, int out_choff,
                                                              int B, int is_fp16) {
    const int CV = C >> 3;
    const float sh = (float)IH / (float)OH, sw = (float)IW / (float)OW;
    const long long total = (long long)B * OH * OW * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long t = i / CV;
        const int ox = (int)(t % OW);
        t /= OW;
        const int oy = (int)(t % OH);
        const long long b = t / OH;
        float fy = fmaxf(sh * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sw * ((float)ox + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = min(y0 + 1, IH - 1), x1 = min(x0 + 1, IW - 1);
        const float wy1 = fy - (float)y0, wx1 = fx - (float)x0;
        const float wy0 = 1.f - wy1, wx0 = 1.f - wx1;
        const uint8_t* base = in + (size_t)b * IH * IW * C * 2 + (size_t)cv * 16;
        const uint4 q00 = __ldg((const uint4*)(base + ((size_t)y0 * IW + x0) * C * 2));
        const uint4 q01 = __ldg((const uint4*)(base + ((size_t)y0 * IW + x1) * C * 2));
        const uint4 q10 = __ldg((const uint4*)(base + ((size_t)y1 * IW + x0) * C * 2));
        const uint4 q11 = __ldg((const uint4*)(base + ((size_t)y1 * IW + x1) * C * 2));
        const uint32_t a00[4] = {q00.x, q00.y, q00.z, q00.w}, a01[4] = {q01.x, q01.y, q01.z, q01.w};
        const uint32_t a10[4] = {q10.x, q10.y, q10.z, q10.w}, a11[4] = {q11.x, q11.y, q11.z, q11.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 v00 = unpack2(a00[k], is_fp16), v01 = unpack2(a01[k], is_fp16);
            const float2 v10 = unpack2(a10[k], is_fp16), v11 = unpack2(a11[k], is_fp16);
            const float rx = wy0 * (wx0 * v00.x + wx1 * v01.x) + wy1 * (wx0 * v10.x + wx1 * v11.x);
            const float ry = wy0 * (wx0 * v00.y + wx1 * v01.y) + wy1 * (wx0 * v10.y + wx1 * v11.y);
            o[k] = pack2(rx, ry, is_fp16);
        }
        uint8_t* dst = out + ((((size_t)b * OH + oy) * OW + ox) * out_ld + out_choff + cv * 8) * 2;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Selection head: per-frame sigmoid / threshold / mask area, then first-max argmax over frames
// (model_attention_aspp.py:54 sigmoid, :71 `prob > 0.05`, :74 / :94 `sum((1,2)).argmax()`).
// Each thread reads float4 logits, evaluates sigmoid in fp32 exactly as `1/(1+exp(-x))`, counts, then the counts
// are reduced with warp shuffles, one shared-memory hop per block and ONE integer atomic per block, so the result
// is exact and order independent.  The optional mask output is the uint8 {0,1} volume the reference builds.
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int above(float v, float thr, int is_prob) {
    return (is_prob ? v : 1.f / (1.f + expf(-v))) > thr;
}
__global__ void __launch_bounds__(256) frame_area_kernel(const float* __restrict__ logits, int is_prob, int HW, float thr,
                                                         int* __restrict__ areas, uint8_t* __restrict__ mask) {
    const int frame = blockIdx.y;
    const float* src = logits + (size_t)frame * HW;
    uint8_t* m = mask ? mask + (size_t)frame * HW : nullptr;
    int cnt = 0;
    const bool vec_ok = ((HW & 3) == 0) && ((((size_t)frame * HW) & 3) == 0);
    if (vec_ok) {
        const int n4 = HW >> 2;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
            const int b0 = above(v.x, thr, is_prob), b1 = above(v.y, thr, is_prob);
            const int b2 = above(v.z, thr, is_prob), b3 = above(v.w, thr, is_prob);
            cnt += b0 + b1 + b2 + b3;
            if (m) reinterpret_cast<uint32_t*>(m)[i] = (uint32_t)b0 | ((uint32_t)b1 << 8) | ((uint32_t)b2 << 16) | ((uint32_t)b3 << 24);
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
            const int b0 = above(__ldg(src + i), thr, is_prob);
            cnt += b0;
            if (m) m[i] = (uint8_t)b0;
        }
    }
    __shared__ int s_part[8];
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0;
        v = warp_sum(v);
        if (threadIdx.x == 0 && v) atomicAdd(areas + frame, v);
    }
}

// `mask_3d.sum((1,2))` on a uint8 volume (model_attention_aspp.py:94): per-frame sum of byte VALUES (equal to the
// area for {0,1} masks); optional binarised copy `(v > 0)` (model_attention_aspp.py:97).
__global__ void __launch_bounds__(256) frame_sum_u8_kernel(const uint8_t* __restrict__ vol, int HW, int* __restrict__ sums,
                                                           uint8_t* __restrict__ mask) {
    const int frame = blockIdx.y;
    const uint8_t* src = vol + (size_t)frame * HW;
    uint8_t* m = mask ? mask + (size_t)frame * HW : nullptr;
    int s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int v = __ldg(src + i);
        s += v;
        if (m) m[i] = v > 0;
    }
    __shared__ int s_part[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0;
        v = warp_sum(v);
        if (threadIdx.x == 0 && v) atomicAdd(sums + frame, v);
    }
}

// first index of the maximum area (numpy argmax tie-break); out[0] = index, out[1] = area at that index
__global__ void __launch_bounds__(1024) area_argmax_kernel(const int* __restrict__ areas, int n, int* __restrict__ out) {
    int best = -1, best_i = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int a = areas[i];
        if (a > best) { best = a; best_i = i; }       // strided scan keeps the smallest index per thread on ties
    }
    __shared__ int s_v[32], s_i[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ov = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = best_i; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        best = threadIdx.x < nw ? s_v[threadIdx.x] : -1;
        best_i = threadIdx.x < nw ? s_i[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ov = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (threadIdx.x == 0) { out[0] = best_i; out[1] = best; }
    }
}

}  // namespace aau
