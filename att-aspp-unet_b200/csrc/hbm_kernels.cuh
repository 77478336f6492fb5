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
// Tile = STEM_TR rows x (256 / (C/8)) columns of one frame.  The (halo-padded, already normalised) input tile sits in
// shared memory, every value duplicated into a float2 so that one LDS.64 yields the (v, v) operand of a packed FMA;
// a thread owns ONE column and ONE group of 8 output channels, keeps its 9x8 folded weights in registers as 36
// float pairs and walks down the rows with a sliding 3x3 window: per pixel 3 shared-memory reads and 36 FFMA2
// (fma.rn.f32x2: two independent IEEE fp32 FMAs per instruction, sm_100) -- the kernel is bound by instruction
// issue, not by its 1 byte in / 2*C bytes out per pixel of HBM traffic, so halving the FMA instructions is the lever.
// The C/8 threads of a pixel write its C*2 bytes back to back and a warp covers 32/(C/8) neighbouring pixels, so
// every store instruction writes whole contiguous 128-byte lines.
// uint8 input is normalised through a 256-entry table of float(v)/255.0f (one IEEE division per thread per block).
enum { STEM_TR = 16, STEM_MAX_TW = 128 };
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pair_f32(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_pair_relu(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return pack2_relu<F16>(lo, hi);
}
template <bool F16>
__global__ void __launch_bounds__(256, 2) stem_conv3x3_kernel(const void* __restrict__ x, int x_dtype, int B, int H, int W,
                                                           const float* __restrict__ w9c,   // [9][C], BN scale folded in
                                                           const float* __restrict__ bias,  // [C]
                                                           uint8_t* __restrict__ out, int out_ld, int out_choff, int C,
                                                           int tiles_x, int tiles_y) {
    __shared__ __align__(8) float2 s_in[(STEM_TR + 2) * (STEM_MAX_TW + 2)];
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    const int CG = C >> 3;                                        // 8-channel groups per pixel
    const int TWc = min(256 / CG, (int)STEM_MAX_TW);              // columns per tile
    const int pitch = TWc + 2;
    const uint32_t pitch_mul = ((1u << 20) + pitch - 1) / pitch;  // i / pitch == (i * pitch_mul) >> 20 for i < 2^11
    const int cg = threadIdx.x % CG, col = threadIdx.x / CG;
    unsigned long long w2[9][4], bz2[4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w9c + t * C + cg * 8));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(w9c + t * C + cg * 8 + 4));
        w2[t][0] = pair_f32(a.x, a.y); w2[t][1] = pair_f32(a.z, a.w); w2[t][2] = pair_f32(b4.x, b4.y); w2[t][3] = pair_f32(b4.z, b4.w);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) bz2[i] = pair_f32(__ldg(bias + cg * 8 + 2 * i), __ldg(bias + cg * 8 + 2 * i + 1));
    const int per_img = tiles_x * tiles_y, ntiles = B * per_img;
    const int n_in = (STEM_TR + 2) * pitch;                       // input pixels of a tile (with halo)
    constexpr int FILL = ((STEM_TR + 2) * (STEM_MAX_TW + 2) + 255) / 256;   // input pixels per thread
    // uint8 frames: the NEXT tile's input bytes are fetched into three registers while the current tile is being
    // computed (a tile's fill was otherwise five dependent global-load latencies long); out-of-frame pixels are 0,
    // which the table maps to 0.0f -- the zero padding of the convolution.
    uint32_t raw[(FILL + 3) / 4];
    auto fetch_u8 = [&](int tile) {
#pragma unroll
        for (int k = 0; k < (FILL + 3) / 4; ++k) raw[k] = 0;
        if (tile >= ntiles) return;
        const int b = tile / per_img, r0 = tile - b * per_img;
        const int tyi = r0 / tiles_x, txi = r0 - tyi * tiles_x;
        const int y0 = tyi * STEM_TR, x0 = txi * TWc;
#pragma unroll
        for (int k = 0; k < FILL; ++k) {
            const int i = threadIdx.x + k * 256;
            const int r = (int)(((uint32_t)i * pitch_mul) >> 20), cc = i - r * pitch;
            const int yy = y0 + r - 1, xx = x0 + cc - 1;
            if (i < n_in && yy >= 0 && yy < H && xx >= 0 && xx < W)
                raw[k >> 2] |= (uint32_t)__ldg((const uint8_t*)x + ((size_t)b * H + yy) * W + xx) << (8 * (k & 3));
        }
    };
    if (x_dtype != 0) fetch_u8(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / per_img, r0 = tile - b * per_img;
        const int tyi = r0 / tiles_x, txi = r0 - tyi * tiles_x;
        const int y0 = tyi * STEM_TR, x0 = txi * TWc;
        __syncthreads();                                          // the previous tile has been consumed (and s_lut is written)
        if (x_dtype != 0) {
#pragma unroll
            for (int k = 0; k < FILL; ++k) {
                const int i = threadIdx.x + k * 256;
                if (i < n_in) {
                    const float v = s_lut[(raw[k >> 2] >> (8 * (k & 3))) & 0xffu];
                    s_in[i] = make_float2(v, v);
                }
            }
        } else {
            for (int i = threadIdx.x; i < n_in; i += 256) {
                const int r = (int)(((uint32_t)i * pitch_mul) >> 20), cc = i - r * pitch;
                const int yy = y0 + r - 1, xx = x0 + cc - 1;
                float v = 0.f;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg((const float*)x + ((size_t)b * H + yy) * W + xx);
                s_in[i] = make_float2(v, v);
            }
        }
        if (x_dtype != 0) fetch_u8(tile + gridDim.x);             // in flight during the compute phase below
        __syncthreads();
        if (col < TWc && x0 + col < W) {
            const unsigned long long* s2 = reinterpret_cast<const unsigned long long*>(s_in);
            unsigned long long v[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { v[0][k] = s2[col + k]; v[1][k] = s2[pitch + col + k]; }
            uint8_t* dst = out + ((((size_t)b * H + y0) * W + x0 + col) * out_ld + out_choff + cg * 8) * 2;
            const size_t row_bytes = (size_t)W * out_ld * 2;
            const int nrow = min((int)STEM_TR, H - y0);
#pragma unroll 3
            for (int r = 0; r < nrow; ++r) {
#pragma unroll
                for (int k = 0; k < 3; ++k) v[2][k] = s2[(r + 2) * pitch + col + k];
                unsigned long long a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = bz2[i];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = ffma2(v[ky][kx], w2[ky * 3 + kx][i], a[i]);
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_pair_relu<F16>(a[0]), pack_pair_relu<F16>(a[1]),
                                                            pack_pair_relu<F16>(a[2]), pack_pair_relu<F16>(a[3]));
                dst += row_bytes;
#pragma unroll
                for (int k = 0; k < 3; ++k) { v[0][k] = v[1][k]; v[1][k] = v[2][k]; }
            }
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
    // pixel lanes: 256 / pairs threads share a channel pair; with more than 256 pairs (base_c >= 80) every thread walks
    // several channel pairs over ALL pixels of the slice instead (lanes == 1)
    const int lanes = max(1, (int)blockDim.x / pairs);
    const int per = (HW + splits - 1) / splits;
    const int p0 = sp * per, p1 = min(HW, p0 + per);
    for (int slot = threadIdx.x; slot < lanes * pairs; slot += blockDim.x) {
        const int pl = slot / pairs, cp = slot - pl * pairs;
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
// grid = (ceil(OW*C/8 / 256), ceil(OH / RESIZE_ROWS), B): a thread produces 8 channels (one 16-byte vector) of one output column.  An axis
// whose size does not change has scale 1, so its second sample has weight exactly 0 and is neither read nor blended
// (v*1 + u*0 == v for finite u): the usual one-axis fix-up is two reads (adjacent rows or pixels: L2 hits), eight
// 2-term blends and one write per thread.  Instantiated per storage type so that unpack / pack are two instructions.
template <bool F16>
__device__ __forceinline__ void blend8(const uint4& p, const uint4& q, float wp, float wq, float (&o)[8]) {
    const uint32_t a[4] = {p.x, p.y, p.z, p.w}, b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 u = unpack2(a[k], F16), v = unpack2(b[k], F16);
        o[2 * k] = wp * u.x + wq * v.x;
        o[2 * k + 1] = wp * u.y + wq * v.y;
    }
}
// A thread walks RESIZE_ROWS consecutive output rows of its (column, 8-channel) slot and keeps the lower source row
// of one output row as the upper source row of the next when they coincide (they do whenever the scale is ~1), so
// the one-row fix-up reads every input row once.
enum { RESIZE_ROWS = 8 };
template <bool F16>
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const uint8_t* __restrict__ in, int IH, int IW, int C,
                                                              uint8_t* __restrict__ out, int OH, int OW, int out_ld, int out_choff,
                                                              int B) {
    const int CV = C >> 3;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= OW * CV) return;
    const int ox = i / CV, cv = i - ox * CV;
    const int oy0 = blockIdx.y * RESIZE_ROWS, b = blockIdx.z;
    const float sh = (float)IH / (float)OH, sw = (float)IW / (float)OW;
    const float fx = fmaxf(sw * ((float)ox + 0.5f) - 0.5f, 0.f);
    const int x0 = (int)fx, x1 = min(x0 + 1, IW - 1);
    const float wx1 = fx - (float)x0, wx0 = 1.f - wx1;
    const bool two_x = IW != OW, two_y = IH != OH;
    const uint8_t* base = in + (size_t)b * IH * IW * C * 2 + (size_t)cv * 16;
    auto load_row = [&](int y, float (&r)[8]) {                   // horizontally blended source row y at this output column
        const uint4 q0 = __ldg((const uint4*)(base + ((size_t)y * IW + x0) * C * 2));
        if (two_x) blend8<F16>(q0, __ldg((const uint4*)(base + ((size_t)y * IW + x1) * C * 2)), wx0, wx1, r);
        else       blend8<F16>(q0, q0, 1.f, 0.f, r);
    };
    float top[8], bot[8];
    int have_top = -1, have_bot = -1;                             // source rows currently held
    uint8_t* dst = out + ((((size_t)b * OH + oy0) * OW + ox) * out_ld + out_choff + cv * 8) * 2;
    const size_t row_bytes = (size_t)OW * out_ld * 2;
#pragma unroll 1
    for (int k = 0; k < RESIZE_ROWS; ++k, dst += row_bytes) {
        const int oy = oy0 + k;
        if (oy >= OH) break;
        const float fy = fmaxf(sh * ((float)oy + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy, y1 = min(y0 + 1, IH - 1);
        const float wy1 = fy - (float)y0, wy0 = 1.f - wy1;
        if (y0 != have_top) {
            if (y0 == have_bot) {
#pragma unroll
                for (int j = 0; j < 8; ++j) top[j] = bot[j];
            } else {
                load_row(y0, top);
            }
            have_top = y0;
        }
        float r[8];
        if (two_y) {
            if (y1 != have_bot) { load_row(y1, bot); have_bot = y1; }
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = wy0 * top[j] + wy1 * bot[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = top[j];
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(r[0], r[1], F16), pack2(r[2], r[3], F16), pack2(r[4], r[5], F16), pack2(r[6], r[7], F16));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Selection head: per-frame threshold / mask area, then first-max argmax over frames
// (model_attention_aspp.py:54 sigmoid, :71 `prob > 0.05`, :74 / :94 `sum((1,2)).argmax()`).
// sigmoid is monotone, so `sigmoid(l) > t` is `l > cut` for ONE fp32 cutoff (SURVEY.md identity i7): the host finds it by
// bisection over the fp32 number line (aau_engine.cu: logit_cutoff -- correctly rounded fp32 sigmoid; the Python layer passes
// the cutoff of the HOST's own torch.sigmoid instead, AAU_IN_LOGIT_CUT), and no transcendental runs per pixel: the decision
// is a pure comparison, bit exact.  Each thread reads float4 values and counts; counts are reduced with warp shuffles, one
// shared-memory hop per block and ONE integer atomic per block, so the result is exact and order independent.  The
// optional mask output is the uint8 {0,1} volume the reference builds.
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int above(float v, float cut, int /*is_prob*/) { return v > cut; }
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

// prob = sigmoid(logits), fp32 -> fp32 (model_attention_aspp.py:54 `torch.sigmoid(self.net(...))`), same expression as `above()`
__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = 1.f / (1.f + expf(-__ldg(x + i)));
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

// {0,1} mask of the frame the argmax picked, without the index ever visiting the host: best[0] = frame, best[1] = its area
// (all-zero mask when the area is 0, model_attention_aspp.py:75-76).  Lets a sweep's whole device part be enqueued in one go,
// so the host tail of one sweep (connected components of this one mask) overlaps the kernels of the next sweep.
__global__ void __launch_bounds__(256) best_frame_mask_kernel(const float* __restrict__ values, int HW, float cut, const int* __restrict__ best,
                                                              uint8_t* __restrict__ mask) {
    const int frame = best[0], area = best[1];
    const float* src = values + (size_t)frame * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
        mask[i] = area > 0 ? (uint8_t)(__ldg(src + i) > cut) : (uint8_t)0;
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

namespace aau {

// ---------------------------------------------------------------------------------------------------------
// Frame conditioning: what the reference does to every frame before the network,
//   cv2.normalize(NORM_MINMAX, 0..255) -> uint8, cv2.createCLAHE(1.0, (8, 8)).apply, cv2.medianBlur(3)
// (model_attention_aspp.py:11-17, inference.py:147-190), reproduced BIT-EXACTLY on uint8 frames: integer histogram
// work plus the few fp32 operations OpenCV performs, in the same order and with the same roundings
//   * normalize : saturate(rint(fmaf(v, (float)scale, (float)shift))), scale = 255 * (1 / (max - min)) in double
//   * CLAHE LUT : per tile histogram over the REFLECT_101-padded frame, clip = max(int(clipLimit*area/256), 1), excess
//                 redistributed (batch + strided residual), lut[i] = saturate(rint(cumsum[i] * (255.f / area)))
//   * CLAHE blend: four LUT values, fp32 mul / add WITHOUT contraction, rint;  median: 3x3, replicated border.
// HBM-bound byte work: three reads and one write of the sweep; the LUTs (16 KB per frame) stay in L2.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_minmax_kernel(const uint8_t* __restrict__ x, int HW, int* __restrict__ mm /* [N][2] = {min, max} */) {
    const int frame = blockIdx.y;
    const uint8_t* src = x + (size_t)frame * HW;
    uint32_t lo4 = 0xffffffffu, hi4 = 0u;                          // four byte lanes each (SIMD-in-word min / max)
    const int head = (int)((16 - ((uintptr_t)src & 15)) & 15);     // bytes before the first 16-byte boundary
    const int n16 = HW > head ? (HW - head) >> 4 : 0;
    const uint4* v16 = reinterpret_cast<const uint4*>(src + head);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
        const uint4 v = __ldg(v16 + i);
        lo4 = __vminu4(__vminu4(lo4, v.x), __vminu4(v.y, __vminu4(v.z, v.w)));
        hi4 = __vmaxu4(__vmaxu4(hi4, v.x), __vmaxu4(v.y, __vmaxu4(v.z, v.w)));
    }
    int lo = min(min(lo4 & 255, (lo4 >> 8) & 255), min((lo4 >> 16) & 255, lo4 >> 24));
    int hi = max(max(hi4 & 255, (hi4 >> 8) & 255), max((hi4 >> 16) & 255, hi4 >> 24));
    if (blockIdx.x == 0) {                                          // unaligned head and tail bytes
        const int tail0 = head + n16 * 16;
        for (int i = threadIdx.x; i < min(head, HW); i += blockDim.x) { const int v = __ldg(src + i); lo = min(lo, v); hi = max(hi, v); }
        for (int i = tail0 + threadIdx.x; i < HW; i += blockDim.x) { const int v = __ldg(src + i); lo = min(lo, v); hi = max(hi, v); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 2 * frame, lo);
        atomicMax(mm + 2 * frame + 1, hi);
    }
}
__global__ void minmax_init_kernel(int* mm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mm[2 * i] = 255; mm[2 * i + 1] = 0; }
}

struct NormCoef { float a, b; };
__device__ __forceinline__ NormCoef norm_coef(const int* mm, int frame) {
    const double smin = (double)mm[2 * frame], smax = (double)mm[2 * frame + 1];
    const double scale = 255.0 * (smax - smin > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    NormCoef c;
    c.a = (float)scale;
    c.b = (float)(0.0 - smin * scale);
    return c;
}
__device__ __forceinline__ int norm_u8(int v, NormCoef c) {
    return min(255, max(0, __float2int_rn(__fmaf_rn((float)v, c.a, c.b))));
}

// grid = (tilesX * tilesY, N): one block builds the 256-entry LUT of one CLAHE tile.  Every warp counts into its own
// 256-bin histogram (speckle frames put most pixels into a few bins: shared-memory atomics of one warp serialise on
// those, eight private copies divide the contention) and walks whole tile rows, so there is no per-pixel division.
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ x, int H, int W, const int* __restrict__ mm,
                                                        int tilesX, int tilesY, int tw, int th, int clip,
                                                        uint8_t* __restrict__ lut /* [N][tilesY*tilesX][256] */) {
    __shared__ int whist[8][256];
    __shared__ int wsum[8];
    __shared__ int s_clipped;
    const int frame = blockIdx.y, tile = blockIdx.x;
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const NormCoef nc = norm_coef(mm, frame);
#pragma unroll
    for (int w = 0; w < 8; ++w) whist[w][threadIdx.x] = 0;
    if (threadIdx.x == 0) s_clipped = 0;
    __syncthreads();
    const uint8_t* src = x + (size_t)frame * H * W;
    for (int r = warp; r < th; r += 8) {
        int yy = ty * th + r;                                       // row in the padded frame
        if (yy >= H) yy = 2 * (H - 1) - yy;                         // BORDER_REFLECT_101
        const uint8_t* row = src + (size_t)yy * W;
        for (int c = lane; c < tw; c += 32) {
            int xx = tx * tw + c;
            if (xx >= W) xx = 2 * (W - 1) - xx;
            atomicAdd(&whist[warp][norm_u8(__ldg(row + xx), nc)], 1);
        }
    }
    __syncthreads();
    int h = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) h += whist[w][threadIdx.x];
    if (clip > 0) {
        const int over = max(h - clip, 0);
        int o = over;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) o += __shfl_xor_sync(0xffffffffu, o, s);
        if (lane == 0 && o) atomicAdd(&s_clipped, o);
        __syncthreads();
        const int clipped = s_clipped;
        h = min(h, clip) + clipped / 256;
        const int residual = clipped - (clipped / 256) * 256;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            if (threadIdx.x % step == 0 && threadIdx.x / step < residual) ++h;
        }
    }
    // inclusive prefix sum over the 256 bins
    int s = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    s += base;
    const float lut_scale = 255.0f / (float)(tw * th);
    lut[((size_t)frame * tilesX * tilesY + tile) * 256 + threadIdx.x] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn((float)s, lut_scale))));
}

// four pixels per 32-bit word: compare-exchange of byte lanes
__device__ __forceinline__ void sort2x4(uint32_t& a, uint32_t& b) { const uint32_t t = __vminu4(a, b); b = __vmaxu4(a, b); a = t; }
__device__ __forceinline__ uint32_t median9x4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t p5, uint32_t p6,
                                              uint32_t p7, uint32_t p8) {      // 19-exchange median-of-9 network, byte-wise
    sort2x4(p1, p2); sort2x4(p4, p5); sort2x4(p7, p8); sort2x4(p0, p1); sort2x4(p3, p4); sort2x4(p6, p7);
    sort2x4(p1, p2); sort2x4(p4, p5); sort2x4(p7, p8); sort2x4(p0, p3); sort2x4(p5, p8); sort2x4(p4, p7);
    sort2x4(p3, p6); sort2x4(p1, p4); sort2x4(p2, p5); sort2x4(p4, p7); sort2x4(p4, p2); sort2x4(p6, p4);
    sort2x4(p4, p2);
    return p4;
}

// grid = (ceil(W/COND_TW), ceil(H/COND_TH), N).  Phase 1: the CLAHE-blended values of a COND_TW x COND_TH output tile
// plus a replicated one-pixel halo go to shared memory (row stride padded to words); the interpolation indices and
// weights of the tile's rows / columns are tabulated once per block.  Phase 2: a thread takes the 3x3 medians of FOUR
// horizontally adjacent pixels at once with byte-lane min / max (the unaligned 3x3 window words come from two aligned
// shared-memory words and a funnel shift) and writes one 32-bit word.
enum { COND_TW = 128, COND_TH = 16, COND_PITCH = COND_TW + 8 };     // halo column at byte 3 of each row, data from byte 4
__global__ void __launch_bounds__(256) clahe_median_kernel(const uint8_t* __restrict__ x, int H, int W, const int* __restrict__ mm,
                                                           const uint8_t* __restrict__ lut, int tilesX, int tilesY, int tw, int th,
                                                           uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t s_v[(COND_TH + 2) * COND_PITCH];
    __shared__ int s_i1[COND_TW + 2 + COND_TH + 2], s_i2[COND_TW + 2 + COND_TH + 2], s_src[COND_TW + 2 + COND_TH + 2];
    __shared__ float s_a[COND_TW + 2 + COND_TH + 2];
    const int frame = blockIdx.z, x0 = blockIdx.x * COND_TW, y0 = blockIdx.y * COND_TH;
    const NormCoef nc = norm_coef(mm, frame);
    const uint8_t* src = x + (size_t)frame * H * W;
    const uint8_t* L = lut + (size_t)frame * tilesX * tilesY * 256;
    // per-column (first COND_TW + 2 entries) and per-row (the rest) source coordinate, LUT plane offsets and weight
    for (int i = threadIdx.x; i < COND_TW + 2 + COND_TH + 2; i += 256) {
        const bool is_col = i < COND_TW + 2;
        const int k = is_col ? i : i - (COND_TW + 2);
        const int lim = is_col ? W : H, tsz = is_col ? tw : th, ntile = is_col ? tilesX : tilesY;
        const int p = min(max((is_col ? x0 : y0) + k - 1, 0), lim - 1);                 // BORDER_REPLICATE of the median
        const float tf = __fadd_rn(__fmul_rn((float)p, 1.0f / (float)tsz), -0.5f);
        const int t1 = (int)floorf(tf);
        s_a[i] = __fadd_rn(tf, -(float)t1);
        s_i1[i] = max(t1, 0) * (is_col ? 256 : tilesX * 256);
        s_i2[i] = min(t1 + 1, ntile - 1) * (is_col ? 256 : tilesX * 256);
        s_src[i] = p;
    }
    __syncthreads();
    // a thread owns one column (its interpolation parameters live in registers) and walks every second row; the two
    // halo columns are done by the first 2 x (COND_TH + 2) threads afterwards
    auto blend = [&](int c, int r, int srcx, int i1c, int i2c, float xa, float xa1) {
        const int ri = COND_TW + 2 + r;
        const int v = norm_u8(__ldg(src + (size_t)s_src[ri] * W + srcx), nc);
        const float ya = s_a[ri], ya1 = __fadd_rn(1.0f, -ya);
        const uint8_t* L1 = L + v + s_i1[ri];
        const uint8_t* L2 = L + v + s_i2[ri];
        const float l11 = (float)__ldg(L1 + i1c), l12 = (float)__ldg(L1 + i2c), l21 = (float)__ldg(L2 + i1c), l22 = (float)__ldg(L2 + i2c);
        const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
        const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
        s_v[r * COND_PITCH + 3 + c] = (uint8_t)min(255, max(0, __float2int_rn(res)));
    };
    {
        const int c = 1 + (threadIdx.x & (COND_TW - 1));
        const int srcx = s_src[c], i1c = s_i1[c], i2c = s_i2[c];
        const float xa = s_a[c], xa1 = __fadd_rn(1.0f, -xa);
        for (int r = threadIdx.x / COND_TW; r < COND_TH + 2; r += 256 / COND_TW) blend(c, r, srcx, i1c, i2c, xa, xa1);
    }
    if (threadIdx.x < 2 * (COND_TH + 2)) {
        const int c = threadIdx.x < COND_TH + 2 ? 0 : COND_TW + 1, r = threadIdx.x < COND_TH + 2 ? threadIdx.x : threadIdx.x - (COND_TH + 2);
        blend(c, r, s_src[c], s_i1[c], s_i2[c], s_a[c], __fadd_rn(1.0f, -s_a[c]));
    }
    __syncthreads();
    const uint32_t* s_w = reinterpret_cast<const uint32_t*>(s_v);
    for (int i = threadIdx.x; i < COND_TH * (COND_TW / 4); i += 256) {
        const int r = i / (COND_TW / 4), q = i - r * (COND_TW / 4);                    // output row, word (4 pixels) in the row
        const int y = y0 + r, xq = x0 + 4 * q;
        if (y >= H || xq >= W) continue;
        uint32_t p[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t* row = s_w + ((r + k) * COND_PITCH >> 2) + q;                // word q holds bytes 4q..4q+3; pixel columns start at byte 4
            const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];
            p[3 * k + 0] = __funnelshift_r(w0, w1, 24);                                 // columns 4q-1 .. 4q+2
            p[3 * k + 1] = w1;                                                          // columns 4q   .. 4q+3
            p[3 * k + 2] = __funnelshift_r(w1, w2, 8);                                  // columns 4q+1 .. 4q+4
        }
        const uint32_t m = median9x4(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
        uint8_t* dst = out + ((size_t)frame * H + y) * W + xq;
        if (xq + 3 < W && (((uintptr_t)dst) & 3) == 0) {
            *reinterpret_cast<uint32_t*>(dst) = m;
        } else {
            for (int k = 0; k < 4 && xq + k < W; ++k) dst[k] = (uint8_t)(m >> (8 * k));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Flip test-time augmentation of the pipeline CLI (attention_aspp_unet_pipeline_stage.py:336-338, test_ablation.py:365-371):
//   prob = sigmoid((net(x) + flip(net(flip(x, W)), W)) / 2)
// `flip_w_kernel` mirrors frames along W (uint8 or fp32 elements), `tta_prob_kernel` averages the logits of the plain
// and the mirrored pass (reading the second one mirrored back) and applies the sigmoid, all in fp32.
template <typename T>
__global__ void __launch_bounds__(256) flip_w_kernel(const T* __restrict__ x, long long rows, int W, T* __restrict__ y) {
    const long long total = rows * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / W;
        const int c = (int)(i - r * W);
        y[i] = __ldg(x + r * W + (W - 1 - c));
    }
}
__global__ void __launch_bounds__(256) tta_prob_kernel(const float* __restrict__ l, const float* __restrict__ lf, long long rows, int W,
                                                       float* __restrict__ prob) {
    const long long total = rows * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / W;
        const int c = (int)(i - r * W);
        const float m = __fmul_rn(__fadd_rn(__ldg(l + i), __ldg(lf + r * W + (W - 1 - c))), 0.5f);
        prob[i] = 1.f / (1.f + expf(-m));
    }
}


// ---------------------------------------------------------------------------------------------------------
// Per-slice head and tail of the pipeline CLI around the network (attention_aspp_unet_pipeline_stage.py:492-498,
// test_ablation.py:826-834): `Resize(512, 512)` of the conditioned uint8 frame before it, and after it
//   prob = cv2.resize(prob, (W, H)); prob = cv2.GaussianBlur(prob, (5, 5), 0); mask = (prob > THR)
// -- so that only uint8 masks (and their areas) ever leave the device.
//
// resize_u8_linear_kernel reproduces cv2.resize(uint8, INTER_LINEAR) BIT-EXACTLY (OpenCV's fixed-point scheme, checked
// against cv2 4.13 for up- and down-scaling in tests/test_host_cpu.py through its numpy twin): per axis
//   f = float((d + 0.5) * (1 / (dst / src)) - 0.5), s = floor(f), f -= s,   coefficients round(f * 2048), round((1 - f) * 2048)
// columns: s < 0 or s >= src - 1 reset f to 0 (and clamp s); rows: no reset, the two source rows are clamped instead;
// horizontal pass in int32 (scale 2^11), vertical pass ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2.
__device__ __forceinline__ void cv_linear_coef(int d, double scale, int src, bool reset, int& s, float& f) {
    const float ff = (float)(((double)d + 0.5) * scale - 0.5);
    s = (int)floorf(ff);
    f = ff - (float)s;
    if (reset) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
    }
}
__global__ void __launch_bounds__(256) resize_u8_linear_kernel(const uint8_t* __restrict__ src, int SH, int SW, uint8_t* __restrict__ dst, int DH, int DW,
                                                               double scale_x, double scale_y) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (dx >= DW || dy >= DH) return;
    const uint8_t* img = src + (size_t)blockIdx.z * SH * SW;
    int sx, sy;
    float fx, fy;
    cv_linear_coef(dx, scale_x, SW, true, sx, fx);
    cv_linear_coef(dy, scale_y, SH, false, sy, fy);
    const int a1 = __float2int_rn(fx * 2048.f), a0 = __float2int_rn((1.f - fx) * 2048.f);
    const int b1 = __float2int_rn(fy * 2048.f), b0 = __float2int_rn((1.f - fy) * 2048.f);
    const int x1 = min(sx + 1, SW - 1), y0 = min(max(sy, 0), SH - 1), y1 = min(max(sy + 1, 0), SH - 1);
    const int h0 = (int)__ldg(img + (size_t)y0 * SW + sx) * a0 + (int)__ldg(img + (size_t)y0 * SW + x1) * a1;
    const int h1 = (int)__ldg(img + (size_t)y1 * SW + sx) * a0 + (int)__ldg(img + (size_t)y1 * SW + x1) * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    dst[((size_t)blockIdx.z * DH + dy) * DW + dx] = (uint8_t)min(max(v, 0), 255);
}

// tail_mask_kernel: one 32x32 output tile per block.  The (32+4)^2 bilinear samples the 5x5 blur of the tile needs are
// formed once in shared memory (cv2.resize float32 INTER_LINEAR: coefficients from the double-precision source coordinate,
// same clamping rules as above; borders of the blur are BORDER_REFLECT_101), then the separable [1 4 6 4 1] / 16 kernel
// (what cv2.GaussianBlur uses for ksize 5, sigma 0) runs as a row pass and a column pass in fp32, the result is compared
// with THR, written as uint8 {0,1} and counted (warp shuffles + one integer atomic per block).  OpenCV's own summation
// order differs in the last bit (vectorised FMA paths), so parity here is a tolerance, not bit equality: probabilities
// within 1e-6, masks identical except where a blurred probability sits within that distance of THR.
enum { TAIL_T = 32, TAIL_R = TAIL_T + 4 };
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
__global__ void __launch_bounds__(256) tail_mask_kernel(const float* __restrict__ prob, int PH, int PW, int H, int W, double scale_x, double scale_y,
                                                        float thr, uint8_t* __restrict__ mask, int* __restrict__ areas) {
    __shared__ float s_r[TAIL_R][TAIL_R + 1];
    __shared__ float s_t[TAIL_R][TAIL_T + 1];
    __shared__ int s_x0[TAIL_R], s_y0[TAIL_R];
    __shared__ float s_fx[TAIL_R], s_fy[TAIL_R];
    __shared__ int s_cnt[8];
    const int frame = blockIdx.z, x0 = blockIdx.x * TAIL_T, y0 = blockIdx.y * TAIL_T;
    const float* src = prob + (size_t)frame * PH * PW;
    if (threadIdx.x < 2 * TAIL_R) {                                  // source index / weight of every row and column of the halo tile
        const bool is_y = threadIdx.x >= TAIL_R;
        const int k = is_y ? threadIdx.x - TAIL_R : threadIdx.x;
        const int d = reflect101((is_y ? y0 : x0) - 2 + k, is_y ? H : W);
        const double c = ((double)d + 0.5) * (is_y ? scale_y : scale_x) - 0.5;
        int sidx = (int)floor(c);
        float f = (float)(c - (double)sidx);
        if (!is_y) {
            if (sidx < 0) { sidx = 0; f = 0.f; }
            if (sidx >= PW - 1) { sidx = PW - 1; f = 0.f; }
            s_x0[k] = sidx; s_fx[k] = f;
        } else {
            s_y0[k] = sidx; s_fy[k] = f;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TAIL_R * TAIL_R; i += 256) {
        const int r = i / TAIL_R, c = i - r * TAIL_R;
        const int sx = s_x0[c], sx1 = min(sx + 1, PW - 1);
        const int sy = s_y0[r], ya = min(max(sy, 0), PH - 1), yb = min(max(sy + 1, 0), PH - 1);
        const float fx = s_fx[c], fy = s_fy[r];
        const float h0 = __fadd_rn(__fmul_rn(__ldg(src + (size_t)ya * PW + sx), 1.f - fx), __fmul_rn(__ldg(src + (size_t)ya * PW + sx1), fx));
        const float h1 = __fadd_rn(__fmul_rn(__ldg(src + (size_t)yb * PW + sx), 1.f - fx), __fmul_rn(__ldg(src + (size_t)yb * PW + sx1), fx));
        s_r[r][c] = __fadd_rn(__fmul_rn(h0, 1.f - fy), __fmul_rn(h1, fy));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TAIL_R * TAIL_T; i += 256) {       // row pass
        const int r = i / TAIL_T, c = i - r * TAIL_T;
        s_t[r][c] = 0.0625f * (s_r[r][c] + s_r[r][c + 4]) + 0.25f * (s_r[r][c + 1] + s_r[r][c + 3]) + 0.375f * s_r[r][c + 2];
    }
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < TAIL_T * TAIL_T; i += 256) {       // column pass, threshold, count
        const int r = i / TAIL_T, c = i - r * TAIL_T;
        const int y = y0 + r, x = x0 + c;
        if (y < H && x < W) {
            const float v = 0.0625f * (s_t[r][c] + s_t[r + 4][c]) + 0.25f * (s_t[r + 1][c] + s_t[r + 3][c]) + 0.375f * s_t[r + 2][c];
            const int b = v > thr;
            cnt += b;
            mask[((size_t)frame * H + y) * W + x] = (uint8_t)b;
        }
    }
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < 8 ? s_cnt[threadIdx.x] : 0;
        v = warp_sum(v);
        if (threadIdx.x == 0 && v) atomicAdd(areas + frame, v);
    }
}

}  // namespace aau
