// Micro-probes behind the igemm design decisions (results: profiles/r01_mma_probe.txt).
//   1. cost of back-to-back tcgen05.mma (M=128, K=16) issued by one thread, as a function of N and of CTAs per SM
//   2. A descriptors that start on a row that is not a multiple of 8 (base-offset field), needed for "row shift" taps
//   3. tcgen05.ld throughput (TMEM -> registers) with 4 and 8 warps
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu
#include "../att-aspp-unet_b200/csrc/ptx_sm100.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>

using namespace aau;

struct ProbeOut { long long issue, total; };

// ---- 1. issue / completion cost of `reps` MMAs of width N
// MODE 0: `if (threadIdx.x == 0)` single-thread role (divergent for the compiler: every UTCHMMA sits in a waterfall loop)
// MODE 1: warp-uniform role (warp index broadcast with a shuffle), MMA / commit under elect.sync
template <int MODE, int KC>
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int reps, ProbeOut* out, int a_row_shift = 0, int swz = 128) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar2[8];
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    for (int i = threadIdx.x; i < (20480 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) ptx::mbar_init(&bar2[i], 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_base_s, 256); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // warp-uniform for the compiler
    if (MODE == 1 ? (warp_u == 0) : (threadIdx.x == 0)) {
        const uint32_t tm = tmem_base_s;
        const uint32_t idesc = ptx::make_idesc_f16(128, N, false);
        const uint64_t da = ptx::make_kmajor_desc(ptx::smem_u32(smem) + (uint32_t)(a_row_shift * swz), swz);
        const uint64_t db = ptx::make_kmajor_desc(ptx::smem_u32(smem + 20480), swz);
        int nc = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
            if (MODE == 0 || ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tm, da + 2 * k, db + 2 * k, idesc, (r | k) > 0);
                if (KC) ptx::umma_commit(&bar2[nc & 7]);          // pipeline-style commit per 4-MMA sub-block (nobody waits)
            }
            if (MODE == 1) __syncwarp();
            ++nc;
        }
        const long long t_issue = clock64() - t0;
        if (MODE == 0 || ptx::elect_one()) ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0, nullptr, 0);
        const long long t1 = clock64() - t0;
        if (blockIdx.x == 0 && threadIdx.x == 0) { out->issue = t_issue; out->total = t1; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base_s, 256); }
}

// ---- 2. row-shifted A descriptor: D = A[row0 + r] . I  (B = 64x64 identity rows, N = 64, K = 64)
__global__ void __launch_bounds__(128) rowshift_kernel(int row0, int use_base_offset, float* dout /* [128][64] */) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                   // 160 rows x 128 B, 128B-swizzled exactly as TMA would write them
    uint8_t* sb = smem + 160 * 128;       // 64 rows x 128 B (1024-aligned: 160*128 = 20480)
    for (int i = threadIdx.x; i < 160 * 64; i += blockDim.x) {
        const int r = i / 64, k = i % 64;
        const float v = (float)(r * 2 + (k == (r % 64) ? 1 : 0));       // A[r][k] = 2r (+1 on a diagonal): exact in bf16 for r < 128
        uint32_t off = (uint32_t)(r * 128 + k * 2);
        off ^= ((off >> 7) & 7) << 4;
        *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16(v);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        uint32_t off = (uint32_t)(n * 128 + k * 2);
        off ^= ((off >> 7) & 7) << 4;
        *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16(n == k ? 1.f : 0.f);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_base_s, 64); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::make_idesc_f16(128, 64, false);
        const uint32_t a_addr = ptx::smem_u32(sa) + (uint32_t)row0 * 128u;
        uint64_t da = ptx::make_kmajor_desc(a_addr, 128);
        if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
        const uint64_t db = ptx::make_kmajor_desc(ptx::smem_u32(sb), 128);
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tm, da + 2 * k, db + 2 * k, idesc, k > 0);
        ptx::umma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0, nullptr, 0);
    ptx::tc_fence_after();
    const int warp = threadIdx.x >> 5;
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        ptx::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) dout[threadIdx.x * 64 + c0 + i] = __uint_as_float(r[i]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 64); }
}

// ---- 3. tcgen05.ld throughput: every warp reads its lane quarter, `cols` columns, `reps` times
__global__ void __launch_bounds__(256) tmem_ld_kernel(int cols, int reps, ProbeOut* out, float* sink) {
    __shared__ uint32_t tmem_base_s;
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_base_s, 256); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = tmem_base_s + ((uint32_t)((threadIdx.x >> 5) & 3) * 32 << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (int c = 0; c < cols; c += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tm + c, v);
            ptx::tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        }
    __syncthreads();
    const long long t1 = clock64() - t0;
    if (threadIdx.x == 0 && blockIdx.x == 0) { out->issue = t1; out->total = t1; }
    if (acc == 12345.f) sink[threadIdx.x] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base_s, 256); }
}

// ---- 4. CTA pair: tcgen05.mma.cta_group::2, M = 256 (128 rows per CTA), N = 64 with the B rows split across the pair
//  A_r[m][k] (CTA r, 128 x 64), B[n][k] (64 x 64; CTA r holds rows n in [32r, 32r+32)), D = A . B^T
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_mma_kernel(float* dout /* [256][64] */, int* status) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t rank = cluster_ctarank();
    uint8_t* sa = smem;                   // 128 rows x 128 B (K = 64 bf16), 128B swizzle
    uint8_t* sb = smem + 16384;           // 32 rows x 128 B: this CTA's half of B
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int m = i / 64, k = i % 64;
        const float v = (float)(((m + 128 * (int)rank) % 7) - 3) + (k == (m % 64) ? 0.5f : 0.f);   // small integers / halves: exact in bf16
        uint32_t off = (uint32_t)(m * 128 + k * 2);
        off ^= ((off >> 7) & 7) << 4;
        *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16(v);
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
        const int nl = i / 64, k = i % 64, n = nl + 32 * (int)rank;
        const float v = (float)(((n * 3 + k) % 5) - 2);
        uint32_t off = (uint32_t)(nl * 128 + k * 2);
        off ^= ((off >> 7) & 7) << 4;
        *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16(v);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // both CTAs have their operands and barriers ready
    ptx::tc_fence_after();
    const uint32_t tm = tmem_base_s;
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (rank == 0 && warp_u == 0) {
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_f16(256, 64, false);
            const uint64_t da = ptx::make_kmajor_desc(ptx::smem_u32(sa), 128), db = ptx::make_kmajor_desc(ptx::smem_u32(sb), 128);
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = k > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                             ::"r"(tm), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(ptx::smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
    }
    uint32_t spins = 0;
    while (!ptx::mbar_try_wait(&bar, 0)) {
        if (++spins > (1u << 20)) { if (threadIdx.x == 0) atomicExch(status, 100 + (int)rank); break; }
    }
    ptx::tc_fence_after();
    const int warp = threadIdx.x >> 5;
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        ptx::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) dout[(size_t)(128 * rank + threadIdx.x) * 64 + c0 + i] = __uint_as_float(r[i]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u) : "memory");
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
    ProbeOut* d;
    CK(cudaMalloc(&d, sizeof(ProbeOut)));
    float* sink;
    CK(cudaMalloc(&sink, 4096));
    CK(cudaFuncSetAttribute(rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int reps = 512;
    printf("# 1. back-to-back tcgen05.mma M=128 K=16, %d MMAs in groups of 4; cycles per MMA (issue loop / until the final commit arrives)\n", reps);
    auto run1 = [&](auto kern, const char* what) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        for (int ctas : {1, 2})
            for (int N : {16, 32, 64, 96, 128, 256}) {
                ProbeOut h;
                for (int it = 0; it < 2; ++it) {
                    kern<<<148 * ctas, 128, 60 * 1024>>>(N, reps, d, 0, 128);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
                printf("%s ctas/SM %d N=%3d : issue %.1f  total %.1f  (nominal N/2 = %d)\n", what, ctas, N, (double)h.issue / reps, (double)h.total / reps, N / 2);
            }
    };
    run1(mma_rate_kernel<0, 0>, "if(thread==0), no commits      ");
    run1(mma_rate_kernel<0, 1>, "if(thread==0), commit per 4    ");
    run1(mma_rate_kernel<1, 0>, "warp-uniform+elect, no commits ");
    run1(mma_rate_kernel<1, 1>, "warp-uniform+elect, commit per 4");
    printf("# 1b. the same with the A descriptor starting `shift` pixel rows into the tile (row-shifted taps), 1 CTA/SM\n");
    for (int swz : {128, 64})
        for (int shift : {0, 1, 2, 8, 33})
            for (int N : {32, 64, 96}) {
                ProbeOut h;
                for (int it = 0; it < 2; ++it) {
                    mma_rate_kernel<1, 1><<<148, 128, 60 * 1024>>>(N, reps, d, shift, swz);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
                printf("swizzle %3d shift %2d N=%3d : total %.1f cycles per MMA\n", swz, shift, N, (double)h.total / reps);
            }
    printf("# 2. row-shifted A descriptor (start row not a multiple of 8)\n");
    float* dout;
    CK(cudaMalloc(&dout, 128 * 64 * 4));
    std::vector<float> ho(128 * 64);
    for (int row0 : {0, 1, 3, 8, 9}) {
        for (int ubo : {0, 1}) {
            rowshift_kernel<<<1, 128, 64 * 1024>>>(row0, ubo, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("row0 %d base_offset %d: CUDA error %s\n", row0, ubo, cudaGetErrorString(e)); return 0; }
            CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int k = 0; k < 64; ++k) {
                    const int ar = r + row0;
                    const float want = (float)(ar * 2 + (k == (ar % 64) ? 1 : 0));
                    if (ho[r * 64 + k] != want) ++bad;
                }
            printf("row0 %d base_offset_field %d : %d wrong of 8192 (D[0][0..2] = %g %g %g, D[1][0..2] = %g %g %g)\n", row0, ubo, bad, ho[0], ho[1], ho[2], ho[64], ho[65], ho[66]);
        }
    }
    printf("# 4. CTA pair MMA (cta_group::2, M=256, N=64, B rows split across the pair)\n");
    {
        CK(cudaFuncSetAttribute(pair_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        float* dd; int* st;
        CK(cudaMalloc(&dd, 256 * 64 * 4)); CK(cudaMalloc(&st, 4)); CK(cudaMemset(st, 0, 4)); CK(cudaMemset(dd, 0xff, 256 * 64 * 4));
        pair_mma_kernel<<<2, 128, 40 * 1024>>>(dd, st);
        cudaError_t e = cudaDeviceSynchronize();
        printf("launch: %s\n", cudaGetErrorString(e));
        if (e == cudaSuccess) {
            std::vector<float> hd(256 * 64); int hs = 0;
            CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost));
            int bad = 0, bad_hi = 0;
            for (int m = 0; m < 256; ++m)
                for (int n = 0; n < 64; ++n) {
                    float want = 0.f;
                    for (int k = 0; k < 64; ++k) {
                        const float a = (float)(((m) % 7) - 3) + (k == ((m % 128) % 64) ? 0.5f : 0.f);
                        const float b = (float)(((n * 3 + k) % 5) - 2);
                        want += a * b;
                    }
                    if (hd[m * 64 + n] != want) { ++bad; if (m >= 128) ++bad_hi; }
                }
            int colbad[8] = {0}, rowbad[8] = {0};
            for (int m = 0; m < 256; ++m)
                for (int n = 0; n < 64; ++n) {
                    float want = 0.f;
                    for (int k = 0; k < 64; ++k) want += ((float)((m % 7) - 3) + (k == ((m % 128) % 64) ? 0.5f : 0.f)) * (float)(((n * 3 + k) % 5) - 2);
                    if (hd[m * 64 + n] != want) { ++colbad[n / 8]; ++rowbad[m / 32]; }
                }
            printf("wrong per 8-column group:"); for (int i = 0; i < 8; ++i) printf(" %d", colbad[i]);
            printf("\nwrong per 32-row group:"); for (int i = 0; i < 8; ++i) printf(" %d", rowbad[i]);
            printf("\nrow 5: "); for (int n = 0; n < 64; ++n) printf("%g ", hd[5 * 64 + n]); printf("\n");
            printf("status %d, wrong %d of 16384 (%d in the second CTA's rows); D[0][0..3] = %g %g %g %g, D[128][0..3] = %g %g %g %g\n", hs, bad, bad_hi,
                   hd[0], hd[1], hd[2], hd[3], hd[128 * 64], hd[128 * 64 + 1], hd[128 * 64 + 2], hd[128 * 64 + 3]);
        } else return 0;
    }
    printf("# 3. tcgen05.ld 32x32b.x32 throughput, one CTA per SM\n");
    for (int threads : {128, 256}) {
        for (int cols : {32, 96, 256}) {
            ProbeOut h;
            const int r2 = 64;
            for (int it = 0; it < 2; ++it) {
                tmem_ld_kernel<<<148, threads>>>(cols, r2, d, sink);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
            const double bytes = (double)r2 * cols * 4 * threads;
            printf("threads %d cols %3d : %.1f cycles per pass, %.1f B/clk/SM\n", threads, cols, (double)h.total / r2, bytes / (double)h.total);
        }
    }
    return 0;
}
