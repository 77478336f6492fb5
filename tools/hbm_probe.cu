// HBM direction probe: pure read, pure write and copy bandwidth of plain 16-byte-per-thread streaming kernels
// (results: profiles/r01_hbm_probe.txt).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_probe tools/hbm_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(uint4* p, size_t n) {
    const uint4 v = make_uint4(threadIdx.x, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_read(const uint4* p, size_t n, unsigned* sink) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const uint4 v = __ldg(p + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void k_copy(const uint4* a, uint4* b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = __ldg(a + i);
}
// 2 reads : 1 write (the attention gate's mix) and 1 read : 2 writes (the transposed conv's mix)
__global__ void k_r2w1(const uint4* a, const uint4* b, uint4* c, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 x = __ldg(a + i), y = __ldg(b + i); x.x ^= y.x; c[i] = x; }
}
__global__ void k_r1w2(const uint4* a, uint4* b, uint4* c, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const uint4 x = __ldg(a + i); b[i] = x; c[i] = x; }
}
int main() {
    const size_t bytes = 1ull << 30, n = bytes / 16;
    uint4 *a, *b, *c; unsigned* sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&c, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes); cudaMemset(c, 3, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148 * 8, 148 * 16, 148 * 32}) {
        for (int mode = 0; mode < 5; ++mode) {
            float best = 1e9f;
            for (int it = 0; it < 6; ++it) {
                cudaEventRecord(e0);
                if (mode == 0) k_write<<<grid, 256>>>(b, n);
                if (mode == 1) k_read<<<grid, 256>>>(a, n, sink);
                if (mode == 2) k_copy<<<grid, 256>>>(a, b, n);
                if (mode == 3) k_r2w1<<<grid, 256>>>(a, b, c, n);
                if (mode == 4) k_r1w2<<<grid, 256>>>(a, b, c, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
            }
            const double moved = bytes * (mode == 0 || mode == 1 ? 1.0 : (mode == 2 ? 2.0 : 3.0));
            const char* nm[] = {"write only", "read only", "copy 1r:1w", "2 reads : 1 write", "1 read : 2 writes"};
            printf("grid %5d  %-18s %7.3f ms  %6.0f GB/s total\n", grid, nm[mode], best, moved / best / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
