// Does cuTensorMapEncodeTiled accept zero / overlapping strides (a 2x2 pooling window folded onto one destination pixel)?
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
int main() {
    cudaFree(0);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    void* p; cudaMalloc(&p, 1 << 24);
    CUtensorMap tm;
    const int C = 32, W2 = 100, H2 = 80;
    for (int variant = 0; variant < 3; ++variant) {
        // dims: (C, xi=2, xo=W2, yi=2, yo=H2); pooled tensor [H2][W2][C] bf16
        cuuint64_t dims[5] = {C, 2, W2, 2, H2};
        cuuint64_t s_xi = variant == 0 ? 0 : (variant == 1 ? 16 : C * 2), s_yi = variant == 0 ? 0 : (variant == 1 ? 16 : (cuuint64_t)W2 * C * 2);
        cuuint64_t strides[4] = {s_xi, (cuuint64_t)C * 2, s_yi, (cuuint64_t)W2 * C * 2};
        cuuint32_t box[5] = {32, 2, 15, 2, 2}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const char* name = nullptr; cuGetErrorName(r, &name);
        printf("variant %d (stride xi %llu, yi %llu): %s\n", variant, (unsigned long long)s_xi, (unsigned long long)s_yi, name ? name : "?");
    }
    return 0;
}
