// Prints the occupancy the runtime reports for igemm_tc_kernel at several dynamic shared-memory sizes.
#include "../att-aspp-unet_b200/csrc/igemm_tc.cuh"
#include <cstdio>
int main() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, (aau::igemm_tc_kernel<2, false, false>));
    printf("regs %d static smem %zu maxDyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
    cudaError_t e1 = cudaFuncSetAttribute((aau::igemm_tc_kernel<2, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 4096);
    cudaError_t e2 = cudaFuncSetAttribute((aau::igemm_tc_kernel<2, false, false>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    printf("set attr: %s / %s\n", cudaGetErrorString(e1), cudaGetErrorString(e2));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("smemPerSM %zu smemPerBlockOptin %zu reserved %zu regsPerSM %d\n", p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.reservedSharedMemPerBlock, p.regsPerMultiprocessor);
    for (size_t s : {16384, 32768, 49152, 65536, 71680, 81920, 110000, 200000}) {
        int occ = -1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (aau::igemm_tc_kernel<2, false, false>), 192, s);
        printf("dyn %zu -> occ %d (%s)\n", s, occ, cudaGetErrorString(e));
    }
    return 0;
}
