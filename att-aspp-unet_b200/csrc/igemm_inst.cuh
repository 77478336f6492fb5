// Instantiation tables of igemm_tc_kernel, split over several translation units so that nvcc compiles them in parallel
// (igemm_generic_{bf16,fp16}.cu, igemm_spec_{bf16,fp16}.cu).  The engine looks kernels up through the functions below.
//
// The generic instantiations (AM = EP = KK = PL = -1) branch on the staging mode and epilogue at run time and serve every
// plan; the hot path additionally gets instantiations with all four fixed at compile time -- a third of the code size
// each, so the single-warp roles miss the instruction cache less and skip the uniform mode branches -- for BOTH 16-bit
// storage types (round 1 had them for bf16 only; fp16 storage is the headline type since round 2: DESIGN.md section 2).
#pragma once

// (NG, fp16, multi-problem, CTA pair, staging mode, epilogue, MMAs per sub-block = KC / 16, fused MaxPool)
#define AAU_IGEMM_GENERIC(X, F) \
    X(2, F, false, false, -1, -1, -1, -1) X(4, F, false, false, -1, -1, -1, -1) X(2, F, true, false, -1, -1, -1, -1) \
    X(2, F, false, true, -1, -1, -1, -1) X(2, F, true, true, -1, -1, -1, -1) X(4, F, false, true, -1, -1, -1, -1)

#define AAU_IGEMM_SPECIALISED(X, F) \
    X(2, F, false, false, AMODE_RS, EPI_STORE, 2, 1) X(2, F, false, false, AMODE_RS, EPI_STORE, 2, 0) X(4, F, false, false, AMODE_RS, EPI_STORE, 4, 0) \
    X(2, F, false, false, AMODE_RS, EPI_OUTCONV, 2, 0) \
    X(2, F, false, false, AMODE_DXN, EPI_STORE, 4, 0) X(2, F, false, false, AMODE_DXN, EPI_STORE, 4, 1) X(2, F, false, false, AMODE_DXN, EPI_STORE, 2, 0) \
    X(2, F, false, true, AMODE_SLAB, EPI_STORE, 4, 0) X(2, F, false, true, AMODE_SLAB, EPI_STORE, 4, 1) \
    X(2, F, false, false, AMODE_TAP, EPI_GATE, 4, 0) X(2, F, false, false, AMODE_TAP, EPI_CONVT, 4, 0) X(4, F, false, false, AMODE_TAP, EPI_CONVT, 4, 0) \
    X(2, F, false, false, AMODE_TAP, EPI_CONVTFIX, 4, 0) X(2, F, false, true, AMODE_TAP, EPI_STORE, 4, 0) X(2, F, true, true, AMODE_TAP, EPI_STORE, 4, 0) \
    X(2, F, false, true, AMODE_RS, EPI_STORE, 2, 1) X(2, F, false, true, AMODE_RS, EPI_STORE, 2, 0) X(4, F, false, true, AMODE_RS, EPI_STORE, 4, 0) X(4, F, false, true, AMODE_RS, EPI_STORE, 4, 1) \
    X(2, F, false, true, AMODE_RS, EPI_OUTCONV, 2, 0) \
    X(2, F, false, true, AMODE_DXN, EPI_STORE, 4, 0) X(2, F, false, true, AMODE_DXN, EPI_STORE, 4, 1) X(2, F, false, true, AMODE_DXN, EPI_STORE, 2, 0) \
    X(2, F, false, true, AMODE_TAP, EPI_CONVT, 4, 0) X(4, F, false, true, AMODE_TAP, EPI_CONVT, 4, 0) \
    X(2, F, false, true, AMODE_DXN, EPI_OUTCONV, 2, 0) X(2, F, false, true, AMODE_DXN, EPI_STORE, 2, 1)

namespace aau {
// kernel lookup: nullptr when the table of that translation unit has no such instantiation
const void* igemm_generic_bf16(int ng, bool multi, bool pair);
const void* igemm_generic_fp16(int ng, bool multi, bool pair);
const void* igemm_spec_bf16(int ng, bool multi, bool pair, int am, int ep, int kk, int pl);
const void* igemm_spec_fp16(int ng, bool multi, bool pair, int am, int ep, int kk, int pl);
// raise the dynamic shared-memory limit of every instantiation of the table (once per process and device)
bool igemm_generic_bf16_raise(int bytes);
bool igemm_generic_fp16_raise(int bytes);
bool igemm_spec_bf16_raise(int bytes);
bool igemm_spec_fp16_raise(int bytes);
}  // namespace aau

// body of one table translation unit
#define AAU_IGEMM_DEFINE_TABLE(LIST, F, RAISE_NAME)                                                            \
    namespace aau {                                                                                                                   \
    static const void* table_lookup(int ng, bool multi, bool pair, int am, int ep, int kk, int pl) {                                  \
        LIST(AAU_IGEMM_X_LOOKUP, F)                                                                                                   \
        return nullptr;                                                                                                               \
    }                                                                                                                                 \
    bool RAISE_NAME(int bytes) {                                                                                                      \
        bool ok = true;                                                                                                               \
        LIST(AAU_IGEMM_X_RAISE, F)                                                                                                    \
        return ok;                                                                                                                    \
    }                                                                                                                                 \
    }
#define AAU_IGEMM_X_LOOKUP(NG, F16, MULTI, PAIR, AM, EP, KK, PL)                                                       \
    if (ng == NG && multi == MULTI && pair == PAIR && am == (AM) && ep == (EP) && kk == (KK) && pl == (PL))            \
        return (const void*)igemm_tc_kernel<NG, F16, MULTI, PAIR, AM, EP, KK, PL>;
#define AAU_IGEMM_X_RAISE(NG, F16, MULTI, PAIR, AM, EP, KK, PL)                                                                                  \
    ok = ok && cudaFuncSetAttribute((const void*)igemm_tc_kernel<NG, F16, MULTI, PAIR, AM, EP, KK, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess && \
         cudaFuncSetAttribute((const void*)igemm_tc_kernel<NG, F16, MULTI, PAIR, AM, EP, KK, PL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess;
