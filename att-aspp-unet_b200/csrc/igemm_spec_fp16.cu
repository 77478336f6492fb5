// igemm_tc_kernel instantiations specialised per (staging mode, epilogue, K step, fused pool), fp16 storage (see igemm_inst.cuh).
#include "igemm_tc.cuh"
#include "igemm_inst.cuh"
AAU_IGEMM_DEFINE_TABLE(AAU_IGEMM_SPECIALISED, true, igemm_spec_fp16_raise)
namespace aau {
const void* igemm_spec_fp16(int ng, bool multi, bool pair, int am, int ep, int kk, int pl) { return table_lookup(ng, multi, pair, am, ep, kk, pl); }
}
