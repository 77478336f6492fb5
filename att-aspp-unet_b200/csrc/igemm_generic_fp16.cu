// Generic igemm_tc_kernel instantiations, fp16 storage (see igemm_inst.cuh).
#include "igemm_tc.cuh"
#include "igemm_inst.cuh"
AAU_IGEMM_DEFINE_TABLE(AAU_IGEMM_GENERIC, true, igemm_generic_fp16_raise)
namespace aau {
const void* igemm_generic_fp16(int ng, bool multi, bool pair) { return table_lookup(ng, multi, pair, -1, -1, -1, -1); }
}
