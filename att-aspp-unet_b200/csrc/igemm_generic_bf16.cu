// Generic igemm_tc_kernel instantiations, bf16 storage (see igemm_inst.cuh).
#include "igemm_tc.cuh"
#include "igemm_inst.cuh"
AAU_IGEMM_DEFINE_TABLE(AAU_IGEMM_GENERIC, false, igemm_generic_bf16_raise)
namespace aau {
const void* igemm_generic_bf16(int ng, bool multi, bool pair) { return table_lookup(ng, multi, pair, -1, -1, -1, -1); }
}
