// placeholder until the tcgen05 kernel lands (next commit)
#include "gemm_epilogue.cuh"
int atmvfi_gemm_conv_tc(const atmvfi_gemm_conv_desc*, cudaStream_t) {
  atmvfi_set_error("gemm_conv(tf32): not built yet");
  return 3;
}
extern "C" int atmvfi_gemm_conv_plan_bytes(void) { return 0; }
extern "C" int atmvfi_gemm_conv_plan(const atmvfi_gemm_conv_desc*, void*) { atmvfi_set_error("gemm_conv_plan: not built yet"); return 3; }
