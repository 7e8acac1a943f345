// fp16 (tcgen05 kind::f16) instantiations of the implicit-GEMM kernel of gemm_conv_tc.cu, compiled as their own translation unit.
#define ATMVFI_TC_F16_TU 1
#include "gemm_conv_tc.cu"
