// C-ABI plumbing of libatmvfi_b200.so: error reporting, device probe, precision dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";
static thread_local int g_round = 0;
static thread_local int g_act_f16 = 0;
int atmvfi_output_rounding() { return g_round; }
int atmvfi_act_f16() { return g_act_f16; }
int atmvfi_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* ev = getenv("ATMVFI_PDL"); on = ev ? (atoi(ev) != 0) : 1; }
  return on;
}

void atmvfi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int atmvfi_gemm_conv_simt(const atmvfi_gemm_conv_desc* d, cudaStream_t st);
int atmvfi_gemm_conv_tc(const atmvfi_gemm_conv_desc* d, cudaStream_t st);

extern "C" {

const char* atmvfi_last_error(void) { return g_err; }
int atmvfi_abi_version(void) { return ATMVFI_ABI_VERSION; }
void atmvfi_set_output_rounding(int on) { g_round = on ? 1 : 0; }
void atmvfi_set_activation_f16(int on) { g_act_f16 = on ? 1 : 0; }

int atmvfi_device_info(int device, char* name, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    atmvfi_set_error("device_info: %s", cudaGetErrorString(e));
    return -1;
  }
  if (name) { strncpy(name, prop.name, 255); name[255] = 0; }
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (prop.major != 10) {
    atmvfi_set_error("device_info: %s is sm_%d%d; this library only carries sm_100a code", prop.name, prop.major, prop.minor);
    return -1;
  }
  return prop.multiProcessorCount;
}

int atmvfi_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    atmvfi_set_error("copy: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int atmvfi_gemm_conv(const atmvfi_gemm_conv_desc* d, void* stream) {
  ATMVFI_REQUIRE(d != nullptr, "gemm_conv: null descriptor");
  ATMVFI_REQUIRE(d->nsrc >= 1 && d->nsrc <= ATMVFI_MAX_SRC, "gemm_conv: nsrc=%d out of range", d->nsrc);
  ATMVFI_REQUIRE(d->ksize == 1 || d->ksize == 3, "gemm_conv: kernel size %d unsupported (1 or 3)", d->ksize);
  ATMVFI_REQUIRE(d->stride >= 1 && d->dil >= 1, "gemm_conv: bad stride/dilation");
  ATMVFI_REQUIRE(d->out_mode >= ATMVFI_OUT_PIXEL && d->out_mode <= ATMVFI_OUT_QKV_HEADS, "gemm_conv: bad out_mode %d", d->out_mode);
  ATMVFI_REQUIRE(d->out_mode != ATMVFI_OUT_QKV_HEADS ||
                     (d->ksize == 1 && d->stride == 1 && !d->residual && !d->out2 && d->qkv_heads > 0 && d->Cout % (12 * d->qkv_heads) == 0),
                 "gemm_conv: QKV_HEADS needs a 1x1 layer without residual / second output and Cout = 3 * heads * hd with hd %% 4 == 0");
  ATMVFI_REQUIRE(d->out_mode != ATMVFI_OUT_SHUFFLE2 || (d->ksize == 1 && d->stride == 1), "gemm_conv: SHUFFLE2 needs ksize=1, stride=1");
  ATMVFI_REQUIRE(!d->out2 || d->prelu2, "gemm_conv: out2 needs prelu2 slopes");
  if (d->precision == ATMVFI_FP32) return atmvfi_gemm_conv_simt(d, (cudaStream_t)stream);
  if (d->precision == ATMVFI_TF32 || d->precision == ATMVFI_TF32X3 || d->precision == ATMVFI_F16) return atmvfi_gemm_conv_tc(d, (cudaStream_t)stream);
  atmvfi_set_error("gemm_conv: unknown precision %d", d->precision);
  return 2;
}

}  // extern "C"
