// Fused epilogue shared by the FP32 (CUDA-core) and TF32 (tcgen05) implicit-GEMM kernels:
// bias -> residual -> PReLU, output-row remapping (plain / ConvTranspose pixel-shuffle / window reverse),
// optional second PReLU'd copy.
#pragma once
#include "common.cuh"

struct EpiParams {
  int Cout;
  const float* bias;
  const float* prelu;
  const float* residual;
  int res_pitch;
  float* out;
  int out_pitch;
  float* out2;
  const float* prelu2;
  int out2_pitch;
  int out_mode;
  int Hout, Wout;          // GEMM rows enumerate (b, y, x) over this grid
  int round;               // round stored values to TF32
  atmvfi_window_geom win;
  int qkv_heads, qkv_hd, qkv_C;   // ATMVFI_OUT_QKV_HEADS: heads, head dim, C = heads * hd
  int64_t qkv_R;                  // ... rows of the GEMM = B * Hout * Wout
  // precision ATMVFI_F16: `residual` and `out2` are fp16 maps; `out` is fp16 unless out_half == 0 (q|k|v, motion heads and the final
  // residual stay fp32); `head32` receives an fp32 copy of output channels [head32_c0, Cout) (the flows + occlusion logit that the
  // warps consume at full precision while the same tensor feeds the next fp16 GEMM).  Pitches count elements of the map's type.
  int cout4;                      // Cout rounded up to 4 when whole-vector stores into the pad lanes are allowed (pad_stores), else Cout
  int act_half, out_half;
  float* head32;
  int head32_pitch, head32_c0;
};

// ATMVFI_OUT_QKV_HEADS: float offset of (GEMM row m, output column co in [0, 3C)) - see include/atmvfi.h.
// Columns of q and k are contiguous inside a head (4 consecutive columns never straddle heads: hd % 4 == 0);
// consecutive columns of v are R floats apart (V is stored transposed).
__device__ __forceinline__ int64_t epi_qkv_offset(const EpiParams& e, int64_t m, int co) {
  const int part = co / e.qkv_C, cc = co - part * e.qkv_C;
  if (part < 2) {
    const int h = cc / e.qkv_hd, d = cc - h * e.qkv_hd;
    return ((int64_t)(part * e.qkv_heads + h) * e.qkv_R + m) * e.qkv_hd + d;
  }
  return 2 * (int64_t)e.qkv_C * e.qkv_R + (int64_t)cc * e.qkv_R + m;
}

// Destination row of GEMM row m (and shuffle block q); -1 = discard (centre-padding token).
__device__ __forceinline__ int64_t epi_out_row(const EpiParams& e, int64_t m, int q) {
  if (e.out_mode == ATMVFI_OUT_PIXEL || e.out_mode == ATMVFI_OUT_QKV_HEADS) return m;
  if (e.out_mode == ATMVFI_OUT_SHUFFLE2) {
    int x = (int)(m % e.Wout);
    int64_t t = m / e.Wout;
    int y = (int)(t % e.Hout);
    int64_t b = t / e.Hout;
    return (b * 2 * e.Hout + 2 * y + (q >> 1)) * (2 * (int64_t)e.Wout) + 2 * x + (q & 1);
  }
  WinPos p = win_decode(e.win, m);
  if (!p.real) return -1;
  return ((int64_t)p.b * e.win.H + p.y) * e.win.W + p.x;
}

// acc = A*W for (GEMM row m, output channel co); orow from epi_out_row.
__device__ __forceinline__ void epi_store(const EpiParams& e, int64_t m, int64_t orow, int co, float acc) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + co);
  if (e.residual) v += __ldg(e.residual + m * e.res_pitch + co);
  if (e.prelu) v = v > 0.f ? v : v * __ldg(e.prelu + co);
  if (e.out_mode == ATMVFI_OUT_QKV_HEADS) {
    e.out[epi_qkv_offset(e, m, co)] = round_tf32_if(v, e.round);
    return;
  }
  e.out[orow * e.out_pitch + co] = round_tf32_if(v, e.round);
  if (e.out2) e.out2[orow * e.out2_pitch + co] = round_tf32_if(v > 0.f ? v : v * __ldg(e.prelu2 + co), e.round);
}

static inline EpiParams make_epi(const atmvfi_gemm_conv_desc* d) {
  EpiParams e;
  e.Cout = d->Cout; e.bias = d->bias; e.prelu = d->prelu; e.residual = d->residual; e.res_pitch = d->res_pitch;
  e.out = d->out; e.out_pitch = d->out_pitch; e.out2 = d->out2; e.prelu2 = d->prelu2; e.out2_pitch = d->out2_pitch;
  e.out_mode = d->out_mode; e.Hout = d->Hout; e.Wout = d->Wout; e.win = d->win; e.round = atmvfi_output_rounding();
  e.qkv_heads = d->qkv_heads > 0 ? d->qkv_heads : 1;
  e.qkv_C = d->Cout / 3;
  e.qkv_hd = e.qkv_C / e.qkv_heads;
  e.qkv_R = (int64_t)d->B * d->Hout * d->Wout;
  e.cout4 = d->pad_stores ? (d->Cout + 3) / 4 * 4 : d->Cout;
  e.act_half = d->precision == ATMVFI_F16 ? 1 : 0;
  e.out_half = (e.act_half && !d->out_f32) ? 1 : 0;
  e.head32 = e.act_half ? d->head32 : nullptr;
  e.head32_pitch = d->head32_pitch; e.head32_c0 = d->head32_c0;
  return e;
}
