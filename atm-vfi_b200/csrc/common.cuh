// Shared helpers for the sm_100a kernels of libatmvfi_b200.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/atmvfi.h"

void atmvfi_set_error(const char* fmt, ...);
// 1: producers of channels-last feature maps round their outputs to TF32 (cvt.rna), because tcgen05 kind::tf32
// TRUNCATES the low 13 mantissa bits of its operands (measured: -2.8e-4 relative magnitude bias per layer).
int atmvfi_output_rounding();

// 1: channels-last feature maps are stored as fp16 (precision ATMVFI_F16): the entry points that read / write such maps pick their
// __half instantiation.  Planar images, flows, masks and the 5-channel motion heads stay fp32 in every mode.
int atmvfi_act_f16();

#define ATMVFI_CHECK_LAUNCH(what)                                                        \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      atmvfi_set_error("%s: launch failed: %s", what, cudaGetErrorString(e__));          \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define ATMVFI_REQUIRE(cond, ...)                                                        \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      atmvfi_set_error(__VA_ARGS__);                                                     \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

constexpr int ATMVFI_MAX_DEVICES = 64;      // per-device launch configuration caches are indexed by the CUDA device ordinal


// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (ATMVFI_PDL, default on): the tensor-core kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so their CTAs may become resident - and run their prologue (barrier init,
// TMEM allocation, tensor-map fetch) - while the previous kernel of the stream drains its last tiles.  pdl_wait() is the point
// before which such a kernel touches no global memory: it returns once the previous kernel has completed and its writes are
// visible.  The persistent kernels (one CTA per SM, all resident) call pdl_launch_dependents() at their start, which lets the
// NEXT kernel's CTAs take over SMs as this kernel's CTAs exit; it does not weaken that kernel's own pdl_wait().  Both are no-ops
// in a launch without the attribute.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int atmvfi_pdl_enabled();          // api.cu: ATMVFI_PDL environment switch (default 1)

// <<<grid, block, 0, stream>>> with the programmatic-stream-serialization attribute (when ATMVFI_PDL is on): for the small kernels
// that sit between two tensor-core launches of a transformer block.  The kernel must call pdl_wait() before its first global access.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = atmvfi_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Row window [y0, y1) of every image of a [B][H][W] grid (include/atmvfi.h "ROW WINDOWS"); y1 == 0 means all rows.
// Host: normalise to (y0, ny).  Returns false when the window is malformed.
static inline bool row_window(int H, int y0, int y1, int* o_y0, int* o_ny) {
  if (y1 == 0 && y0 == 0) { *o_y0 = 0; *o_ny = H; return true; }
  if (y0 < 0 || y1 > H || y1 < y0) return false;
  *o_y0 = y0; *o_ny = y1 - y0;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Window bookkeeping (attention.py:8-71, 275-331) as index arithmetic.
// A window-major row r enumerates (image b, window row wy, window col wx, token ty, tx) exactly like
// window_partition() does on the centre-padded, cyclically shifted map.
// ---------------------------------------------------------------------------------------------
struct WinPos {
  int b;        // image on the batch axis
  int yr, xr;   // coordinates in the padded + rolled frame (what the window sees)
  int y, x;     // coordinates in the un-padded token grid, valid only if `real`
  bool real;    // false: a centre-padding token
};

__device__ __forceinline__ WinPos win_decode(const atmvfi_window_geom& g, int64_t r) {
  const int N = g.ws * g.ws;
  const int nwx = g.Wp / g.ws, nwy = g.Hp / g.ws;
  int n = (int)(r % N);
  int64_t w = r / N;
  int wx = (int)(w % nwx);
  w /= nwx;
  int wy = (int)(w % nwy);
  WinPos p;
  p.b = (int)(w / nwy);
  p.yr = wy * g.ws + n / g.ws;
  p.xr = wx * g.ws + n % g.ws;
  // torch.roll(x, -shift): rolled[i] = padded[(i + shift) mod size]
  int yp = p.yr + g.shift;
  if (yp >= g.Hp) yp -= g.Hp;
  int xp = p.xr + g.shift;
  if (xp >= g.Wp) xp -= g.Wp;
  p.y = yp - g.pad_top;
  p.x = xp - g.pad_left;
  p.real = (p.y >= 0) && (p.y < g.H) && (p.x >= 0) && (p.x < g.W);
  return p;
}

// 32-bit flavour for the kernels that decode one row per warp (64-bit divisions cost ~100 instructions each, as much as the
// LayerNorm of the row itself); the caller guarantees r < 2^31.
__device__ __forceinline__ WinPos win_decode32(const atmvfi_window_geom& g, int r) {
  const int N = g.ws * g.ws;
  const int nwx = g.Wp / g.ws, nwy = g.Hp / g.ws;
  int w = r / N;
  const int n = r - w * N;
  const int ty = n / g.ws, tx = n - ty * g.ws;
  int q = w / nwx;
  const int wx = w - q * nwx;
  WinPos p;
  p.b = q / nwy;
  const int wy = q - p.b * nwy;
  p.yr = wy * g.ws + ty;
  p.xr = wx * g.ws + tx;
  int yp = p.yr + g.shift;
  if (yp >= g.Hp) yp -= g.Hp;
  int xp = p.xr + g.shift;
  if (xp >= g.Wp) xp -= g.Wp;
  p.y = yp - g.pad_top;
  p.x = xp - g.pad_left;
  p.real = (p.y >= 0) && (p.y < g.H) && (p.x >= 0) && (p.x < g.W);
  return p;
}

// Region labels of the two additive -100 masks.  The reference builds the centre-padding mask on the
// UN-rolled padded frame and applies it to the rolled windows (attention.py:273-303), so the padding
// label is a function of the window-frame coordinates, not of where the token came from.
__device__ __forceinline__ int win_pad_label(const atmvfi_window_geom& g, int yr, int xr) {
  int ly = yr < g.pad_top ? 0 : (yr < g.pad_top + g.H ? 1 : 2);
  int lx = xr < g.pad_left ? 0 : (xr < g.pad_left + g.W ? 1 : 2);
  return ly * 3 + lx;
}
__device__ __forceinline__ int win_shift_label(const atmvfi_window_geom& g, int yr, int xr) {
  int ly = yr < g.Hp - g.ws ? 0 : (yr < g.Hp - g.shift ? 1 : 2);
  int lx = xr < g.Wp - g.ws ? 0 : (xr < g.Wp - g.shift ? 1 : 2);
  return ly * 3 + lx;
}
// one packed label: differing labels <=> masked pair
__device__ __forceinline__ int win_mask_label(const atmvfi_window_geom& g, int yr, int xr) {
  int lab = 0;
  if (g.Hp != g.H || g.Wp != g.W) lab = win_pad_label(g, yr, xr);
  if (g.shift) lab = lab * 9 + win_shift_label(g, yr, xr);
  return lab;
}

// ---------------------------------------------------------------------------------------------
// grid_sample(bilinear, zeros, align_corners=True) coordinate arithmetic of flow_warp.py:35-40 replayed
// operation by operation in fp32 (no FMA contraction): pixel + flow -> normalise -> un-normalise.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_src_coord(float pix, float flow, int size) {
  float p = __fadd_rn(pix, flow);
  float gsz = (float)(size - 1);
  float g = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, p), gsz), -1.f);          // 2*x/(w-1) - 1
  return __fmul_rn(__fadd_rn(g, 1.f), __fdiv_rn(gsz, 2.f));              // (g+1) * ((w-1)/2)
}

// flat index over the row window [B][ny][W] -> image b and offset rem = y*W + x inside the FULL [H][W] plane
__device__ __forceinline__ void rw_decode(int64_t i, int W, int y0, int ny, int& b, int& y, int& x) {
  const int64_t per = (int64_t)ny * W;
  b = (int)(i / per);
  const int r = (int)(i - (int64_t)b * per);
  const int yy = r / W;
  y = y0 + yy;
  x = r - yy * W;
}

struct Bilin {
  int x0, y0;
  float wnw, wne, wsw, wse;
  bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ Bilin bilin_setup(float ix, float iy, int W, int H) {
  Bilin s;
  float fx = floorf(ix), fy = floorf(iy);
  s.x0 = (int)fx;
  s.y0 = (int)fy;
  float tx = __fsub_rn(ix, fx), ty = __fsub_rn(iy, fy);
  float ux = __fsub_rn(__fadd_rn(fx, 1.f), ix), uy = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  s.wnw = __fmul_rn(ux, uy);
  s.wne = __fmul_rn(tx, uy);
  s.wsw = __fmul_rn(ux, ty);
  s.wse = __fmul_rn(tx, ty);
  s.vx0 = s.x0 >= 0 && s.x0 < W;
  s.vx1 = s.x0 + 1 >= 0 && s.x0 + 1 < W;
  s.vy0 = s.y0 >= 0 && s.y0 < H;
  s.vy1 = s.y0 + 1 >= 0 && s.y0 + 1 < H;
  // non-finite coordinates (NaN flow) sample nothing, like ATen's bounds checks
  if (!(ix > -2.0e9f && ix < 2.0e9f) || !(iy > -2.0e9f && iy < 2.0e9f)) s.vx0 = s.vx1 = s.vy0 = s.vy1 = false;
  return s;
}

__device__ __forceinline__ float round_tf32_if(float v, bool on) {
  if (on) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    v = __uint_as_float(u);
  }
  return v;
}
__device__ __forceinline__ float4 round_tf32_if(float4 v, bool on) {
  return make_float4(round_tf32_if(v.x, on), round_tf32_if(v.y, on), round_tf32_if(v.z, on), round_tf32_if(v.w, on));
}

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

// GELU for the TF32 / fp16 modes (the result is rounded to a 10-bit mantissa right after): erf by Abramowitz-Stegun 7.1.26,
// 1 - (a1 t + ... + a5 t^5) exp(-x^2), t = 1 / (1 + p |x|): branch-free, 2 MUFU + 13 FP32 instructions instead of the ~45 predicated
// instructions of erff (no denormal scaling around ex2, the final 0.5 v (1 + erf) as one FMA).  |GELU error| <= 5e-7 against
// float64 over [-8.5, 8.5], three orders of magnitude below the TF32 rounding step.  One definition for the stand-alone DWConv
// kernels and the fused Mlp tail, so that they produce identical bits.  The FP32 mode keeps erff.
__device__ __forceinline__ float gelu_fast(float v) {
  const float x = v * 0.70710678118654752440f;
  const float y = v * 0.84932180028801904272f;          // x * sqrt(log2 e): exp(-x^2) = 2^(-y^2)
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, fabsf(x), 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-y * y));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float r = copysignf(fmaf(-p, e, 1.f), x);
  const float h = 0.5f * v;
  return fmaf(h, r, h);
}


// ---------------------------------------------------------------------------------------------
// Element access of channels-last feature maps: fp32, or fp16 storage with fp32 arithmetic (ATMVFI_F16).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Act;
template <>
struct Act<float> {
  static constexpr bool kHalf = false;
  static __device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ float4 lds4(const void* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <>
struct Act<__half> {
  static constexpr bool kHalf = true;
  static __device__ __forceinline__ float4 unpack(uint2 u) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ uint2 pack(float4 v) {
    uint2 u;
    *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v.z, v.w);
    return u;
  }
  static __device__ __forceinline__ float4 ld4(const __half* p) { return unpack(__ldg(reinterpret_cast<const uint2*>(p))); }
  static __device__ __forceinline__ float4 lds4(const void* p) { return unpack(*reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ void st4(__half* p, float4 v) { *reinterpret_cast<uint2*>(p) = pack(v); }
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(__ldg(p)); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  // 8 channels = 16 bytes per access: what the streaming kernels need to keep as many bytes in flight as their fp32 versions
  static __device__ __forceinline__ void ld8(const __half* p, float4& a, float4& b) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    a = unpack(make_uint2(u.x, u.y));
    b = unpack(make_uint2(u.z, u.w));
  }
  static __device__ __forceinline__ void lds8(const void* p, float4& a, float4& b) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    a = unpack(make_uint2(u.x, u.y));
    b = unpack(make_uint2(u.z, u.w));
  }
  static __device__ __forceinline__ void st8(__half* p, float4 a, float4 b) {
    const uint2 x = pack(a), y = pack(b);
    *reinterpret_cast<uint4*>(p) = make_uint4(x.x, x.y, y.x, y.y);
  }
};
