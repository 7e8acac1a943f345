// HBM-bound kernels of the ATM-VFI forward: LayerNorm (plain and fused with the window gather),
// depth-wise 3x3 + GELU, backward warps (+ occlusion blend), align_corners resize, layout packers.
// Every kernel is a coalesced streaming pass; grids are sized in multiples of the SM count (148).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

int atmvfi_dwconv_tma_launch(const void* in, void* out, int B, int H, int W, int C, int pitch, const float* w9c,
                             const float* bias, int y0, int ny, bool rnd, bool f16, cudaStream_t st);

namespace {

constexpr int kSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp normalises one row of C floats (C % 4 == 0); row kept in registers between the passes.
template <int MAXV, typename T>   // 4-channel vectors per lane
__device__ __forceinline__ void ln_row(const T* __restrict__ src, T* __restrict__ dst, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float eps, int lane, bool rnd) {
  float4 v[MAXV];
  const int nv = C >> 2;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int idx = lane + i * 32;
    if (idx < nv) {
      v[i] = Act<T>::ld4(src + 4 * idx);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int idx = lane + i * 32;
    if (idx < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int idx = lane + i * 32;
    if (idx < nv) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
      float4 b = __ldg(reinterpret_cast<const float4*>(beta) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      Act<T>::st4(dst + 4 * idx, round_tf32_if(o, rnd));
    }
  }
}

// fp16 rows with C % 8 == 0: 8 channels (16 bytes) per lane and step, same structure
template <int MAXV>
__device__ __forceinline__ void ln_row8(const __half* __restrict__ src, __half* __restrict__ dst, int C, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float eps, int lane) {
  float4 a[MAXV], b[MAXV];
  const int nv = C >> 3;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      Act<__half>::ld8(src + 8 * idx, a[i], b[i]);
      s += ((a[i].x + a[i].y) + (a[i].z + a[i].w)) + ((b[i].x + b[i].y) + (b[i].z + b[i].w));
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float d0 = a[i].x - mean, d1 = a[i].y - mean, d2 = a[i].z - mean, d3 = a[i].w - mean;
      const float d4 = b[i].x - mean, d5 = b[i].y - mean, d6 = b[i].z - mean, d7 = b[i].w - mean;
      q += ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7));
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * idx), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * idx + 1);
      float4 o0, o1;
      o0.x = (a[i].x - mean) * rstd * g0.x + b0.x; o0.y = (a[i].y - mean) * rstd * g0.y + b0.y;
      o0.z = (a[i].z - mean) * rstd * g0.z + b0.z; o0.w = (a[i].w - mean) * rstd * g0.w + b0.w;
      o1.x = (b[i].x - mean) * rstd * g1.x + b1.x; o1.y = (b[i].y - mean) * rstd * g1.y + b1.y;
      o1.z = (b[i].z - mean) * rstd * g1.z + b1.z; o1.w = (b[i].w - mean) * rstd * g1.w + b1.w;
      Act<__half>::st8(dst + 8 * idx, o0, o1);
    }
  }
}

constexpr int kLnMaxV = 8;   // up to C = 1024

template <typename T>
__device__ __forceinline__ void ln_any(const T* __restrict__ src, T* __restrict__ dst, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int lane, bool rnd) {
  ln_row<kLnMaxV, T>(src, dst, C, gamma, beta, eps, lane, rnd);
}
template <>
__device__ __forceinline__ void ln_any<__half>(const __half* __restrict__ src, __half* __restrict__ dst, int C, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int lane, bool rnd) {
  // (an 8-channel-per-lane variant, ln_row8, measured SLOWER on B200 - 0.30 vs 0.24 ms for the 8 LayerNorm launches of a Base 1080p
  // forward: with C = 384 only 48 of the 64 vector slots of a warp are used, and the kernel is latency- not bandwidth-bound)
  ln_row<kLnMaxV, __half>(src, dst, C, gamma, beta, eps, lane, rnd);
}

template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ in, int in_pitch,
                                                        T* __restrict__ out, int out_pitch, int64_t rows, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, bool rnd) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps)
    ln_any<T>(in + r * in_pitch, out + r * out_pitch, C, gamma, beta, eps, lane, rnd);
}

template <typename T>
__global__ void __launch_bounds__(256) window_gather_ln_kernel(const T* __restrict__ tok, int tok_pitch,
                                                               T* __restrict__ win, int win_pitch, int C,
                                                               atmvfi_window_geom g, int64_t rows, int wy0, int nwy,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, bool rnd) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t per_img = (int64_t)g.Hp * g.Wp, per_win = (int64_t)nwy * g.ws * g.Wp;   // tokens per image: all / in the row window
  const bool small = (int64_t)g.B2 * per_img < (1ll << 31);
  for (int64_t rr = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rr < rows; rr += warps) {
    int64_t r;
    WinPos p;
    if (small) {
      const int bi = (int)rr / (int)per_win;
      const int r32 = bi * (int)per_img + wy0 * g.ws * g.Wp + ((int)rr - bi * (int)per_win);
      r = r32;
      p = win_decode32(g, r32);
    } else {
      const int64_t bi = rr / per_win;
      r = bi * per_img + (int64_t)wy0 * g.ws * g.Wp + (rr - bi * per_win);
      p = win_decode(g, r);
    }
    T* dst = win + r * win_pitch;
    if (p.real) {
      const T* src = tok + ((int64_t)(p.b * g.H + p.y) * g.W + p.x) * tok_pitch;
      ln_any<T>(src, dst, C, gamma, beta, eps, lane, rnd);
    } else {
      // LayerNorm of an all-zero token: (0-0)*rstd*gamma + beta = beta (attention.py:273,316)
      for (int i = lane; i < (C >> 2); i += 32)
        Act<T>::st4(dst + 4 * i, round_tf32_if(__ldg(reinterpret_cast<const float4*>(beta) + i), rnd));
    }
  }
}

// depth-wise 3x3, pad 1, + bias + exact GELU.  One thread owns (x, V channels) and walks down a strip of kDwRows rows.
// Every input row is fetched once per strip and SCATTERED into the accumulators of the three output rows it feeds (taps
// arrive in row-major order, so the sum is formed exactly like a per-pixel loop starting from the bias).  Only the
// weights, three accumulators and two input rows (the current one and the prefetched next one) live in registers:
// A CTA covers kDwCg consecutive channel groups (one warp = one contiguous segment) x kDwX neighbouring columns, so the
// left / right neighbours a thread needs are the centre loads of its CTA mates and hit in L1: only ~1.25 x 1.25 of the
// map crosses the L2 (with one thread per (column, group) flattened over the row, every neighbour came from another
// CTA and the kernel ran at the L2 bandwidth of 3 reads per element).
constexpr int kDwRows = 8;
constexpr int kDwCg = 32, kDwX = 8;          // 256 threads
__device__ __forceinline__ float gelu_exact(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
// (TF32 / fp16 modes: gelu_fast of common.cuh; the FP32 mode keeps erff)

template <int V>
struct DwVec {
  float v[V];
};
template <int V>
__device__ __forceinline__ DwVec<V> dw_load(const __half* p) {      // fp16 maps: V == 4 only
  DwVec<V> r;
  const float4 t = Act<__half>::ld4(p);
  r.v[0] = t.x; r.v[1] = t.y; r.v[2 % V] = t.z; r.v[3 % V] = t.w;
  return r;
}
template <int V>
__device__ __forceinline__ DwVec<V> dw_load(const float* p) {
  DwVec<V> r;
  if (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2 % V] = t.z; r.v[3 % V] = t.w;
  } else {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}

template <int V, bool kFastErf, int kDepth, typename T = float>
__global__ void __launch_bounds__(256) dwconv_gelu_kernel(const T* __restrict__ in, T* __restrict__ out, int B,
                                                          int H, int W, int C, int pitch,
                                                          const float* __restrict__ w9c, const float* __restrict__ bias,
                                                          int wy0, int wy1, bool rnd) {
  // grid: x = (column block, channel-group block), y = row strip, z = image
  const int cv = C / V;
  const int cgb = (cv + kDwCg - 1) / kDwCg;
  const int xb = blockIdx.x / cgb;
  const int cg = (blockIdx.x - xb * cgb) * kDwCg + (threadIdx.x & (kDwCg - 1));
  const int x = xb * kDwX + (threadIdx.x >> 5);
  if (cg >= cv || x >= W) return;
  const int y0 = wy0 + blockIdx.y * kDwRows;
  const int y1 = min(y0 + kDwRows, wy1);
  const int b = blockIdx.z;
  DwVec<V> k[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = dw_load<V>(w9c + t * C + cg * V);
  const DwVec<V> bz = dw_load<V>(bias + cg * V);
  const bool xl = x > 0, xr = x + 1 < W;
  const T* base = in + ((size_t)b * H * W + x) * pitch + cg * V;
  const size_t row_stride = (size_t)W * pitch;
  DwVec<V> zero;
#pragma unroll
  for (int e = 0; e < V; ++e) zero.v[e] = 0.f;
  auto load_row = [&](int yy, DwVec<V>& l, DwVec<V>& m, DwVec<V>& r) {
    if (yy < 0 || yy >= H) { l = m = r = zero; return; }
    const T* rowp = base + yy * row_stride;
    m = dw_load<V>(rowp);
    l = xl ? dw_load<V>(rowp - pitch) : zero;
    r = xr ? dw_load<V>(rowp + pitch) : zero;
  };
  T* obase = out + ((size_t)b * H * W + x) * pitch + cg * V;
  // acc0: output row yy-1 (complete after input row yy), acc1: row yy, acc2: row yy+1
  DwVec<V> acc0 = bz, acc1 = bz, acc2 = bz;
  // register ring of input rows: row (y0 - 1 + i) lives in slot i % (kDepth + 1); kDepth rows are in flight ahead of the
  // one being consumed (the kernel is latency-bound: ncu showed 52 % of the stalls on the first use of the next row)
  DwVec<V> rl[kDepth + 1], rm[kDepth + 1], rr[kDepth + 1];
#pragma unroll
  for (int i = 0; i < kDepth; ++i) load_row(y0 - 1 + i, rl[i], rm[i], rr[i]);
#pragma unroll
  for (int i = 0; i < kDwRows + 2; ++i) {
    const int yy = y0 - 1 + i;
    if (yy > y1) break;
    if (i + kDepth < kDwRows + 2 && yy + kDepth <= y1)
      load_row(yy + kDepth, rl[(i + kDepth) % (kDepth + 1)], rm[(i + kDepth) % (kDepth + 1)], rr[(i + kDepth) % (kDepth + 1)]);
    const DwVec<V>& l = rl[i % (kDepth + 1)];
    const DwVec<V>& m = rm[i % (kDepth + 1)];
    const DwVec<V>& r = rr[i % (kDepth + 1)];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      // input row yy is the bottom row (taps 6-8) of output yy-1, the middle row (3-5) of yy, the top row (0-2) of yy+1
      acc0.v[e] = fmaf(l.v[e], k[6].v[e], acc0.v[e]); acc0.v[e] = fmaf(m.v[e], k[7].v[e], acc0.v[e]); acc0.v[e] = fmaf(r.v[e], k[8].v[e], acc0.v[e]);
      acc1.v[e] = fmaf(l.v[e], k[3].v[e], acc1.v[e]); acc1.v[e] = fmaf(m.v[e], k[4].v[e], acc1.v[e]); acc1.v[e] = fmaf(r.v[e], k[5].v[e], acc1.v[e]);
      acc2.v[e] = fmaf(l.v[e], k[0].v[e], acc2.v[e]); acc2.v[e] = fmaf(m.v[e], k[1].v[e], acc2.v[e]); acc2.v[e] = fmaf(r.v[e], k[2].v[e], acc2.v[e]);
    }
    if (i >= 2) {                                         // output row yy-1 has received its three input rows
      float o[V];
#pragma unroll
      for (int e = 0; e < V; ++e) o[e] = round_tf32_if(kFastErf ? gelu_fast(acc0.v[e]) : gelu_exact(acc0.v[e]), rnd);
      T* op = obase + (size_t)(yy - 1) * row_stride;
      if (V == 4) Act<T>::st4(op, make_float4(o[0], o[1], o[2 % V], o[3 % V]));
      else *reinterpret_cast<float2*>(op) = make_float2(o[0], o[1]);
    }
    acc0 = acc1; acc1 = acc2; acc2 = bz;
  }
}

// First encoder layer: 3x3 conv (pad 1) on a planar 3-channel image + bias + PReLU -> NHWC.  One thread computes all
// COUT channels of one pixel from 27 cached planar loads; weights/bias/slopes sit in shared memory (broadcast reads).
// (Tried: four adjacent pixels per thread - a quarter of the weight reads, float4 image loads - measured 25 % SLOWER on B200: the
// 384-byte lane stride of its stores costs more than the shared-memory reads it saves.  Packed FFMA2 arithmetic, and staging a warp's 3 KB output block in shared memory
// for fully coalesced stores: no change either - neither instruction issue nor the 96-byte-stride stores bound it.)
template <int COUT, typename T>
__global__ void __launch_bounds__(128) conv3x3_first_kernel(const float* __restrict__ img, const float* __restrict__ wk,
                                                            int ldw, const float* __restrict__ bias,
                                                            const float* __restrict__ prelu, T* __restrict__ out,
                                                            int out_pitch, int B, int H, int W, int wy0, int ny, bool rnd) {
  __shared__ __align__(16) float sw[27 * COUT];
  __shared__ float sb[COUT], sp[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = wk[(i / COUT) * ldw + (i % COUT)];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) { sb[i] = bias[i]; sp[i] = prelu ? prelu[i] : 1.f; }
  __syncthreads();
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * ny * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    rw_decode(j, W, wy0, ny, b, y, x);
    const int64_t i = (int64_t)b * hw + (int64_t)y * W + x;
    float v[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = ok ? __ldg(img + ((int64_t)b * 3 + c) * hw + (int64_t)yy * W + xx) : 0.f;
      }
    T* o = out + i * out_pitch;
#pragma unroll
    for (int co = 0; co < COUT; co += 4) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 w4 = *reinterpret_cast<const float4*>(&sw[k * COUT + co]);
        a[0] = fmaf(v[k], w4.x, a[0]); a[1] = fmaf(v[k], w4.y, a[1]); a[2] = fmaf(v[k], w4.z, a[2]); a[3] = fmaf(v[k], w4.w, a[3]);
      }
      float4 r;
      float* rp = &r.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = a[e] + sb[co + e];
        t = t > 0.f ? t : t * sp[co + e];
        rp[e] = round_tf32_if(t, rnd);
      }
      Act<T>::st4(o + co, r);
    }
  }
}

// Up to 5 planar 3-channel images -> 15 (+1 zero) consecutive channels of an NHWC buffer in one pass
// (the image part of torch.cat([feat, im0, I_t_0, im1, I_t_1, I_t], 1), network_base.py:418).
template <typename T>
__global__ void __launch_bounds__(256) pack5_planar_kernel(const float* __restrict__ s0, const float* __restrict__ s1,
                                                           const float* __restrict__ s2, const float* __restrict__ s3,
                                                           const float* __restrict__ s4, T* __restrict__ out,
                                                           int out_pitch, int B, int H, int W, int wy0, int ny, bool rnd) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * ny * W;
  const float* src[5] = {s0, s1, s2, s3, s4};
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    rw_decode(j, W, wy0, ny, b, y, x);
    const int64_t rem = (int64_t)y * W + x, i = (int64_t)b * hw + rem;
    float v[16];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) v[j * 3 + c] = round_tf32_if(__ldg(src[j] + ((int64_t)b * 3 + c) * hw + rem), rnd);
    v[15] = 0.f;
    T* o = out + i * out_pitch;
#pragma unroll
    for (int q = 0; q < 4; ++q) Act<T>::st4(o + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
}

// Row-slab mode: source rows [lo[i], lo[i+1]) live in the arena of another GPU, `delta[i]` bytes away from the local copy
// of the buffer in the peer-mapped address space (0 = local).  The warp reads them in place over NVLink.
struct RowOwners {
  int nseg;
  int lo[ATMVFI_P2P_MAX_PEERS * 2 + 1];
  long long delta[ATMVFI_P2P_MAX_PEERS * 2];
};
__device__ __forceinline__ long long owner_delta(const RowOwners& o, int y) {
  long long d = 0;
#pragma unroll 1
  for (int i = 0; i < o.nseg; ++i)
    if (y >= o.lo[i] && y < o.lo[i + 1]) d = o.delta[i];
  return d;
}

// d0 / d1: byte distance to the GPU that holds source row y0 / y0+1 (row slabs; 0 = local)
__device__ __forceinline__ float sample_plane(const float* __restrict__ p, const Bilin& s, int W, long long d0 = 0, long long d1 = 0) {
  float o = 0.f;
  const float* r0 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p + (int64_t)s.y0 * W + s.x0) + d0);
  const float* r1 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p + (int64_t)(s.y0 + 1) * W + s.x0) + d1);
  if (s.vy0 && s.vx0) o = __fmul_rn(__ldg(r0), s.wnw);
  if (s.vy0 && s.vx1) o = __fadd_rn(o, __fmul_rn(__ldg(r0 + 1), s.wne));
  if (s.vy1 && s.vx0) o = __fadd_rn(o, __fmul_rn(__ldg(r1), s.wsw));
  if (s.vy1 && s.vx1) o = __fadd_rn(o, __fmul_rn(__ldg(r1 + 1), s.wse));
  return o;
}

__global__ void __launch_bounds__(256) flow_warp_nchw_kernel(const float* __restrict__ img, const float* __restrict__ flow,
                                                             float* __restrict__ out, int B, int C, int H, int W, int wy0, int ny) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * ny * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    rw_decode(j, W, wy0, ny, b, y, x);
    const int64_t rem = (int64_t)y * W + x;
    float ix = warp_src_coord((float)x, __ldg(flow + (int64_t)b * 2 * hw + rem), W);
    float iy = warp_src_coord((float)y, __ldg(flow + ((int64_t)b * 2 + 1) * hw + rem), H);
    Bilin s = bilin_setup(ix, iy, W, H);
    for (int c = 0; c < C; ++c) out[((int64_t)b * C + c) * hw + rem] = sample_plane(img + ((int64_t)b * C + c) * hw, s, W);
  }
}

// NHWC gather: a group of (C/4) lanes serves one output pixel, each lane moves one float4 per corner.
template <bool kPeers, typename T = float>
__global__ void __launch_bounds__(256) flow_warp_nhwc_kernel(const T* __restrict__ src, int src_pitch,
                                                             const float* __restrict__ head, int head_pitch, int flow_off,
                                                             T* __restrict__ out, int out_pitch, int B, int C, int H,
                                                             int W, int wy0, int ny, bool rnd, const __grid_constant__ RowOwners own) {
  // One warp per pixel, lanes over its 4-channel groups: the index decode, the coordinate round trip and the tap weights (the
  // bulk of the instructions: four fp32 divisions) are evaluated once per pixel instead of once per channel group, and the four
  // corner reads of a warp are contiguous 512-byte pieces of the source rows.
  const int cv = C >> 2;
  const int lane = threadIdx.x & 31;
  const int64_t npix = (int64_t)B * ny * W, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pi < npix; pi += warps) {
    int b, y, x;
    rw_decode(pi, W, wy0, ny, b, y, x);
    const int64_t pix = ((int64_t)b * H + y) * W + x;
    const float* hp = head + pix * head_pitch + flow_off;
    float ix = warp_src_coord((float)x, __ldg(hp), W);
    float iy = warp_src_coord((float)y, __ldg(hp + 1), H);
    Bilin s = bilin_setup(ix, iy, W, H);
    const T* base = src + (int64_t)b * H * W * src_pitch;
    long long d0 = 0, d1 = 0;
    if (kPeers) {
      if (s.vy0) d0 = owner_delta(own, s.y0);
      if (s.vy1) d1 = owner_delta(own, s.y0 + 1);
    }
    for (int c4 = lane; c4 < cv; c4 += 32) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    auto corner = [&](int yy, int xx, float w, bool first) {
      const T* rowp = base + ((int64_t)yy * W + xx) * src_pitch;
      if (kPeers) rowp = reinterpret_cast<const T*>(reinterpret_cast<const char*>(rowp) + (yy == s.y0 ? d0 : d1));
      float4 v = Act<T>::ld4(rowp + 4 * c4);
      if (first) {
        o.x = __fmul_rn(v.x, w); o.y = __fmul_rn(v.y, w); o.z = __fmul_rn(v.z, w); o.w = __fmul_rn(v.w, w);
      } else {
        o.x = __fadd_rn(o.x, __fmul_rn(v.x, w)); o.y = __fadd_rn(o.y, __fmul_rn(v.y, w));
        o.z = __fadd_rn(o.z, __fmul_rn(v.z, w)); o.w = __fadd_rn(o.w, __fmul_rn(v.w, w));
      }
    };
    if (s.vy0 && s.vx0) corner(s.y0, s.x0, s.wnw, false);
    if (s.vy0 && s.vx1) corner(s.y0, s.x0 + 1, s.wne, false);
    if (s.vy1 && s.vx0) corner(s.y0 + 1, s.x0, s.wsw, false);
    if (s.vy1 && s.vx1) corner(s.y0 + 1, s.x0 + 1, s.wse, false);
    Act<T>::st4(out + pix * out_pitch + 4 * c4, round_tf32_if(o, rnd));
    }
  }
}

template <bool kPeers>
__global__ void __launch_bounds__(256) warp_blend_kernel(const float* __restrict__ im0, const float* __restrict__ im1,
                                                         const float* __restrict__ head, int head_pitch, int head_off,
                                                         float* __restrict__ w0, float* __restrict__ w1,
                                                         float* __restrict__ it, float* __restrict__ flow0,
                                                         float* __restrict__ flow1, float* __restrict__ occ1,
                                                         float* __restrict__ occ2, int B, int H, int W, int wy0, int ny,
                                                         const __grid_constant__ RowOwners own) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * ny * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    rw_decode(j, W, wy0, ny, b, y, x);
    const int64_t rem = (int64_t)y * W + x, i = (int64_t)b * hw + rem;
    const float* hp = head + i * head_pitch + head_off;
    float f0x = __ldg(hp), f0y = __ldg(hp + 1), f1x = __ldg(hp + 2), f1y = __ldg(hp + 3), lg = __ldg(hp + 4);
    Bilin s0 = bilin_setup(warp_src_coord((float)x, f0x, W), warp_src_coord((float)y, f0y, H), W, H);
    Bilin s1 = bilin_setup(warp_src_coord((float)x, f1x, W), warp_src_coord((float)y, f1y, H), W, H);
    float m1 = sigmoidf_exact(lg);
    float m2 = __fsub_rn(1.f, m1);
    long long d00 = 0, d01 = 0, d10 = 0, d11 = 0;       // both sources share the row layout
    if (kPeers) {
      if (s0.vy0) d00 = owner_delta(own, s0.y0);
      if (s0.vy1) d01 = owner_delta(own, s0.y0 + 1);
      if (s1.vy0) d10 = owner_delta(own, s1.y0);
      if (s1.vy1) d11 = owner_delta(own, s1.y0 + 1);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int64_t o = ((int64_t)b * 3 + c) * hw;
      float a = sample_plane(im0 + o, s0, W, d00, d01);
      float bb = sample_plane(im1 + o, s1, W, d10, d11);
      w0[o + rem] = a;
      w1[o + rem] = bb;
      it[o + rem] = __fadd_rn(__fmul_rn(m1, a), __fmul_rn(m2, bb));
    }
    if (flow0) {
      flow0[(int64_t)b * 2 * hw + rem] = f0x;
      flow0[((int64_t)b * 2 + 1) * hw + rem] = f0y;
    }
    if (flow1) {
      flow1[(int64_t)b * 2 * hw + rem] = f1x;
      flow1[((int64_t)b * 2 + 1) * hw + rem] = f1y;
    }
    if (occ1) occ1[i] = m1;
    if (occ2) occ2[i] = m2;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// One level of the global-motion pyramid warp (network_base.py:480-485) in ONE launch: for both frames, (optionally) the x2
// align_corners up-sampling of the coarser level's flow with doubled values (upsample_flow, network_base.py:11-18) fused into
// the backward warp that consumes it (flow_warp.py:50-60).  Replaces 2 resize + 2 warp launches per level.
// A CTA produces a 32 x 8 pixel tile of BOTH frames.  The flows of a tile are smooth (they come from the 1/16 grid), so the
// footprint of the tile in the source image is the tile shifted by its flow plus a small apron: the CTA stages that source
// window (3 channels, zero outside the image = grid_sample's zero padding) in shared memory with coalesced row reads and every
// bilinear tap becomes a shared-memory read.  A tile whose footprint does not fit the window (very divergent flows) falls back
// to the direct global gather.  The arithmetic (coordinate round trip, tap weights, accumulation order) is that of
// resize_ac_kernel / flow_warp_nchw_kernel, operation by operation: the outputs are bit-identical to the unfused launches.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPwTx = 32, kPwTy = 8, kPwApron = 7;
constexpr int kPwSw = kPwTx + 2 * kPwApron + 2, kPwSh = kPwTy + 2 * kPwApron + 2;     // staged source window (48 x 24)

__device__ __forceinline__ float upsample2_ac(const float* __restrict__ src, int Hin, int Win, float sh, float sw, int y, int x) {
  float fy = __fmul_rn(sh, (float)y), fx = __fmul_rn(sw, (float)x);
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
  float ly = __fsub_rn(fy, (float)y0), lx = __fsub_rn(fx, (float)x0);
  float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  float v00 = __ldg(src + (int64_t)y0 * Win + x0), v01 = __ldg(src + (int64_t)y0 * Win + x1);
  float v10 = __ldg(src + (int64_t)y1 * Win + x0), v11 = __ldg(src + (int64_t)y1 * Win + x1);
  float top = __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01));
  float bot = __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11));
  return __fmul_rn(__fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot)), 2.f);
}

__global__ void __launch_bounds__(kPwTx * kPwTy) pyramid_warp_kernel(const float* __restrict__ im0, const float* __restrict__ im1,
                                                                     const float* __restrict__ fin0, const float* __restrict__ fin1, int up,
                                                                     float* __restrict__ out0, float* __restrict__ out1,
                                                                     float* __restrict__ fout0, float* __restrict__ fout1, int B, int H, int W,
                                                                     float sh, float sw, int wy0, int ny) {
  __shared__ float tile[2][3][kPwSh][kPwSw];
  __shared__ int s_box[2][4];                       // per frame: min x0, min y0, max x0, max y0 of the taps' top-left corners
  const int tx = threadIdx.x % kPwTx, ty = threadIdx.x / kPwTx;
  const int tiles_x = (W + kPwTx - 1) / kPwTx, tiles_y = (ny + kPwTy - 1) / kPwTy;
  const int64_t ntiles = (int64_t)B * tiles_y * tiles_x, hw = (int64_t)H * W;
  const int Hc = H >> 1, Wc = W >> 1;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // (32-bit decode: ntiles < 2^31 is checked by the launcher; 64-bit divisions cost ~100 instructions each)
    const int t32 = (int)t, trow = t32 / tiles_x, bx = t32 - trow * tiles_x, b = trow / tiles_y, by = trow - b * tiles_y;
    const int x = bx * kPwTx + tx, y = wy0 + by * kPwTy + ty;
    const bool in_img = x < W && y < wy0 + ny;
    if (threadIdx.x < 8) s_box[threadIdx.x >> 2][threadIdx.x & 3] = (threadIdx.x & 2) ? -(1 << 30) : (1 << 30);
    __syncthreads();
    Bilin sm[2];
    bool finite[2] = {false, false};
    int fx0[2] = {1 << 30, 1 << 30}, fy0[2] = {1 << 30, 1 << 30}, fx1[2] = {-(1 << 30), -(1 << 30)}, fy1[2] = {-(1 << 30), -(1 << 30)};
    if (in_img) {
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float* fin = f ? fin1 : fin0;
        float fx, fy;
        if (up) {
          fx = upsample2_ac(fin + (int64_t)b * 2 * Hc * Wc, Hc, Wc, sh, sw, y, x);
          fy = upsample2_ac(fin + ((int64_t)b * 2 + 1) * Hc * Wc, Hc, Wc, sh, sw, y, x);
          float* fo = f ? fout1 : fout0;
          if (fo) { fo[(int64_t)b * 2 * hw + (int64_t)y * W + x] = fx; fo[((int64_t)b * 2 + 1) * hw + (int64_t)y * W + x] = fy; }
        } else {
          fx = __ldg(fin + (int64_t)b * 2 * hw + (int64_t)y * W + x);
          fy = __ldg(fin + ((int64_t)b * 2 + 1) * hw + (int64_t)y * W + x);
        }
        const float ix = warp_src_coord((float)x, fx, W), iy = warp_src_coord((float)y, fy, H);
        sm[f] = bilin_setup(ix, iy, W, H);
        finite[f] = (ix > -2.0e9f && ix < 2.0e9f) && (iy > -2.0e9f && iy < 2.0e9f);
        if (finite[f] && (sm[f].vx0 || sm[f].vx1) && (sm[f].vy0 || sm[f].vy1)) {      // samples that touch the image define the footprint
          fx0[f] = fx1[f] = sm[f].x0;
          fy0[f] = fy1[f] = sm[f].y0;
        }
      }
    }
    // footprint of the tile: warp-level min / max first, then one shared-memory atomic per warp and bound (256 threads hammering
    // eight words with atomics serialised the kernel: 140 us for the full-resolution level, measured with ncu)
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int a = __reduce_min_sync(0xffffffffu, fx0[f]), b2 = __reduce_min_sync(0xffffffffu, fy0[f]);
      const int c2 = __reduce_max_sync(0xffffffffu, fx1[f]), d2 = __reduce_max_sync(0xffffffffu, fy1[f]);
      if ((threadIdx.x & 31) == 0 && c2 >= a) {
        atomicMin(&s_box[f][0], a); atomicMin(&s_box[f][1], b2);
        atomicMax(&s_box[f][2], c2); atomicMax(&s_box[f][3], d2);
      }
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const float* img = (f ? im1 : im0) + (int64_t)b * 3 * hw;
      float* out = (f ? out1 : out0) + (int64_t)b * 3 * hw;
      const int bx0 = s_box[f][0], by0 = s_box[f][1], bx1 = s_box[f][2], by1 = s_box[f][3];
      const bool any = bx1 >= bx0;
      const bool staged = any && (bx1 - bx0 + 2 <= kPwSw) && (by1 - by0 + 2 <= kPwSh);
      if (staged) {
        const int wx = bx1 - bx0 + 2, wy = by1 - by0 + 2;          // window that holds every tap: [bx0, bx0 + wx) x [by0, by0 + wy)
        // one warp per source row, lanes along it: coalesced reads and no index divisions (the flat loop with its three runtime
        // divisions per element cost ten times the warp arithmetic itself: 140 us for the full-resolution level)
        for (int r = threadIdx.x >> 5; r < wy; r += (kPwTx * kPwTy) >> 5) {
          const int sy = by0 + r;
          const bool rok = sy >= 0 && sy < H;
          for (int q = threadIdx.x & 31; q < wx; q += 32) {
            const int sx = bx0 + q;
            const bool ok = rok && sx >= 0 && sx < W;
            const int64_t off = (int64_t)sy * W + sx;
#pragma unroll
            for (int c = 0; c < 3; ++c) tile[f][c][r][q] = ok ? __ldg(img + (int64_t)c * hw + off) : 0.f;
          }
        }
      }
      __syncthreads();
      if (in_img) {
        const Bilin& s = sm[f];
        const int64_t rem = (int64_t)y * W + x;
        const bool touches = finite[f] && (s.vx0 || s.vx1) && (s.vy0 || s.vy1);
        if (staged && touches) {
          const int r = s.y0 - by0, q = s.x0 - bx0;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // same accumulation as sample_plane: invalid taps are skipped there and read a staged 0 here (x + 0 = x)
            float o = 0.f;
            if (s.vy0 && s.vx0) o = __fmul_rn(tile[f][c][r][q], s.wnw);
            if (s.vy0 && s.vx1) o = __fadd_rn(o, __fmul_rn(tile[f][c][r][q + 1], s.wne));
            if (s.vy1 && s.vx0) o = __fadd_rn(o, __fmul_rn(tile[f][c][r + 1][q], s.wsw));
            if (s.vy1 && s.vx1) o = __fadd_rn(o, __fmul_rn(tile[f][c][r + 1][q + 1], s.wse));
            out[(int64_t)c * hw + rem] = o;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) out[(int64_t)c * hw + rem] = sample_plane(img + (int64_t)c * hw, s, W);
        }
      }
    }
    __syncthreads();
  }
}

// F.interpolate(bilinear, align_corners=True): src = dst * (in-1)/(out-1); ATen's weights w1 = src - floor, w0 = 1 - w1.
__global__ void __launch_bounds__(256) resize_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                        int Hin, int Win, int Hout, int Wout, float sh, float sw,
                                                        float scale, int wy0, int ny) {
  const int64_t total = (int64_t)planes * ny * Wout;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int pl, y, x;
    rw_decode(j, Wout, wy0, ny, pl, y, x);
    const int64_t p = pl, i = ((int64_t)pl * Hout + y) * Wout + x;
    float fy = __fmul_rn(sh, (float)y), fx = __fmul_rn(sw, (float)x);
    int y0 = (int)fy, x0 = (int)fx;
    int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
    float ly = __fsub_rn(fy, (float)y0), lx = __fsub_rn(fx, (float)x0);
    float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    const float* src = in + p * Hin * Win;
    float v00 = __ldg(src + (int64_t)y0 * Win + x0), v01 = __ldg(src + (int64_t)y0 * Win + x1);
    float v10 = __ldg(src + (int64_t)y1 * Win + x0), v11 = __ldg(src + (int64_t)y1 * Win + x1);
    float top = __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01));
    float bot = __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11));
    float v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
    out[i] = scale == 1.f ? v : __fmul_rn(v, scale);
  }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int out_pitch, int chan_off, int B, int C, int H, int W,
                                                           int zero_to, int wy0, int ny, bool rnd) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * ny * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    rw_decode(j, W, wy0, ny, b, y, x);
    const int64_t rem = (int64_t)y * W + x, i = (int64_t)b * hw + rem;
    float* o = out + i * out_pitch + chan_off;
    for (int c = 0; c < C; ++c) o[c] = round_tf32_if(__ldg(in + ((int64_t)b * C + c) * hw + rem), rnd);
    for (int c = chan_off + C; c < zero_to; ++c) out[i * out_pitch + c] = 0.f;
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, int in_pitch, float* __restrict__ out, int B,
                                                           int C, int H, int W) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int64_t rem = i - b * hw;
    for (int c = 0; c < C; ++c) out[((int64_t)b * C + c) * hw + rem] = __ldg(in + i * in_pitch + c);
  }
}

// mean |a - b| per sample, deterministic: fixed-shape tree per block -> kL1Blocks partial sums per sample -> one warp adds
// them in double in a fixed order.
constexpr int kL1Blocks = 512;
__global__ void __launch_bounds__(256) l1_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         float* __restrict__ partial, int64_t n) {
  __shared__ float red[256];
  const int s = blockIdx.y;
  const float* pa = a + (int64_t)s * n;
  const float* pb = b + (int64_t)s * n;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += fabsf(__ldg(pa + i) - __ldg(pb + i));
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(int64_t)s * kL1Blocks + blockIdx.x] = red[0];
}
__global__ void l1_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t n) {
  const int s = blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kL1Blocks; ++i) t += (double)partial[(int64_t)s * kL1Blocks + i];
    out[s] = (float)(t / (double)n);
  }
}

// per sample: the candidate whose loss is the smallest, ties resolved like the reference's if / elif / else chain
// (network_base.py:596-611)
__global__ void __launch_bounds__(256) select_min3_kernel(const float* __restrict__ l0, const float* __restrict__ l1,
                                                          const float* __restrict__ l2, const float* __restrict__ c0,
                                                          const float* __restrict__ c1, const float* __restrict__ c2,
                                                          float* __restrict__ out, int B, int64_t n) {
  const int64_t total = (int64_t)B * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(i / n);
    const float a = __ldg(l0 + s), b = __ldg(l1 + s), c = __ldg(l2 + s);
    const float m = fminf(a, fminf(b, c));
    out[i] = a == m ? __ldg(c0 + i) : (b == m ? __ldg(c1 + i) : __ldg(c2 + i));
  }
}

__global__ void __launch_bounds__(256) residual_finish_kernel(const float* __restrict__ res, int res_pitch,
                                                              const float* __restrict__ it, float* __restrict__ it_sum,
                                                              float* __restrict__ it_clamped, int B, int H, int W, int wy0, int ny) {
  const int64_t hw = (int64_t)H * W, total = (int64_t)B * ny * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    rw_decode(j, W, wy0, ny, b, y, x);
    const int64_t rem = (int64_t)y * W + x, i = (int64_t)b * hw + rem;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int64_t o = ((int64_t)b * 3 + c) * hw + rem;
      float r = __fsub_rn(__fmul_rn(2.f, sigmoidf_exact(__ldg(res + i * res_pitch + c))), 1.f);
      float s = __fadd_rn(__ldg(it + o), r);
      if (it_sum) it_sum[o] = s;
      it_clamped[o] = fminf(fmaxf(s, 0.f), 1.f);
    }
  }
}

__global__ void __launch_bounds__(256) u8_to_planar_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int H,
                                                           int W, int Hp, int Wp, int top, int left, int bgr, int y0, int y1) {
  const int64_t total = (int64_t)Hp * Wp, lo = (int64_t)y0 * Wp, hi = (int64_t)y1 * Wp;      // padded rows [y0, y1) only
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % Wp), y = (int)(i / Wp);
    int sx = min(max(x - left, 0), W - 1), sy = min(max(y - top, 0), H - 1);   // replicate pad (utils.py:66-69)
    const uint8_t* p = in + ((int64_t)sy * W + sx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(int64_t)c * total + i] = __fdiv_rn((float)p[bgr ? 2 - c : c], 255.f);
  }
}

__global__ void __launch_bounds__(256) planar_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int H,
                                                           int W, int Hp, int Wp, int top, int left, int bgr) {
  const int64_t total = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W), y = (int)(i / W);
    int64_t s = (int64_t)(y + top) * Wp + (x + left);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = rintf(__fmul_rn(__ldg(in + (int64_t)c * Hp * Wp + s), 255.f));   // np.round: half to even
      out[i * 3 + (bgr ? 2 - c : c)] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
    }
  }
}

inline int grid_for(int64_t work_items, int block) {
  int64_t blocks = (work_items + block - 1) / block;
  int64_t cap = (int64_t)kSMs * 16;   // 16 resident 256-thread CTAs/SM worth of grid-stride work
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

// fp32 rows [rows][C] (pitch in floats) -> fp16 rows (pitch in halves): the tiny fp32 side products (per-token motion, 5-channel
// motion heads) that also feed an fp16 GEMM as one of its concatenated sources.
__global__ void __launch_bounds__(256) cast_f32_f16_kernel(const float* __restrict__ in, int in_pitch, __half* __restrict__ out, int out_pitch,
                                                           int64_t rows, int C, int zero_to) {
  const int64_t total = rows * zero_to;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / zero_to;
    const int c = (int)(i - r * zero_to);
    out[r * out_pitch + c] = c < C ? __float2half_rn(__ldg(in + r * in_pitch + c)) : __float2half_rn(0.f);
  }
}

extern "C" {

int atmvfi_cast_f32_to_f16(const float* in, int in_pitch, void* out, int out_pitch, int64_t rows, int C, int zero_fill_to, void* stream) {
  ATMVFI_REQUIRE(C > 0 && zero_fill_to <= out_pitch && C <= in_pitch, "cast_f32_to_f16: bad channel counts");
  if (rows <= 0) return 0;
  const int zt = zero_fill_to > C ? zero_fill_to : C;
  cast_f32_f16_kernel<<<grid_for(rows * zt, 256), 256, 0, (cudaStream_t)stream>>>(in, in_pitch, reinterpret_cast<__half*>(out), out_pitch, rows, C, zt);
  ATMVFI_CHECK_LAUNCH("cast_f32_to_f16");
  return 0;
}

int atmvfi_layernorm(const float* in, int in_pitch, float* out, int out_pitch, int64_t rows, int C, const float* gamma,
                     const float* beta, float eps, void* stream) {
  ATMVFI_REQUIRE(C % 4 == 0 && C <= kLnMaxV * 128 && in_pitch % 4 == 0 && out_pitch % 4 == 0,
                 "layernorm: C=%d pitches %d/%d unsupported (need C%%4==0, C<=%d)", C, in_pitch, out_pitch, kLnMaxV * 128);
  if (rows <= 0) return 0;
  if (atmvfi_act_f16())
    launch_pdl(layernorm_kernel<__half>, grid_for(rows, 8), 256, (cudaStream_t)stream, reinterpret_cast<const __half*>(in), in_pitch, reinterpret_cast<__half*>(out), out_pitch, rows, C, gamma, beta, eps, false);
  else
    launch_pdl(layernorm_kernel<float>, grid_for(rows, 8), 256, (cudaStream_t)stream, in, in_pitch, out, out_pitch, rows, C, gamma, beta, eps, atmvfi_output_rounding() != 0);
  ATMVFI_CHECK_LAUNCH("layernorm");
  return 0;
}

int atmvfi_window_gather_ln(const float* tok, int tok_pitch, float* win, int win_pitch, int C, const atmvfi_window_geom* g,
                            const float* gamma, const float* beta, float eps, int wy0, int wy1, void* stream) {
  ATMVFI_REQUIRE(C % 4 == 0 && C <= kLnMaxV * 128 && tok_pitch % 4 == 0 && win_pitch % 4 == 0, "window_gather_ln: C=%d unsupported", C);
  ATMVFI_REQUIRE(g->Hp % g->ws == 0 && g->Wp % g->ws == 0 && g->shift >= 0 && g->shift < g->ws, "window_gather_ln: bad geometry");
  int w0, nwy;
  ATMVFI_REQUIRE(row_window(g->Hp / g->ws, wy0, wy1, &w0, &nwy), "window_gather_ln: bad window-row range [%d,%d)", wy0, wy1);
  int64_t rows = (int64_t)g->B2 * nwy * g->ws * g->Wp;
  if (rows <= 0) return 0;
  if (atmvfi_act_f16())
    launch_pdl(window_gather_ln_kernel<__half>, grid_for(rows, 8), 256, (cudaStream_t)stream, reinterpret_cast<const __half*>(tok), tok_pitch, reinterpret_cast<__half*>(win), win_pitch, C, *g, rows, w0, nwy, gamma, beta, eps, false);
  else
    launch_pdl(window_gather_ln_kernel<float>, grid_for(rows, 8), 256, (cudaStream_t)stream, tok, tok_pitch, win, win_pitch, C, *g, rows, w0, nwy, gamma, beta, eps, atmvfi_output_rounding() != 0);
  ATMVFI_CHECK_LAUNCH("window_gather_ln");
  return 0;
}

int atmvfi_conv3x3_first(const float* img, const float* wk, int ldw, const float* bias, const float* prelu, float* out,
                         int out_pitch, int B, int H, int W, int Cout, int y0, int y1, void* stream) {
  ATMVFI_REQUIRE(out_pitch % 4 == 0 && ((uintptr_t)out & 15) == 0, "conv3x3_first: output must be 16-byte aligned, pitch %% 4 == 0");
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "conv3x3_first: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W;
  if (n <= 0) return 0;
  const bool rnd = atmvfi_output_rounding() != 0;
  int grid = grid_for(n, 128);
  cudaStream_t st = (cudaStream_t)stream;
  const bool f16 = atmvfi_act_f16() != 0;
  __half* outh = reinterpret_cast<__half*>(out);
  switch (Cout) {
    case 16:
      if (f16) conv3x3_first_kernel<16, __half><<<grid, 128, 0, st>>>(img, wk, ldw, bias, prelu, outh, out_pitch, B, H, W, y0, ny, false);
      else conv3x3_first_kernel<16, float><<<grid, 128, 0, st>>>(img, wk, ldw, bias, prelu, out, out_pitch, B, H, W, y0, ny, rnd);
      break;
    case 24:
      if (f16) conv3x3_first_kernel<24, __half><<<grid, 128, 0, st>>>(img, wk, ldw, bias, prelu, outh, out_pitch, B, H, W, y0, ny, false);
      else conv3x3_first_kernel<24, float><<<grid, 128, 0, st>>>(img, wk, ldw, bias, prelu, out, out_pitch, B, H, W, y0, ny, rnd);
      break;
    case 32:
      if (f16) conv3x3_first_kernel<32, __half><<<grid, 128, 0, st>>>(img, wk, ldw, bias, prelu, outh, out_pitch, B, H, W, y0, ny, false);
      else conv3x3_first_kernel<32, float><<<grid, 128, 0, st>>>(img, wk, ldw, bias, prelu, out, out_pitch, B, H, W, y0, ny, rnd);
      break;
    default:
      atmvfi_set_error("conv3x3_first: Cout=%d not instantiated (16, 24, 32)", Cout);
      return 2;
  }
  ATMVFI_CHECK_LAUNCH("conv3x3_first");
  return 0;
}

int atmvfi_pack5_planar(const float* s0, const float* s1, const float* s2, const float* s3, const float* s4, float* out,
                        int out_pitch, int B, int H, int W, int y0, int y1, void* stream) {
  ATMVFI_REQUIRE(out_pitch >= 16 && out_pitch % 4 == 0 && ((uintptr_t)out & 15) == 0, "pack5_planar: output needs pitch >= 16 and 16-byte alignment");
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "pack5_planar: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W;
  if (n <= 0) return 0;
  if (atmvfi_act_f16())
    pack5_planar_kernel<__half><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s0, s1, s2, s3, s4, reinterpret_cast<__half*>(out), out_pitch, B, H, W, y0, ny, false);
  else
    pack5_planar_kernel<float><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s0, s1, s2, s3, s4, out, out_pitch, B, H, W, y0, ny, atmvfi_output_rounding() != 0);
  ATMVFI_CHECK_LAUNCH("pack5_planar");
  return 0;
}

int atmvfi_dwconv3x3_gelu(const float* in, float* out, int B, int H, int W, int C, int pitch, const float* w9c,
                          const float* bias, int y0, int y1, void* stream) {
  ATMVFI_REQUIRE(C % 4 == 0 && pitch % 4 == 0, "dwconv3x3_gelu: C=%d pitch=%d must be multiples of 4", C, pitch);
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "dwconv3x3_gelu: bad row window [%d,%d)", y0, y1);
  if ((int64_t)B * ny * W <= 0) return 0;
  ATMVFI_REQUIRE((int64_t)W * (C / 4) < (1 << 30) && B <= 65535, "dwconv3x3_gelu: shape out of range");
  static_assert(kDwCg * kDwX == 256, "dwconv CTA shape");
  const bool f16 = atmvfi_act_f16() != 0;
  {   // TMA-fed streaming kernel (dwconv_tma.cu); falls through to the register-ring kernel when it does not apply
    const int rc = atmvfi_dwconv_tma_launch(in, out, B, H, W, C, pitch, w9c, bias, y0, ny, !f16 && atmvfi_output_rounding() != 0, f16, (cudaStream_t)stream);
    if (rc != 3) return rc;
  }
  if (f16) {      // fp16 maps: register-ring kernel, 4 channels per thread
    dim3 grid((unsigned)(((C / 4 + kDwCg - 1) / kDwCg) * ((W + kDwX - 1) / kDwX)), (unsigned)((ny + kDwRows - 1) / kDwRows), (unsigned)B);
    dwconv_gelu_kernel<4, true, 3, __half><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out), B, H, W, C, pitch, w9c, bias, y0, y0 + ny, false);
    ATMVFI_CHECK_LAUNCH("dwconv3x3_gelu");
    return 0;
  }
  static int vsel = 0, dsel = 0;
  if (!vsel) { const char* ev = getenv("ATMVFI_DW_V"); vsel = ev ? atoi(ev) : 4; }
  if (!dsel) { const char* ev = getenv("ATMVFI_DW_DEPTH"); dsel = ev ? atoi(ev) : 3; }
  const bool rnd = atmvfi_output_rounding() != 0;
#define ATMVFI_DW_LAUNCH(V_, D_)                                                                                                        \
  do {                                                                                                                                  \
    dim3 grid((unsigned)(((C / V_ + kDwCg - 1) / kDwCg) * ((W + kDwX - 1) / kDwX)), (unsigned)((ny + kDwRows - 1) / kDwRows), (unsigned)B); \
    if (rnd) dwconv_gelu_kernel<V_, true, D_><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, pitch, w9c, bias, y0, y0 + ny, rnd);  \
    else dwconv_gelu_kernel<V_, false, D_><<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, pitch, w9c, bias, y0, y0 + ny, rnd);     \
  } while (0)
  if (vsel == 2) {
    if (dsel == 1) ATMVFI_DW_LAUNCH(2, 1); else if (dsel == 2) ATMVFI_DW_LAUNCH(2, 2); else if (dsel == 3) ATMVFI_DW_LAUNCH(2, 3); else ATMVFI_DW_LAUNCH(2, 4);
  } else {
    if (dsel == 1) ATMVFI_DW_LAUNCH(4, 1); else if (dsel == 2) ATMVFI_DW_LAUNCH(4, 2); else if (dsel == 3) ATMVFI_DW_LAUNCH(4, 3); else ATMVFI_DW_LAUNCH(4, 4);
  }
#undef ATMVFI_DW_LAUNCH
  ATMVFI_CHECK_LAUNCH("dwconv3x3_gelu");
  return 0;
}

int atmvfi_flow_warp_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W, int y0, int y1, void* stream) {
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "flow_warp_nchw: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W;
  if (n <= 0 || C <= 0) return 0;
  flow_warp_nchw_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(img, flow, out, B, C, H, W, y0, ny);
  ATMVFI_CHECK_LAUNCH("flow_warp_nchw");
  return 0;
}

static int flow_warp_nhwc_impl(const float* src, int src_pitch, const float* head, int head_pitch, int flow_off, float* out,
                               int out_pitch, int B, int C, int H, int W, int y0, int y1, const atmvfi_row_owners* owners, void* stream) {
  ATMVFI_REQUIRE(C % 4 == 0 && src_pitch % 4 == 0 && out_pitch % 4 == 0, "flow_warp_nhwc: C=%d must be a multiple of 4", C);
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "flow_warp_nhwc: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W * 32;          // one warp per pixel
  if (n <= 0) return 0;
  RowOwners own;
  memset(&own, 0, sizeof(own));
  const bool rnd = atmvfi_output_rounding() != 0;
  if (owners && owners->nseg > 0) {
    ATMVFI_REQUIRE(!atmvfi_act_f16(), "flow_warp_nhwc_p2p: the row-slab mode runs on fp32 feature maps (tf32 / fp32 / fp32x3 precision)");
    ATMVFI_REQUIRE(owners->nseg <= ATMVFI_P2P_MAX_PEERS * 2, "flow_warp_nhwc: %d owner segments (max %d)", owners->nseg, ATMVFI_P2P_MAX_PEERS * 2);
    own.nseg = owners->nseg;
    for (int i = 0; i < owners->nseg; ++i) { own.lo[i] = owners->row_lo[i]; own.delta[i] = owners->byte_delta[i]; }
    own.lo[owners->nseg] = owners->row_lo[owners->nseg];
    flow_warp_nhwc_kernel<true><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, src_pitch, head, head_pitch, flow_off, out, out_pitch, B, C, H, W, y0, ny, rnd, own);
  } else if (atmvfi_act_f16()) {      // fp16 feature maps (the flows in `head` stay fp32)
    flow_warp_nhwc_kernel<false, __half><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(src), src_pitch, head, head_pitch, flow_off, reinterpret_cast<__half*>(out), out_pitch, B, C, H, W, y0, ny, false, own);
  } else {
    flow_warp_nhwc_kernel<false><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, src_pitch, head, head_pitch, flow_off, out, out_pitch, B, C, H, W, y0, ny, rnd, own);
  }
  ATMVFI_CHECK_LAUNCH("flow_warp_nhwc");
  return 0;
}

int atmvfi_flow_warp_nhwc(const float* src, int src_pitch, const float* head, int head_pitch, int flow_off, float* out,
                          int out_pitch, int B, int C, int H, int W, int y0, int y1, void* stream) {
  return flow_warp_nhwc_impl(src, src_pitch, head, head_pitch, flow_off, out, out_pitch, B, C, H, W, y0, y1, nullptr, stream);
}

int atmvfi_flow_warp_nhwc_p2p(const float* src, int src_pitch, const float* head, int head_pitch, int flow_off, float* out,
                              int out_pitch, int B, int C, int H, int W, int y0, int y1, const atmvfi_row_owners* owners, void* stream) {
  return flow_warp_nhwc_impl(src, src_pitch, head, head_pitch, flow_off, out, out_pitch, B, C, H, W, y0, y1, owners, stream);
}

static int warp_blend_impl(const float* im0, const float* im1, const float* head, int head_pitch, int head_off, float* w0,
                           float* w1, float* it, float* flow0, float* flow1, float* occ1, float* occ2, int B, int H, int W,
                           int y0, int y1, const atmvfi_row_owners* owners, void* stream) {
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "warp_blend: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W;
  if (n <= 0) return 0;
  RowOwners own;
  memset(&own, 0, sizeof(own));
  if (owners && owners->nseg > 0) {
    ATMVFI_REQUIRE(owners->nseg <= ATMVFI_P2P_MAX_PEERS * 2, "warp_blend: %d owner segments (max %d)", owners->nseg, ATMVFI_P2P_MAX_PEERS * 2);
    own.nseg = owners->nseg;
    for (int i = 0; i < owners->nseg; ++i) { own.lo[i] = owners->row_lo[i]; own.delta[i] = owners->byte_delta[i]; }
    own.lo[owners->nseg] = owners->row_lo[owners->nseg];
    warp_blend_kernel<true><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(im0, im1, head, head_pitch, head_off, w0, w1, it, flow0, flow1, occ1, occ2, B, H, W, y0, ny, own);
  } else {
    warp_blend_kernel<false><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(im0, im1, head, head_pitch, head_off, w0, w1, it, flow0, flow1, occ1, occ2, B, H, W, y0, ny, own);
  }
  ATMVFI_CHECK_LAUNCH("warp_blend");
  return 0;
}

int atmvfi_warp_blend(const float* im0, const float* im1, const float* head, int head_pitch, int head_off, float* w0,
                      float* w1, float* it, float* flow0, float* flow1, float* occ1, float* occ2, int B, int H, int W,
                      int y0, int y1, void* stream) {
  return warp_blend_impl(im0, im1, head, head_pitch, head_off, w0, w1, it, flow0, flow1, occ1, occ2, B, H, W, y0, y1, nullptr, stream);
}

int atmvfi_warp_blend_p2p(const float* im0, const float* im1, const float* head, int head_pitch, int head_off, float* w0,
                          float* w1, float* it, float* flow0, float* flow1, float* occ1, float* occ2, int B, int H, int W,
                          int y0, int y1, const atmvfi_row_owners* owners, void* stream) {
  return warp_blend_impl(im0, im1, head, head_pitch, head_off, w0, w1, it, flow0, flow1, occ1, occ2, B, H, W, y0, y1, owners, stream);
}

int atmvfi_resize_bilinear_ac(const float* in, float* out, int planes, int Hin, int Win, int Hout, int Wout, float scale,
                              int y0, int y1, void* stream) {
  int ny;
  ATMVFI_REQUIRE(row_window(Hout, y0, y1, &y0, &ny), "resize_bilinear_ac: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)planes * ny * Wout;
  if (n <= 0) return 0;
  // ATen area_pixel_compute_scale(align_corners=True): (in-1)/(out-1), 0 when out == 1
  float sh = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  float sw = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  resize_ac_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, Hin, Win, Hout, Wout, sh, sw, scale, y0, ny);
  ATMVFI_CHECK_LAUNCH("resize_bilinear_ac");
  return 0;
}

int atmvfi_pyramid_warp(const float* im0, const float* im1, const float* flow0, const float* flow1, int upsample, float* out0, float* out1,
                        float* flow0_out, float* flow1_out, int B, int H, int W, int y0, int y1, void* stream) {
  ATMVFI_REQUIRE(im0 && im1 && flow0 && flow1 && out0 && out1, "pyramid_warp: null argument");
  ATMVFI_REQUIRE(!upsample || (H % 2 == 0 && W % 2 == 0), "pyramid_warp: the x2 flow up-sampling needs even H, W (got %dx%d)", H, W);
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "pyramid_warp: bad row window [%d,%d)", y0, y1);
  const int64_t tiles = (int64_t)B * ((ny + kPwTy - 1) / kPwTy) * ((W + kPwTx - 1) / kPwTx);
  if (tiles <= 0) return 0;
  ATMVFI_REQUIRE(tiles < (1ll << 31), "pyramid_warp: shape out of range");
  const int Hc = H / 2, Wc = W / 2;
  const float sh = H > 1 ? (float)(Hc - 1) / (float)(H - 1) : 0.f, sw = W > 1 ? (float)(Wc - 1) / (float)(W - 1) : 0.f;
  const int grid = (int)(tiles < (int64_t)kSMs * 8 ? tiles : (int64_t)kSMs * 8);
  pyramid_warp_kernel<<<grid, kPwTx * kPwTy, 0, (cudaStream_t)stream>>>(im0, im1, flow0, flow1, upsample, out0, out1, flow0_out, flow1_out, B, H, W, sh,
                                                                        sw, y0, ny);
  ATMVFI_CHECK_LAUNCH("pyramid_warp");
  return 0;
}

int atmvfi_nchw_to_nhwc(const float* in, float* out, int out_pitch, int chan_off, int B, int C, int H, int W,
                        int zero_fill_to, int y0, int y1, void* stream) {
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "nchw_to_nhwc: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W;
  if (n <= 0) return 0;
  nchw_to_nhwc_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, out_pitch, chan_off, B, C, H, W, zero_fill_to, y0, ny, atmvfi_output_rounding() != 0);
  ATMVFI_CHECK_LAUNCH("nchw_to_nhwc");
  return 0;
}

int atmvfi_nhwc_to_nchw(const float* in, int in_pitch, float* out, int B, int C, int H, int W, void* stream) {
  int64_t n = (int64_t)B * H * W;
  if (n <= 0 || C <= 0) return 0;
  nhwc_to_nchw_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, in_pitch, out, B, C, H, W);
  ATMVFI_CHECK_LAUNCH("nhwc_to_nchw");
  return 0;
}

int atmvfi_l1_mean_scratch_floats(int samples) { return samples * kL1Blocks; }

int atmvfi_l1_mean(const float* a, const float* b, float* out, float* scratch, int samples, int64_t n, void* stream) {
  ATMVFI_REQUIRE(samples > 0 && samples <= 65535 && n > 0 && scratch != nullptr, "l1_mean: bad arguments");
  l1_partial_kernel<<<dim3(kL1Blocks, (unsigned)samples), 256, 0, (cudaStream_t)stream>>>(a, b, scratch, n);
  ATMVFI_CHECK_LAUNCH("l1_mean(partial)");
  l1_final_kernel<<<samples, 32, 0, (cudaStream_t)stream>>>(scratch, out, n);
  ATMVFI_CHECK_LAUNCH("l1_mean(final)");
  return 0;
}

int atmvfi_select_min3(const float* l0, const float* l1, const float* l2, const float* c0, const float* c1, const float* c2,
                       float* out, int samples, int64_t n, void* stream) {
  if ((int64_t)samples * n <= 0) return 0;
  select_min3_kernel<<<grid_for((int64_t)samples * n, 256), 256, 0, (cudaStream_t)stream>>>(l0, l1, l2, c0, c1, c2, out, samples, n);
  ATMVFI_CHECK_LAUNCH("select_min3");
  return 0;
}

int atmvfi_residual_finish(const float* res, int res_pitch, const float* it, float* it_sum, float* it_clamped, int B, int H,
                           int W, int y0, int y1, void* stream) {
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "residual_finish: bad row window [%d,%d)", y0, y1);
  int64_t n = (int64_t)B * ny * W;
  if (n <= 0) return 0;
  residual_finish_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(res, res_pitch, it, it_sum, it_clamped, B, H, W, y0, ny);
  ATMVFI_CHECK_LAUNCH("residual_finish");
  return 0;
}

int atmvfi_u8_to_planar(const uint8_t* in, float* out, int H, int W, int Hp, int Wp, int top, int left, int bgr, void* stream) {
  u8_to_planar_kernel<<<grid_for((int64_t)Hp * Wp, 256), 256, 0, (cudaStream_t)stream>>>(in, out, H, W, Hp, Wp, top, left, bgr, 0, Hp);
  ATMVFI_CHECK_LAUNCH("u8_to_planar");
  return 0;
}

int atmvfi_u8_to_planar_rows(const uint8_t* in, float* out, int H, int W, int Hp, int Wp, int top, int left, int bgr, int y0, int y1,
                             void* stream) {
  ATMVFI_REQUIRE(0 <= y0 && y0 <= y1 && y1 <= Hp, "u8_to_planar_rows: bad row window [%d,%d) of %d padded rows", y0, y1, Hp);
  if (y1 == y0) return 0;
  u8_to_planar_kernel<<<grid_for((int64_t)(y1 - y0) * Wp, 256), 256, 0, (cudaStream_t)stream>>>(in, out, H, W, Hp, Wp, top, left, bgr, y0, y1);
  ATMVFI_CHECK_LAUNCH("u8_to_planar_rows");
  return 0;
}

int atmvfi_planar_to_u8(const float* in, uint8_t* out, int H, int W, int Hp, int Wp, int top, int left, int bgr, void* stream) {
  planar_to_u8_kernel<<<grid_for((int64_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(in, out, H, W, Hp, Wp, top, left, bgr);
  ATMVFI_CHECK_LAUNCH("planar_to_u8");
  return 0;
}

}  // extern "C"
