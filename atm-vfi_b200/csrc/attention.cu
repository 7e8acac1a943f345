// Window attention with the attention-to-motion reduction (attention.py:187-213, 370-390), fp32 SIMT.
//
// One CTA per (window, head); thread i owns query row i and streams the keys with an online softmax,
// accumulating  P*V  and the motion expectation  sum_j P_ij * relative_coord[:, i, j]  in registers.
// K and V of the (possibly other-frame) window live in shared memory and are read as warp broadcasts.
// Windows are addressed in place: masks are evaluated from the geometry, never materialised.
// A second tiny kernel applies the head-mix MLP (8 -> 4 -> GELU -> 1) and scatters the motion to the
// token grid, undoing window partition / roll / centre padding.
#include "common.cuh"

namespace {

template <int HD>
__global__ void __launch_bounds__(256) window_attention_kernel(const float* __restrict__ qkv, int qkv_pitch,
                                                               float* __restrict__ out, int out_pitch, int C, int heads,
                                                               atmvfi_window_geom g, int cross,
                                                               const float* __restrict__ rc, float* __restrict__ motion_raw, int wy0, int nwy, bool rnd,
                                                               int layout) {
  extern __shared__ float smem[];
  const int N = g.ws * g.ws;
  float* sk = smem;             // [N][HD]
  float* sv = smem + N * HD;    // [N][HD]
  int* slab = reinterpret_cast<int*>(smem + 2 * N * HD);   // [N] mask labels

  const int h = blockIdx.y;
  const int nW = (g.Hp / g.ws) * (g.Wp / g.ws);
  // blockIdx.x enumerates the windows of the row window [wy0, wy0+nwy) of every image
  const int per_img = nwy * (g.Wp / g.ws);
  const int64_t win = (int64_t)(blockIdx.x / per_img) * nW + wy0 * (g.Wp / g.ws) + blockIdx.x % per_img;
  const int64_t total_win = (int64_t)g.B2 * nW;
  // the other frame's copy of this window sits half the window batch away (attention.py:318)
  const int64_t kv_win = cross ? (win + total_win / 2) % total_win : win;

  const int64_t R = total_win * N;                 // head-major layout (ATMVFI_OUT_QKV_HEADS): rows of the whole tensor
  if (layout == 0) {
    const float* kbase = qkv + kv_win * N * qkv_pitch + C + h * HD;
    const float* vbase = kbase + C;
    for (int i = threadIdx.x; i < N * (HD / 4); i += blockDim.x) {
      int r = i / (HD / 4), c4 = i % (HD / 4);
      reinterpret_cast<float4*>(sk + r * HD)[c4] = __ldg(reinterpret_cast<const float4*>(kbase + (int64_t)r * qkv_pitch) + c4);
      reinterpret_cast<float4*>(sv + r * HD)[c4] = __ldg(reinterpret_cast<const float4*>(vbase + (int64_t)r * qkv_pitch) + c4);
    }
  } else {
    const float* kbase = qkv + ((int64_t)(heads + h) * R + kv_win * N) * HD;          // K[h][r][d]
    for (int i = threadIdx.x; i < N * (HD / 4); i += blockDim.x)
      reinterpret_cast<float4*>(sk)[i] = __ldg(reinterpret_cast<const float4*>(kbase) + i);
    const float* vbase = qkv + 2 * (int64_t)C * R + (int64_t)h * HD * R + kv_win * N;   // V^T[h][d][r]
    for (int i = threadIdx.x; i < N * HD; i += blockDim.x) {
      const int c = i / N, r = i - c * N;
      sv[r * HD + c] = __ldg(vbase + (int64_t)c * R + r);
    }
  }
  const bool masked = (g.shift != 0) || g.Hp != g.H || g.Wp != g.W;
  const int wy = (int)((win % nW) / (g.Wp / g.ws)), wx = (int)((win % nW) % (g.Wp / g.ws));
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    slab[i] = masked ? win_mask_label(g, wy * g.ws + i / g.ws, wx * g.ws + i % g.ws) : 0;
  __syncthreads();

  const int i = threadIdx.x;
  if (i >= N) return;
  const int64_t row = win * N + i;
  float q[HD];
  {
    const float4* qp = reinterpret_cast<const float4*>(layout == 0 ? qkv + row * qkv_pitch + h * HD : qkv + ((int64_t)h * R + row) * HD);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      float4 t = __ldg(qp + d);
      q[4 * d] = t.x; q[4 * d + 1] = t.y; q[4 * d + 2] = t.z; q[4 * d + 3] = t.w;
    }
  }
  const float scale = (float)(1.0 / sqrt((double)HD));   // python: head_dim ** -0.5, folded at compile time
  const int mylab = slab[i];
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f, mx = 0.f, my = 0.f;
  const float* rcx = rc ? rc + (int64_t)i * N : nullptr;
  const float* rcy = rc ? rc + (int64_t)N * N + (int64_t)i * N : nullptr;
  for (int j = 0; j < N; ++j) {
    const float4* kp = reinterpret_cast<const float4*>(sk + j * HD);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      float4 k4 = kp[d];
      s0 = fmaf(q[4 * d], k4.x, s0);
      s1 = fmaf(q[4 * d + 1], k4.y, s1);
      s2 = fmaf(q[4 * d + 2], k4.z, s2);
      s3 = fmaf(q[4 * d + 3], k4.w, s3);
    }
    float s = ((s0 + s1) + (s2 + s3)) * scale;
    if (slab[j] != mylab) s += -100.0f;           // additive mask, -100 not -inf (attention.py:55-57)
    float mn = fmaxf(m, s);
    float corr = expf(m - mn);                    // exp(-inf) = 0 on the first key
    float p = expf(s - mn);
    l = l * corr + p;
    const float4* vp = reinterpret_cast<const float4*>(sv + j * HD);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      float4 v4 = vp[d];
      acc[4 * d] = fmaf(p, v4.x, acc[4 * d] * corr);
      acc[4 * d + 1] = fmaf(p, v4.y, acc[4 * d + 1] * corr);
      acc[4 * d + 2] = fmaf(p, v4.z, acc[4 * d + 2] * corr);
      acc[4 * d + 3] = fmaf(p, v4.w, acc[4 * d + 3] * corr);
    }
    if (rc) {
      mx = fmaf(p, __ldg(rcx + j), mx * corr);
      my = fmaf(p, __ldg(rcy + j), my * corr);
    }
    m = mn;
  }
  const float inv = 1.0f / l;
  float4* op = reinterpret_cast<float4*>(out + row * out_pitch + h * HD);
#pragma unroll
  for (int d = 0; d < HD / 4; ++d)
    op[d] = round_tf32_if(make_float4(acc[4 * d] * inv, acc[4 * d + 1] * inv, acc[4 * d + 2] * inv, acc[4 * d + 3] * inv), rnd);
  if (motion_raw) {
    motion_raw[(row * heads + h) * 2 + 0] = mx * inv;
    motion_raw[(row * heads + h) * 2 + 1] = my * inv;
  }
}

// motion[b, y, x, off + frame*2 + xy] = w2 . gelu(w0 . m_heads + b0) + b2     (attention.py:143-146, 209-211)
__global__ void __launch_bounds__(256) motion_mix_kernel(const float* __restrict__ motion_raw, int heads,
                                                         atmvfi_window_geom g, int64_t rows, int wy0, int nwy,
                                                         const float* __restrict__ w0, const float* __restrict__ b0,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ motion, int motion_pitch, int motion_off) {
  const int hid = heads / 2;
  const int pairs = g.B2 / 2;
  const int64_t per_img = (int64_t)g.Hp * g.Wp, per_win = (int64_t)nwy * g.ws * g.Wp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * 2; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rr = t >> 1;
    const int64_t bi = rr / per_win;
    const int64_t r = bi * per_img + (int64_t)wy0 * g.ws * g.Wp + (rr - bi * per_win);
    int xy = (int)(t & 1);
    WinPos p = win_decode(g, r);
    if (!p.real) continue;
    float o = __ldg(b2);
    for (int k = 0; k < hid; ++k) {
      float a = __ldg(b0 + k);
      for (int hh = 0; hh < heads; ++hh) a = fmaf(__ldg(w0 + k * heads + hh), __ldg(motion_raw + (r * heads + hh) * 2 + xy), a);
      a = 0.5f * a * (1.f + erff(a * 0.70710678118654752440f));
      o = fmaf(__ldg(w2 + k), a, o);
    }
    int frame = p.b >= pairs ? 1 : 0;
    int pb = p.b - frame * pairs;
    motion[((int64_t)(pb * g.H + p.y) * g.W + p.x) * motion_pitch + motion_off + frame * 2 + xy] = o;
  }
}

template <int HD>
int launch_attention(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                     const atmvfi_window_geom& g, int cross, const float* rc, float* motion_raw, int wy0, int nwy, int layout, cudaStream_t st) {
  const int N = g.ws * g.ws;
  const int64_t wins = (int64_t)g.B2 * nwy * (g.Wp / g.ws);
  if (wins <= 0) return 0;
  size_t smem = (size_t)(2 * N * HD) * sizeof(float) + N * sizeof(int);
  auto kern = window_attention_kernel<HD>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      atmvfi_set_error("window_attention: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return 1;
    }
  }
  dim3 grid((unsigned)wins, (unsigned)heads);
  int threads = ((N + 31) / 32) * 32;
  kern<<<grid, threads, smem, st>>>(qkv, qkv_pitch, out, out_pitch, C, heads, g, cross, rc, motion_raw, wy0, nwy, atmvfi_output_rounding() != 0, layout);
  ATMVFI_CHECK_LAUNCH("window_attention");
  return 0;
}

}  // namespace

int atmvfi_window_attention_tc_launch(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                                      const atmvfi_window_geom* g, int cross, const float* rc, float* motion_raw, int wy0, int nwy, int layout,
                                      cudaStream_t st);

static int attention_impl(bool tensor_cores, int rc_closed_form, const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                          const atmvfi_window_geom* g, int cross, const float* relative_coord, const float* mix_w0, const float* mix_b0,
                          const float* mix_w2, const float* mix_b2, float* motion, int motion_pitch, int motion_off, float* scratch,
                          int wy0, int wy1, int layout, void* stream);

extern "C" int atmvfi_window_attention_tc(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                                          const atmvfi_window_geom* g, int cross, const float* relative_coord, int rc_closed_form,
                                          const float* mix_w0, const float* mix_b0, const float* mix_w2, const float* mix_b2,
                                          float* motion, int motion_pitch, int motion_off, float* scratch, int wy0, int wy1,
                                          int qkv_layout, void* stream) {
  return attention_impl(true, rc_closed_form, qkv, qkv_pitch, out, out_pitch, C, heads, g, cross, relative_coord, mix_w0, mix_b0, mix_w2,
                        mix_b2, motion, motion_pitch, motion_off, scratch, wy0, wy1, qkv_layout, stream);
}

extern "C" int atmvfi_window_attention(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                                       const atmvfi_window_geom* g, int cross, const float* relative_coord,
                                       const float* mix_w0, const float* mix_b0, const float* mix_w2, const float* mix_b2,
                                       float* motion, int motion_pitch, int motion_off, float* scratch, int wy0, int wy1,
                                       int qkv_layout, void* stream) {
  return attention_impl(false, 0, qkv, qkv_pitch, out, out_pitch, C, heads, g, cross, relative_coord, mix_w0, mix_b0, mix_w2, mix_b2,
                        motion, motion_pitch, motion_off, scratch, wy0, wy1, qkv_layout, stream);
}

static int attention_impl(bool tensor_cores, int rc_closed_form, const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                          const atmvfi_window_geom* g, int cross, const float* relative_coord, const float* mix_w0, const float* mix_b0,
                          const float* mix_w2, const float* mix_b2, float* motion, int motion_pitch, int motion_off, float* scratch,
                          int wy0, int wy1, int layout, void* stream) {
  ATMVFI_REQUIRE(heads > 0 && C % heads == 0, "window_attention: dim %d should be divided by num_heads %d", C, heads);
  const int hd = C / heads, N = g->ws * g->ws;
  ATMVFI_REQUIRE(N <= 256, "window_attention: window %d too large (max 16)", g->ws);
  ATMVFI_REQUIRE(g->Hp % g->ws == 0 && g->Wp % g->ws == 0 && g->shift >= 0 && g->shift < g->ws, "window_attention: bad geometry");
  ATMVFI_REQUIRE(!cross || g->B2 % 2 == 0, "window_attention: cross attention needs an even batch");
  ATMVFI_REQUIRE(qkv_pitch % 4 == 0 && out_pitch % 4 == 0 && hd % 4 == 0, "window_attention: pitches / head dim must be multiples of 4");
  ATMVFI_REQUIRE(layout == 0 || layout == 1, "window_attention: unknown qkv layout %d", layout);
  int nwy;
  ATMVFI_REQUIRE(row_window(g->Hp / g->ws, wy0, wy1, &wy0, &nwy), "window_attention: bad window-row range [%d,%d)", wy0, wy1);
  const bool want_motion = motion != nullptr;
  ATMVFI_REQUIRE(!want_motion || (relative_coord && scratch && mix_w0 && mix_b0 && mix_w2 && mix_b2 && cross),
                 "window_attention: motion output needs relative_coord, scratch and the head-mix MLP");
  cudaStream_t st = (cudaStream_t)stream;
  float* raw = want_motion ? scratch : nullptr;
  const float* rc = want_motion ? relative_coord : nullptr;
  int rcode = 3;
  if (tensor_cores)
    rcode = atmvfi_window_attention_tc_launch(qkv, qkv_pitch, out, out_pitch, C, heads, g, cross, (want_motion && !rc_closed_form) ? rc : nullptr, raw, wy0, nwy, layout, st);
  ATMVFI_REQUIRE(!(rcode == 3 && atmvfi_act_f16()), "window_attention: fp16 feature maps need the tcgen05 attention kernel, which does not "
                 "take this shape (window %d, head dim %d)", g->ws, hd);
  if (rcode == 3)      // shape outside the tcgen05 kernel's envelope (or fp32 requested): CUDA-core kernel
  switch (hd) {
    case 28: rcode = launch_attention<28>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    case 44: rcode = launch_attention<44>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    case 48: rcode = launch_attention<48>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    case 84: rcode = launch_attention<84>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    case 16: rcode = launch_attention<16>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    case 32: rcode = launch_attention<32>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    case 64: rcode = launch_attention<64>(qkv, qkv_pitch, out, out_pitch, C, heads, *g, cross, rc, raw, wy0, nwy, layout, st); break;
    default:
      atmvfi_set_error("window_attention: head dim %d not instantiated (have 16,28,32,44,48,64,84)", hd);
      return 2;
  }
  if (rcode) return rcode;
  if (want_motion) {
    int64_t rows = (int64_t)g->B2 * nwy * g->ws * g->Wp;
    int blocks = (int)((rows * 2 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks > 0) motion_mix_kernel<<<blocks, 256, 0, st>>>(raw, heads, *g, rows, wy0, nwy, mix_w0, mix_b0, mix_w2, mix_b2, motion, motion_pitch, motion_off);
    ATMVFI_CHECK_LAUNCH("motion_mix");
  }
  return 0;
}
