// Depth-wise 3x3 + bias + GELU (Mlp middle, attention.py:74-85, 118-119) as a TMA-fed streaming kernel.
//
// The register-ring kernel in elementwise.cu tops out at ~3.2 TB/s (ncu: latency-bound, 52 % of the stalls on the first use
// of the next row; more prefetch depth costs registers and occupancy).  Here the loads leave the register file altogether:
//   * the NHWC map is a 4-D TMA tensor {C, W, H, B}; one box {64 channels, 18 columns, 1 row} = 4.6 KB is one pipeline
//     stage; TMA zero-fills columns -1 / W and rows -1 / H, which IS the convolution's zero padding, and channels beyond C;
//   * a CTA (256 threads = 16 columns x 16 channel quads) walks down a strip of rows; a ring of kStages boxes is kept in
//     flight by one elected thread (full / empty mbarriers), so ~3 CTAs x 6 rows x 4.6 KB per SM are always outstanding;
//   * every input row is read from shared memory once (3 conflict-free LDS.128 per thread) and SCATTERED into the
//     accumulators of the three output rows it feeds, exactly like the register kernel, so the two produce identical bits.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int kCc = 64;            // channels per CTA (256 B per pixel in the box)
constexpr int kX = 16;             // output columns per CTA
constexpr int kStages = 6;
constexpr int kRowsPerCta = 34;    // output rows per strip
constexpr int kBoxBytes = kCc * 4 * (kX + 2);       // fp32 maps; fp16 maps fill half of every slot
constexpr int kThreadsD = 256;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float gelu_erff(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
struct DwParams {
  CUtensorMap map;
  void* out;
  const float* w9c;
  const float* bias;
  int B, H, W, C, pitch;
  int y0, y1;                      // row window
  int cblocks, xblocks, strips;
};

template <bool kFastErf, typename T = float>
__global__ void __launch_bounds__(kThreadsD) dwconv_tma_kernel(const __grid_constant__ DwParams p, const bool rnd) {
  constexpr int kEs = (int)sizeof(T), kBox = kCc * kEs * (kX + 2);
  __shared__ __align__(128) uint8_t ring[kStages][kBoxBytes];
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages];

  int item = blockIdx.x;
  const int cb = item % p.cblocks; item /= p.cblocks;
  const int xb = item % p.xblocks; item /= p.xblocks;
  const int sb = item % p.strips;
  const int b = item / p.strips;
  const int c0 = cb * kCc, x0 = xb * kX;
  const int ys = p.y0 + sb * kRowsPerCta, ye = min(ys + kRowsPerCta, p.y1);
  const int nin = ye - ys + 2;                       // input rows ys-1 .. ye

  const int tid = threadIdx.x, lane = tid & 31;
  const int c4 = tid & 15, xl = tid >> 4;            // a warp = 2 columns x 16 channel quads: 2 x 256 contiguous bytes
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&empty[s])), "r"(kThreadsD / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int i) {                          // input row i of the strip -> slot i % kStages (thread 0 only)
    const int s = i % kStages;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full[s])), "r"(kBox) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     su32(ring[s])),
                 "l"(&p.map), "r"(su32(&full[s])), "r"(c0), "r"(x0 - 1), "r"(ys - 1 + i), "r"(b)
                 : "memory");
  };
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}\n" : "=r"(ok) : "r"(su32(bar)), "r"(parity) : "memory");
  };
  if (tid == 0)
    for (int i = 0; i < kStages && i < nin; ++i) issue(i);

  const int x = x0 + xl, c = c0 + 4 * c4;
  const bool active = x < p.W && c < p.C;
  float4 k[9], bz = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = active ? __ldg(reinterpret_cast<const float4*>(p.w9c + t * p.C + c)) : bz;
  if (active) bz = __ldg(reinterpret_cast<const float4*>(p.bias + c));
  T* obase = reinterpret_cast<T*>(p.out) + ((size_t)b * p.H * p.W + x) * p.pitch + c;
  const size_t row_stride = (size_t)p.W * p.pitch;

  auto fma4 = [](const float4& a, const float4& w, float4& acc) {
    acc.x = fmaf(a.x, w.x, acc.x); acc.y = fmaf(a.y, w.y, acc.y); acc.z = fmaf(a.z, w.z, acc.z); acc.w = fmaf(a.w, w.w, acc.w);
  };
  float4 acc0 = bz, acc1 = bz, acc2 = bz;            // output rows yy-1, yy, yy+1 while input row yy is being consumed
  for (int i = 0; i < nin; ++i) {
    const int s = i % kStages;
    wait(&full[s], (uint32_t)(i / kStages) & 1);
    const uint8_t* rowp = ring[s] + xl * (kCc * kEs) + c4 * (4 * kEs);
    const float4 l = Act<T>::lds4(rowp);
    const float4 m = Act<T>::lds4(rowp + kCc * kEs);
    const float4 r = Act<T>::lds4(rowp + 2 * kCc * kEs);
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&empty[s])) : "memory");
    if (tid == 0 && i + kStages < nin) {             // refill this slot once all 8 warps have read it
      wait(&empty[s], (uint32_t)(i / kStages) & 1);
      issue(i + kStages);
    }
    // taps arrive in row-major order for every output row: bias, then the top row (taps 0-2), the middle row, the bottom row
    fma4(l, k[6], acc0); fma4(m, k[7], acc0); fma4(r, k[8], acc0);
    fma4(l, k[3], acc1); fma4(m, k[4], acc1); fma4(r, k[5], acc1);
    fma4(l, k[0], acc2); fma4(m, k[1], acc2); fma4(r, k[2], acc2);
    if (i >= 2 && active) {
      float4 o;
      if (kFastErf) o = make_float4(gelu_fast(acc0.x), gelu_fast(acc0.y), gelu_fast(acc0.z), gelu_fast(acc0.w));
      else o = make_float4(gelu_erff(acc0.x), gelu_erff(acc0.y), gelu_erff(acc0.z), gelu_erff(acc0.w));
      Act<T>::st4(obase + (size_t)(ys + i - 2) * row_stride, round_tf32_if(o, rnd));
    }
    acc0 = acc1; acc1 = acc2; acc2 = bz;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn dw_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// Returns 0 on success, 3 when this kernel does not apply (the caller falls back to the register-ring kernel).
int atmvfi_dwconv_tma_launch(const void* in, void* out, int B, int H, int W, int C, int pitch, const float* w9c,
                             const float* bias, int y0, int ny, bool rnd, bool f16, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) { const char* ev = getenv("ATMVFI_DW_TMA"); enabled = ev ? atoi(ev) : 1; }
  if (!enabled) return 3;
  if (C % 4 || pitch % 4 || ((uintptr_t)in & 15) || ((uintptr_t)out & 15) || ((uintptr_t)w9c & 15) || ((uintptr_t)bias & 15)) return 3;
  EncodeTiledFn enc = dw_get_encode();
  if (!enc) return 3;
  DwParams p;
  memset(&p, 0, sizeof(p));
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t es = f16 ? 2 : 4;
  if (f16 && pitch % 8) return 3;                 // TMA strides must be multiples of 16 bytes
  cuuint64_t gstr[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * W, (cuuint64_t)pitch * es * W * H};
  cuuint32_t box[4] = {(cuuint32_t)kCc, (cuuint32_t)(kX + 2), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&p.map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 3;
  p.out = out; p.w9c = w9c; p.bias = bias;
  p.B = B; p.H = H; p.W = W; p.C = C; p.pitch = pitch;
  p.y0 = y0; p.y1 = y0 + ny;
  p.cblocks = (C + kCc - 1) / kCc;
  p.xblocks = (W + kX - 1) / kX;
  p.strips = (ny + kRowsPerCta - 1) / kRowsPerCta;
  const int64_t ctas = (int64_t)p.cblocks * p.xblocks * p.strips * B;
  if (ctas <= 0) return 0;
  if (ctas > 0x7fffffff) return 3;
  if (f16) dwconv_tma_kernel<true, __half><<<(unsigned)ctas, kThreadsD, 0, st>>>(p, false);      // fp16 storage: the fast erf's 2.6e-7 is far below half precision
  else if (rnd) dwconv_tma_kernel<true><<<(unsigned)ctas, kThreadsD, 0, st>>>(p, rnd);
  else dwconv_tma_kernel<false><<<(unsigned)ctas, kThreadsD, 0, st>>>(p, rnd);
  ATMVFI_CHECK_LAUNCH("dwconv3x3_gelu(tma)");
  return 0;
}
