// FP32 (CUDA-core FFMA) implicit-GEMM convolution / linear layer.  This is the exact-fp32 datapath
// (precision = ATMVFI_FP32): it anchors parity against the CPU oracle and serves the thin layers
// (3-channel input, 3/5-channel heads) that cannot fill a tensor-core tile.
//
// GEMM view: M = B*Hout*Wout output pixels, N = Cout (x4 for the k2s2 transposed conv), K = taps * sum(C_src).
// A is gathered on the fly from up to 4 NHWC sources (virtual channel concat), zero outside the image.
// CTA tile 128 x 64, K step 16, 256 threads, 8 x 4 register tile per thread.
#include "gemm_epilogue.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, THREADS = 256;

struct SimtParams {
  int nsrc;
  const float* sptr[ATMVFI_MAX_SRC];
  int sC[ATMVFI_MAX_SRC];
  int spitch[ATMVFI_MAX_SRC];
  int Hin, Win, ksize, stride, dil, pad;
  int Hout, Wout;
  int row0, nrows;          // row window: GEMM rows enumerate [B][nrows][Wout], output row = row0 + local row
  int64_t M;
  int Ntot, K, Ctot;
  const float* weight;
  int ldw;
  EpiParams epi;
};

__global__ void __launch_bounds__(THREADS) gemm_conv_simt_kernel(const SimtParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // the 8 pixels this thread gathers for A
  const int a_kk = tid % BK;
  int a_base[8], a_iy[8], a_ix[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + tid / BK + 16 * i;
    if (m < p.M) {
      int ox = (int)(m % p.Wout);
      int64_t t = m / p.Wout;
      int oy = p.row0 + (int)(t % p.nrows);
      int b = (int)(t / p.nrows);
      a_base[i] = b * p.Hin * p.Win;
      a_iy[i] = oy * p.stride - p.pad;
      a_ix[i] = ox * p.stride - p.pad;
    } else {
      a_base[i] = 0;
      a_iy[i] = -(1 << 28);     // always out of bounds -> zero
      a_ix[i] = 0;
    }
  }
  const int b_kk = tid / 16, b_n = (tid % 16) * 4;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    {  // ---- A tile ----
      int k = k0 + a_kk;
      const float* sp = nullptr;
      int pitch = 0, dy = 0, dx = 0;
      if (k < p.K) {
        int tap = k / p.Ctot, c = k - tap * p.Ctot;
        int s = 0;
        while (s < p.nsrc - 1 && c >= p.sC[s]) { c -= p.sC[s]; ++s; }
        sp = p.sptr[s] + c;
        pitch = p.spitch[s];
        if (p.ksize == 3) { dy = (tap / 3) * p.dil; dx = (tap % 3) * p.dil; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int iy = a_iy[i] + dy, ix = a_ix[i] + dx;
        float v = 0.f;
        if (sp && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
          v = __ldg(sp + (int64_t)(a_base[i] + iy * p.Win + ix) * pitch);
        As[a_kk][tid / BK + 16 * i] = v;
      }
    }
    {  // ---- B tile ----
      int k = k0 + b_kk, n = n0 + b_n;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < p.K && n < p.ldw) w = __ldg(reinterpret_cast<const float4*>(p.weight + (int64_t)k * p.ldw + n));
      *reinterpret_cast<float4*>(&Bs[b_kk][b_n]) = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    {   // window-local GEMM row -> row of the full [B][Hout][Wout] grid
      const int ox = (int)(m % p.Wout);
      const int64_t t = m / p.Wout;
      m = ((t / p.nrows) * p.Hout + p.row0 + t % p.nrows) * p.Wout + ox;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.Ntot) continue;
      int q = 0, co = n;
      if (p.epi.out_mode == ATMVFI_OUT_SHUFFLE2) { q = n / p.epi.Cout; co = n - q * p.epi.Cout; }
      int64_t orow = epi_out_row(p.epi, m, q);
      if (orow >= 0) epi_store(p.epi, m, orow, co, acc[i][j]);
    }
  }
}

}  // namespace

int atmvfi_gemm_conv_simt(const atmvfi_gemm_conv_desc* d, cudaStream_t st) {
  SimtParams p;
  p.nsrc = d->nsrc;
  p.Ctot = 0;
  for (int s = 0; s < ATMVFI_MAX_SRC; ++s) {
    p.sptr[s] = s < d->nsrc ? d->src[s].ptr : nullptr;
    p.sC[s] = s < d->nsrc ? d->src[s].C : 0;
    p.spitch[s] = s < d->nsrc ? d->src[s].pitch : 0;
    p.Ctot += p.sC[s];
  }
  p.Hin = d->Hin; p.Win = d->Win; p.ksize = d->ksize; p.stride = d->stride; p.dil = d->dil;
  p.pad = d->dil * (d->ksize - 1) / 2;
  p.Hout = d->Hout; p.Wout = d->Wout;
  ATMVFI_REQUIRE(row_window(d->Hout, d->row_begin, d->row_end, &p.row0, &p.nrows), "gemm_conv(fp32): bad row window [%d,%d) for Hout=%d",
                 d->row_begin, d->row_end, d->Hout);
  p.M = (int64_t)d->B * p.nrows * d->Wout;
  p.Ntot = d->out_mode == ATMVFI_OUT_SHUFFLE2 ? 4 * d->Cout : d->Cout;
  p.K = d->ksize * d->ksize * p.Ctot;
  p.weight = d->weight;
  p.ldw = d->ldw;
  p.epi = make_epi(d);
  ATMVFI_REQUIRE(d->ldw % 4 == 0 && d->ldw >= p.Ntot, "gemm_conv(fp32): ldw=%d must be a multiple of 4 and >= N=%d", d->ldw, p.Ntot);
  if (p.M <= 0) return 0;
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.Ntot + BN - 1) / BN));
  gemm_conv_simt_kernel<<<grid, THREADS, 0, st>>>(p);
  ATMVFI_CHECK_LAUNCH("gemm_conv(fp32)");
  return 0;
}
