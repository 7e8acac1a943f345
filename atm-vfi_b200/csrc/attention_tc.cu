// Window attention + attention-to-motion on the 5th-generation tensor cores (attention.py:187-213, 370-390).
//
// One CTA (128 threads) per work item = (window group, head, 128-row query tile):
//   * windows of N <= 64 tokens are processed 128/N at a time (a block-diagonal 128 x 128 problem, the
//     off-diagonal blocks are excluded in the softmax), larger windows (N = 144 for the global branch) one
//     at a time in query tiles of 128 rows;
//   * Q, K (K-major) and V^T are staged in shared memory in the SWIZZLE_128B layout, rounded to TF32;
//   * S = Q K^T   : tcgen05.mma kind::tf32, M = 128, N = keys (multiple of 16), accumulator in TMEM;
//   * softmax     : thread r owns row r = TMEM lane r: scale, additive -100 masks evaluated from the window
//                   geometry, max, exp, sum; the motion expectation sum_j p_ij * relative_coord[:, i, j] is
//                   accumulated in the same pass; P (unnormalised, TF32) overwrites the Q/K staging area;
//   * O = P V     : second tcgen05.mma chain into other TMEM columns; rows are scaled by 1/l on the way out.
// The reference materialises [B',8,N,N] attention and a [B',8,2,N,N] product for the motion; here nothing but
// the per-head outputs and two floats per (token, head) ever reach HBM.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int kMaxParts = 4;       // kP threads per query row split the 16-key chunks (2 or 4); warps w, w+4, ... share TMEM lanes

struct AttnParams {
  const float* qkv;
  int qkv_pitch;
  float* out;                   // fp32, or __half when out_f16 (precision ATMVFI_F16: the attention output feeds an fp16 GEMM)
  int out_pitch;
  int out_f16;
  int C, heads, hd;
  atmvfi_window_geom g;
  int cross;
  const float* rc;            // [2][N][N] or nullptr (closed form: key position - query position)
  float* motion_raw;          // [rows][heads][2] or nullptr
  unsigned long long* prof;   // debug: 6 phase cycle counters (stage, QK^T, softmax max, softmax exp, PV, output) or nullptr
  int N, wpi, mtiles, KP, HP, chunksH, chunksK;
  int total_windows, nW;
  int wy0, per_img, virt_windows;   // row window: windows [wy0*nwx, +per_img) of every image, virt_windows = B2*per_img
  int tmem_cols;
  int round;
  float scale;
  // shared memory carve-up (bytes)
  int offK, offV, offLab, offBar;
  // layout 1 (ATMVFI_OUT_QKV_HEADS): Q[h][r][d], K[h][r][d], V^T[h][d][r] - operands are fetched by TMA straight into the swizzled
  // UMMA layouts: mapQ / mapK = 3-D {hd, R, 2*heads} (box 32 x rows x 1), mapV = 2-D {R, C} (box 32 keys x HP channels)
  int layout, qrows;
  CUtensorMap mapQ, mapK, mapV;
  // output through shared memory + TMA store: mapO = 3-D {hd, heads, R} over `out`, box {hd + o_pad, 1, o_rows}; the box is wider than a
  // head on purpose - the staging rows are padded to a conflict-free pitch and the TMA unit clips the columns beyond hd
  int o_tma, o_pad, o_rows;
  CUtensorMap mapO;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t swz(int row, int c) {   // byte offset of float column c (0..31) of a 128-byte row
  return (uint32_t)(row * 128 + ((((c >> 2) ^ (row & 7)) << 4) | ((c & 3) << 2)));
}
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(s_u32(bar)), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {   // no wait: several loads may be in flight
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tf32r(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// kHD > 0: head dimension known at compile time - the Q / K / V staging loops unroll and their global loads are issued
// back to back (with a run-time bound every 16-byte load waited for the previous one: ~18k of the ~46k cycles of a CTA).
template <int kHD, int kP>
__global__ void __launch_bounds__(128 * kP) window_attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                         // chunksH x [128 x 128 B]; later P: chunksK x [128 x 128 B]
  uint8_t* sK = smem + p.offK;                // chunksH x [KP x 128 B]
  uint8_t* sV = smem + p.offV;                // chunksK x [HP x 128 B]   (V transposed: row = channel, column = key)
  int* sLab = reinterpret_cast<int*>(smem + p.offLab);           // [KP] mask label per key
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  uint64_t* tbar = bars + 4;                  // TMA completion (head-major layout)

  constexpr int kThreadsA = kRows * kP;
  const int tid = threadIdx.x & (kRows - 1), part = threadIdx.x >> 7, warp = (threadIdx.x >> 5) & 3;
  float* sRed = reinterpret_cast<float*>(smem + p.offBar + 64);      // [4][2][128] partial max / sum / motion x / motion y
  // work item
  int item = blockIdx.x;
  const int mt = item % p.mtiles;
  item /= p.mtiles;
  const int h = item % p.heads;
  const int wg = item / p.heads;
  const int64_t total_win = p.total_windows;
  const int64_t win0 = (int64_t)wg * p.wpi;

  pdl_wait();                                 // programmatic dependent launch (common.cuh): q | k | v come from the previous kernel
  const int N = p.N, hd = kHD ? kHD : p.hd;
  const int HP = kHD ? (kHD + 15) / 16 * 16 : p.HP;
  const int nwx = p.g.Wp / p.g.ws;
  // virtual window index v (dense over the row window) -> window id in the full window-major tensor
  auto actual = [&](int64_t v) -> int64_t { const int vi = (int)v, bi = vi / p.per_img; return (int64_t)bi * p.nW + p.wy0 * nwx + (vi - bi * p.per_img); };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(tbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.layout == 1) {
      // Head-major operands: this thread issues a handful of TMA boxes per window right away - they land in the swizzled K-major
      // layouts of the two MMAs (Q, K: rows = tokens; V^T: rows = channels) while the CTA builds its row / key tables.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const int nkc = (N + 31) >> 5;                               // 32-key chunks of V^T per window
      uint32_t bytes = 0;
      for (int wl = 0; wl < p.wpi; ++wl)
        if (win0 + wl < p.virt_windows) bytes += (uint32_t)(p.chunksH * p.qrows * 128 + p.chunksH * N * 128 + nkc * HP * 128);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(tbar)), "r"(bytes) : "memory");
      for (int wl = 0; wl < p.wpi; ++wl) {
        if (win0 + wl >= p.virt_windows) continue;
        const int64_t w = actual(win0 + wl);
        const int64_t wk = p.cross ? (w + total_win / 2) % total_win : w;       // the other frame's copy of the window (attention.py:318)
        const int qrow0 = (int)(w * N) + mt * kRows, krow0 = (int)(wk * N);
        for (int c = 0; c < p.chunksH; ++c) {
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                           s_u32(sQ + c * (kRows * 128) + wl * N * 128)),
                       "l"(&p.mapQ), "r"(s_u32(tbar)), "r"(c * 32), "r"(qrow0), "r"(h)
                       : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                           s_u32(sK + c * (p.KP * 128) + wl * N * 128)),
                       "l"(&p.mapK), "r"(s_u32(tbar)), "r"(c * 32), "r"(krow0), "r"(p.heads + h)
                       : "memory");
        }
        for (int kc = 0; kc < nkc; ++kc)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                           s_u32(sV + (((wl * N) >> 5) + kc) * (HP * 128))),
                       "l"(&p.mapV), "r"(s_u32(tbar)), "r"(krow0 + kc * 32), "r"(h * hd)
                       : "memory");
      }
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  long long t_prev = p.prof ? clock64() : 0;      // phase timers (debug: ATMVFI_ATTN_PROF)
  // ---- row / key bookkeeping (one row per thread: no per-element index arithmetic) ----------------------------
  const int nslH = (hd + 7) >> 3;             // K = 8 slices of the head dimension
  const int hd_pad = nslH * 8;
  const bool masked = (p.g.shift != 0) || p.g.Hp != p.g.H || p.g.Wp != p.g.W;
  const int ws = p.g.ws;
  // (window-local index, token) of row / key `r`; window < 0: padding row of the tile
  auto locate = [&](int r, int base_tok, int& wl, int& tok) -> int64_t {
    if (p.wpi == 1) { wl = 0; tok = base_tok + r; return tok < N ? actual(win0) : -1; }
    wl = r / N;
    tok = r - wl * N;
    return (wl < p.wpi && win0 + wl < p.virt_windows) ? actual(win0 + wl) : -1;
  };
  auto mask_label = [&](int64_t w, int tok) -> int {
    if (!masked) return 0;
    const int wi = (int)(w % p.nW);
    return win_mask_label(p.g, (wi / nwx) * ws + tok / ws, (wi % nwx) * ws + tok % ws);
  };

  // ---- stage Q, K (K-major) and V^T, TF32-rounded, swizzled ---------------------------------------------------------
  // Row / key bookkeeping goes to shared memory first; the copy loops then walk (row, 16-byte column) pairs with the
  // column fastest, so a warp reads whole 192-byte (hd = 48) head slices instead of 32 scattered 16-byte pieces
  // (measured: the thread-per-row version spent 16.5k of the 35k cycles of a CTA here, bound by L1 tag lookups).
  int wl_i, tok_i;
  const int64_t win_i = locate(tid, mt * kRows, wl_i, tok_i);
  const bool row_ok = win_i >= 0;
  int* sRowSrc = reinterpret_cast<int*>(sRed + 4 * kMaxParts * kRows);      // [128] global row of each query row, -1 = padding row
  int* sKeySrc = sRowSrc + kRows;                               // [KP]  global row of each key (other frame's window if cross)
  float* sKx = reinterpret_cast<float*>(sKeySrc + 256);         // [KP]  window coordinates of each key as floats (closed-form motion)
  float* sKy = sKx + 256;
  if (part == 0) sRowSrc[tid] = row_ok ? (int)(win_i * N + tok_i) : -1;
  for (int kk = threadIdx.x; kk < p.KP; kk += kThreadsA) {
    int wl, tok;
    const int64_t w = locate(kk, 0, wl, tok);
    const bool ok = w >= 0;
    int src = -1;
    if (ok) {
      const int64_t wk = p.cross ? (w + total_win / 2) % total_win : w;       // the other frame's copy of the window (attention.py:318)
      src = (int)(wk * N + tok);
    }
    sKeySrc[kk] = src;
    // key meta: bits [0,12) mask label, [12,16) window-local index, [16,24) x, [24,32) y; -1 = excluded key
    sLab[kk] = ok ? (mask_label(w, tok) | (wl << 12) | ((tok % ws) << 16) | ((tok / ws) << 24)) : -1;
    sKx[kk] = (float)(tok % ws);
    sKy[kk] = (float)(tok / ws);
  }
  // row constants of the softmax, computed while the operands are in flight
  const int lab_i = row_ok ? (mask_label(win_i, tok_i) | (wl_i << 12)) : 0;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const float sc2 = p.scale * 1.4426950408889634f;              // logits are kept in the log2 domain: exp(x) = exp2(x*log2e)
  const float mask2 = -100.0f * 1.4426950408889634f;
  auto logit2 = [&](float s, int meta) -> float {               // -inf: excluded; else (scaled logit + additive mask) * log2(e)
    if (meta < 0 || ((meta ^ lab_i) & 0xF000)) return -INFINITY; // padding key, or a key of another window of the group
    float x = s * sc2;
    if ((meta ^ lab_i) & 0xFFF) x += mask2;
    return x;
  };
  // the two threads of a row take alternate 16-key chunks; chunks that belong entirely to another window of the group
  // (N % 16 == 0) are never read: their probabilities are exactly zero
  // (decided per WARP - tcgen05.ld is .aligned - from the first and last row of the warp; rows are window-ordered)
  const bool skip_foreign = p.wpi > 1;
  const int own_lo = __shfl_sync(0xffffffffu, wl_i * N, 0), own_hi = __shfl_sync(0xffffffffu, wl_i * N + N, 31);
  // this thread's chunks: c0 = first + 32 j < lim.  With at most two of them (8x8 windows: 64 keys per row, two threads per
  // row) S stays in registers between the two passes and both tcgen05.ld are in flight together; larger windows re-read
  // S from TMEM in the second pass (keeping 5 chunks live cost more in register pressure than the reload).
  constexpr int kMaxOwn = 2;
  constexpr int kStep = 16 * kP;                                // distance between two chunks of one thread
  int first = part * 16, lim = p.KP;
  if (skip_foreign) {
    int ci = own_lo >> 4;
    ci += (part - ci) & (kP - 1);                     // first chunk index >= ci that is congruent to part
    first = ci * 16;
    lim = min(own_hi, p.KP);
  }
  const bool keep = first + kStep * kMaxOwn >= lim;             // warp-uniform
  const int xi = tok_i % ws, yi = tok_i / ws;
  __syncthreads();
  if (p.layout == 1) {
    // (operands were requested by thread 0 at the top of the kernel)
    const int nkc = (N + 31) >> 5;
    for (int wl = 0; wl < p.wpi; ++wl) {                         // windows of the group that do not exist: zero their V^T columns
      if (win0 + wl < p.virt_windows) continue;                  // (P is zero there, but 0 x garbage could be NaN)
      uint8_t* z = sV + ((wl * N) >> 5) * (HP * 128);
      for (int i = threadIdx.x * 16; i < nkc * HP * 128; i += kThreadsA * 16) *reinterpret_cast<float4*>(z + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    bar_wait(tbar, 0);
  } else {
    // Asynchronous copies (cp.async): every load of the CTA is in flight at once and no registers are tied up.  The
    // head slices are 192-byte pieces of 4.6 KB rows, so DRAM latency under this access pattern is long (measured:
    // 4-6 dependent round trips of ~4k cycles each with register staging); here it is paid once.  Operands are NOT
    // re-rounded: the runtime's producer (the qkv GEMM) already stores TF32 values; a caller that passes unrounded
    // fp32 gets the tensor core's truncation instead of round-to-nearest.
    const int F4 = hd >> 2, F4P = hd_pad >> 2;                  // float4 per row: real / padded to the K = 8 slices of the MMA
    const float* qbase = p.qkv + h * hd;
    const float* kbase = qbase + p.C;
    const float* vbase = kbase + p.C;
    auto cp16 = [](uint32_t dst, const void* src, bool valid) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
    };
    auto cp4 = [](uint32_t dst, const void* src, bool valid) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
    };
    const uint32_t aQ = s_u32(sQ), aK = s_u32(sK), aV = s_u32(sV);
    for (int i = threadIdx.x; i < kRows * F4P; i += kThreadsA) {
      const int r = i / F4P, c4 = i - r * F4P, c = c4 << 2;
      const int src = sRowSrc[r];
      const bool ok = src >= 0 && c4 < F4;
      cp16(aQ + (c >> 5) * (kRows * 128) + swz(r, c & 31), ok ? qbase + (int64_t)src * p.qkv_pitch + c : p.qkv, ok);
    }
    for (int i = threadIdx.x; i < p.KP * F4P; i += kThreadsA) {
      const int kk = i / F4P, c4 = i - kk * F4P, c = c4 << 2;
      const int src = sKeySrc[kk];
      const bool ok = src >= 0 && c4 < F4;
      cp16(aK + (c >> 5) * (p.KP * 128) + swz(kk, c & 31), ok ? kbase + (int64_t)src * p.qkv_pitch + c : p.qkv, ok);
    }
    // V^T (row = channel, column = key): 4-byte copies, lanes along the channels of one key (coalesced reads).  Columns
    // of excluded keys are zero-filled (P is zero there, but 0 x garbage could be NaN); rows hd.. feed output columns
    // that are never stored.
    for (int i = threadIdx.x; i < p.KP * hd; i += kThreadsA) {
      const int kk = i / hd, c = i - kk * hd;
      const int src = sKeySrc[kk];
      const bool ok = src >= 0;
      cp4(aV + (kk >> 5) * (HP * 128) + swz(c, kk & 31), ok ? vbase + (int64_t)src * p.qkv_pitch + c : p.qkv, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&p.prof[0], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + p.KP;

  // ---- S = Q K^T ------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    const uint32_t id = idesc_tf32(p.KP);
    for (int s = 0; s < nslH; ++s) {
      const int c = s >> 2, j = s & 3;
      mma_tf32(tS, sdesc(s_u32(sQ + c * (kRows * 128))) + 2 * j, sdesc(s_u32(sK + c * (p.KP * 128))) + 2 * j, id, s ? 1u : 0u);
    }
    commit(&bars[0]);
  }
  bar_wait(&bars[0], 0);
  if (p.prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&p.prof[1], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- softmax + motion, row per thread ---------------------------------------------------------------------
  uint32_t sraw[kMaxOwn][16];
  float mx = -INFINITY;
  float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};       // four independent chains
  if (keep) {
#pragma unroll
    for (int j = 0; j < kMaxOwn; ++j)
      if (first + kStep * j < lim) ld16_issue(tS + lane_addr + first + kStep * j, sraw[j]);
    ld_wait();
#pragma unroll
    for (int j = 0; j < kMaxOwn; ++j) {
      const int c0 = first + kStep * j;
      if (c0 < lim) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int4 m4 = *reinterpret_cast<const int4*>(sLab + c0 + e);          // one broadcast read for four keys
          const int mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float x = logit2(__uint_as_float(sraw[j][e + k]), mm[k]);
            sraw[j][e + k] = __float_as_uint(x);
            mx4[k] = fmaxf(mx4[k], x);
          }
        }
      }
    }
  } else {
    // larger windows: two chunks in flight per step, labels four at a time
    for (int c0 = first; c0 < lim; c0 += 2 * kStep) {
      uint32_t ra[16], rb[16];
      const bool two = c0 + kStep < lim;
      ld16_issue(tS + lane_addr + c0, ra);
      if (two) ld16_issue(tS + lane_addr + c0 + kStep, rb);
      ld_wait();
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const int4 m4 = *reinterpret_cast<const int4*>(sLab + c0 + e);
        const int mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) mx4[k] = fmaxf(mx4[k], logit2(__uint_as_float(ra[e + k]), mm[k]));
      }
      if (two) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int4 m4 = *reinterpret_cast<const int4*>(sLab + c0 + kStep + e);
          const int mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) mx4[k] = fmaxf(mx4[k], logit2(__uint_as_float(rb[e + k]), mm[k]));
        }
      }
    }
  }
  mx = fmaxf(fmaxf(mx, mx4[0]), fmaxf(fmaxf(mx4[1], mx4[2]), mx4[3]));
  sRed[part * kRows + tid] = mx;
  __syncthreads();
  if (p.prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&p.prof[2], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
  mx = sRed[tid];
#pragma unroll
  for (int k = 1; k < kP; ++k) mx = fmaxf(mx, sRed[k * kRows + tid]);
  // P may overwrite the Q/K area now: the MMA that read it has retired (bars[0]) and S lives in TMEM / registers
  float l = 0.f, mvx = 0.f, mvy = 0.f;
  const bool want_motion = p.motion_raw != nullptr;
  // chunks of this thread that are not its own (another window of the group, or beyond KP): P = 0
  for (int c0 = part * 16; c0 < p.chunksK * 32; c0 += kStep) {
    if (c0 >= first && c0 < lim) continue;
    uint8_t* prow = sQ + (c0 >> 5) * (kRows * 128);
#pragma unroll
    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(prow + swz(tid, (c0 & 31) + e)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float mxs = mx > -INFINITY ? mx : 0.f;                   // (a row without any admissible key: 2^(-inf - 0) = 0, not NaN)
  auto emit_chunk = [&](const int c0, const float (&x16)[16]) {      // x16: masked logits (log2 domain) of 16 keys
    uint8_t* prow = sQ + (c0 >> 5) * (kRows * 128);
    float pv[16];
    if (want_motion && p.rc) {          // relative_coord table given explicitly (not the reference's own buffer): per-entry loads
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int meta = sLab[c0 + e];
        const float x = x16[e];
        const float pe = (row_ok && x > -INFINITY) ? exp2f(x - mx) : 0.f;
        l += pe;
        const int tj = ((meta >> 24) & 0xFF) * ws + ((meta >> 16) & 0xFF);
        const float dx = pe != 0.f ? __ldg(p.rc + (int64_t)tok_i * N + tj) : 0.f;
        const float dy = pe != 0.f ? __ldg(p.rc + (int64_t)N * N + (int64_t)tok_i * N + tj) : 0.f;
        mvx = fmaf(pe, dx, mvx);
        mvy = fmaf(pe, dy, mvy);
        pv[e] = tf32r(pe);
      }
    } else {
      // closed form: sum_j p_j (k_j - q) = sum_j p_j k_j - q l; the key coordinates come from shared memory four at a time
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const float4 kx4 = *reinterpret_cast<const float4*>(sKx + c0 + e), ky4 = *reinterpret_cast<const float4*>(sKy + c0 + e);
        const float kx[4] = {kx4.x, kx4.y, kx4.z, kx4.w}, ky[4] = {ky4.x, ky4.y, ky4.z, ky4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float pe;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(x16[e + k] - mxs));
          pe = row_ok ? pe : 0.f;                                                            // padding rows of the tile contribute nothing
          l += pe;
          mvx = fmaf(pe, kx[k], mvx);
          mvy = fmaf(pe, ky[k], mvy);
          pv[e + k] = __uint_as_float((__float_as_uint(pe) + 0x1000u) & 0xFFFFE000u);      // cvt.rna.tf32 of a finite non-negative value
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 16; e += 4)
      *reinterpret_cast<float4*>(prow + swz(tid, (c0 & 31) + e)) = make_float4(pv[e], pv[e + 1], pv[e + 2], pv[e + 3]);
  };
  if (keep) {
#pragma unroll
    for (int j = 0; j < kMaxOwn; ++j) {
      const int c0 = first + kStep * j;
      if (c0 < lim) {
        float x16[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) x16[e] = __uint_as_float(sraw[j][e]);
        emit_chunk(c0, x16);
      }
    }
  } else {
    uint32_t rn[16];
    ld16_issue(tS + lane_addr + first, rn);
    for (int c0 = first; c0 < lim; c0 += kStep) {
      float x16[16];
      ld_wait();
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const int4 m4 = *reinterpret_cast<const int4*>(sLab + c0 + e);
        const int mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) x16[e + k] = logit2(__uint_as_float(rn[e + k]), mm[k]);
      }
      if (c0 + kStep < lim) ld16_issue(tS + lane_addr + c0 + kStep, rn);        // next chunk's accumulators while this one is exponentiated
      emit_chunk(c0, x16);
    }
  }
  if (!p.rc) { mvx = fmaf(-(float)xi, l, mvx); mvy = fmaf(-(float)yi, l, mvy); }
  sRed[(kMaxParts + part) * kRows + tid] = l;
  sRed[(2 * kMaxParts + part) * kRows + tid] = mvx;
  sRed[(3 * kMaxParts + part) * kRows + tid] = mvy;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&p.prof[3], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- O = P V ------------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    const uint32_t id = idesc_tf32(HP);
    const int nslK = p.KP >> 3;
    for (int s = 0; s < nslK; ++s) {
      const int c = s >> 2, j = s & 3;
      mma_tf32(tO, sdesc(s_u32(sQ + c * (kRows * 128))) + 2 * j, sdesc(s_u32(sV + c * (HP * 128))) + 2 * j, id, s ? 1u : 0u);
    }
    commit(&bars[1]);
  }
  bar_wait(&bars[1], 0);
  if (p.prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&p.prof[4], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  l = 0.f; mvx = 0.f; mvy = 0.f;
#pragma unroll
  for (int k = 0; k < kP; ++k) {
    l += sRed[(kMaxParts + k) * kRows + tid];
    mvx += sRed[(2 * kMaxParts + k) * kRows + tid];
    mvy += sRed[(3 * kMaxParts + k) * kRows + tid];
  }
  const float inv = row_ok ? 1.0f / l : 0.f;
  const int64_t grow = row_ok ? win_i * N + tok_i : 0;
  if (p.o_tma) {
    // Output through shared memory: every thread drops its 16-channel pieces into a [128][hd + pad] tile in the (now free) Q / P
    // region - the pad makes the row pitch an odd number of 16-byte units, so the stores of a quarter-warp hit different banks -
    // and thread 0 hands one box per window (or per 16 rows of a large window) to the TMA unit, which clips the pad columns.
    const int es = p.out_f16 ? 2 : 4;
    const int pitch_b = (hd + p.o_pad) * es;
    uint8_t* so = sQ + tid * pitch_b;
#pragma unroll
    for (int c0 = part * 16; c0 < HP; c0 += kStep) {
      float o[16];
      ld16(tO + lane_addr + c0, o);
#pragma unroll
      for (int e = 0; e < 16; e += 8) {
        if (c0 + e < hd) {
          const float4 v0 = make_float4(o[e] * inv, o[e + 1] * inv, o[e + 2] * inv, o[e + 3] * inv);
          const float4 v1 = make_float4(o[e + 4] * inv, o[e + 5] * inv, o[e + 6] * inv, o[e + 7] * inv);
          if (p.out_f16) {
            const uint2 a = Act<__half>::pack(v0), b = Act<__half>::pack(v1);
            if (c0 + e + 4 < hd) *reinterpret_cast<uint4*>(so + (c0 + e) * 2) = make_uint4(a.x, a.y, b.x, b.y);
            else *reinterpret_cast<uint2*>(so + (c0 + e) * 2) = a;
          } else {
            *reinterpret_cast<float4*>(so + (c0 + e) * 4) = round_tf32_if(v0, p.round != 0);
            if (c0 + e + 4 < hd) *reinterpret_cast<float4*>(so + (c0 + e + 4) * 4) = round_tf32_if(v1, p.round != 0);
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int wl = 0; wl < p.wpi; ++wl) {
        if (win0 + wl >= p.virt_windows) continue;
        const int64_t w = actual(win0 + wl);
        const int nrows = p.wpi > 1 ? N : min(kRows, N - mt * kRows);           // real rows of this window in the tile
        const int row_g = (int)(w * N) + mt * kRows, row_s = p.wpi > 1 ? wl * N : 0;
        for (int r0 = 0; r0 < nrows; r0 += p.o_rows)
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&p.mapO),
                       "r"(s_u32(sQ + (row_s + r0) * pitch_b)), "r"(0), "r"(h), "r"(row_g + r0)
                       : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else {
#pragma unroll
  for (int c0 = part * 16; c0 < HP; c0 += kStep) {
    float o[16];
    ld16(tO + lane_addr + c0, o);
    if (row_ok) {
      const int64_t off = grow * p.out_pitch + h * hd + c0;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        if (c0 + e < hd) {
          float4 v = make_float4(o[e] * inv, o[e + 1] * inv, o[e + 2] * inv, o[e + 3] * inv);
          if (p.out_f16) Act<__half>::st4(reinterpret_cast<__half*>(p.out) + off + e, v);
          else *reinterpret_cast<float4*>(p.out + off + e) = round_tf32_if(v, p.round != 0);
        }
      }
    }
  }
  }
  if (row_ok && p.motion_raw && part == 0) {
    p.motion_raw[(grow * p.heads + h) * 2 + 0] = mvx * inv;
    p.motion_raw[(grow * p.heads + h) * 2 + 1] = mvy * inv;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&p.prof[5], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
  }
  if (p.o_tma && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the staging tile lives until the boxes have left it
}

inline int rup(int x, int m) { return (x + m - 1) / m * m; }

unsigned long long* g_prof = nullptr;

}  // namespace

// Debug aid: ATMVFI_ATTN_PROF=1 makes thread 0 of every CTA accumulate the cycles of the six phases of the kernel into a
// device buffer (read back with atmvfi_attn_prof_read).
static unsigned long long* atmvfi_attn_prof_buffer() {
  static int on = -1;
  if (on < 0) {
    const char* ev = getenv("ATMVFI_ATTN_PROF");
    on = ev && atoi(ev) ? 1 : 0;
    if (on) { cudaMalloc(&g_prof, 8 * sizeof(unsigned long long)); cudaMemset(g_prof, 0, 8 * sizeof(unsigned long long)); }
  }
  return on ? g_prof : nullptr;
}
extern "C" int atmvfi_attn_prof_read(unsigned long long* out6) {
  if (!g_prof) return 1;
  cudaMemcpy(out6, g_prof, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(g_prof, 0, 8 * sizeof(unsigned long long));
  return 0;
}

// Returns 0 on success, 3 if the shape is outside what this kernel handles (caller falls back to the fp32 kernel).
int atmvfi_window_attention_tc_launch(const float* qkv, int qkv_pitch, float* out, int out_pitch, int C, int heads,
                                      const atmvfi_window_geom* g, int cross, const float* rc, float* motion_raw, int wy0, int nwy, int layout,
                                      cudaStream_t st) {
  AttnParams p;
  p.qkv = qkv; p.qkv_pitch = qkv_pitch; p.out = out; p.out_pitch = out_pitch; p.C = C; p.heads = heads; p.hd = C / heads;
  p.g = *g; p.cross = cross; p.rc = rc; p.motion_raw = motion_raw;
  p.out_f16 = atmvfi_act_f16();
  p.prof = atmvfi_attn_prof_buffer();
  p.N = g->ws * g->ws;
  if (p.N > 256 || p.hd % 4 != 0 || p.hd > 96) return 3;
  p.nW = (g->Hp / g->ws) * (g->Wp / g->ws);
  p.total_windows = g->B2 * p.nW;
  p.wy0 = wy0; p.per_img = nwy * (g->Wp / g->ws); p.virt_windows = g->B2 * p.per_img;
  if (p.N <= 64) { p.wpi = kRows / p.N; p.mtiles = 1; } else { p.wpi = 1; p.mtiles = (p.N + kRows - 1) / kRows; }
  p.KP = rup(p.wpi * p.N, 16);
  p.HP = rup(p.hd, 16);
  p.chunksH = (rup(p.hd, 8) + 31) / 32;
  p.chunksK = (p.KP + 31) / 32;
  int need = p.KP + p.HP, cols = 32;
  while (cols < need) cols <<= 1;
  if (cols > 512) return 3;
  p.tmem_cols = cols;
  p.round = atmvfi_output_rounding();
  p.scale = (float)(1.0 / sqrt((double)p.hd));
  const int bytesQ = p.chunksH * kRows * 128, bytesK = p.chunksH * p.KP * 128, bytesP = p.chunksK * kRows * 128;
  const int regionQK = bytesQ + bytesK > bytesP ? bytesQ + bytesK : bytesP;
  p.offK = bytesQ;
  p.offV = rup(regionQK, 1024);
  p.offLab = p.offV + p.chunksK * p.HP * 128;
  p.offBar = rup(p.offLab + p.KP * 4, 16);
  p.layout = layout;
  p.qrows = p.wpi > 1 ? p.N : kRows;
  if (layout == 1) {
    // V^T chunks of a window must start on a 32-key boundary of the tile; operands must satisfy the TMA alignment rules
    if ((p.wpi > 1 && p.N % 32 != 0) || ((uintptr_t)qkv & 15) || p.hd % 4) return 3;
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
      void* fp = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        enc = reinterpret_cast<EncodeTiledFn>(fp);
    }
    if (!enc) return 3;
    const cuuint64_t R = (cuuint64_t)p.total_windows * p.N;
    cuuint64_t gdim[3] = {(cuuint64_t)p.hd, R, (cuuint64_t)(2 * heads)};
    cuuint64_t gstr[2] = {(cuuint64_t)p.hd * 4, R * p.hd * 4};
    cuuint32_t estr[3] = {1, 1, 1};
    cuuint32_t boxq[3] = {32, (cuuint32_t)p.qrows, 1}, boxk[3] = {32, (cuuint32_t)p.N, 1};
    CUresult r1 = enc(&p.mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(qkv), gdim, gstr, boxq, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&p.mapK, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(qkv), gdim, gstr, boxk, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t vdim[2] = {R, (cuuint64_t)C};
    cuuint64_t vstr[1] = {R * 4};
    cuuint32_t boxv[2] = {32, (cuuint32_t)p.HP};
    CUresult r3 = enc(&p.mapV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(qkv) + 2 * (size_t)C * R, vdim, vstr, boxv, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS || r3 != CUDA_SUCCESS) return 3;
  }
  {
    // output staging + TMA store: needs 16-byte aligned head slices and whole 16-row boxes for windows larger than the tile rows
    static int o_tma_on = -1;
    if (o_tma_on < 0) { const char* ev = getenv("ATMVFI_ATTN_TMASTORE"); o_tma_on = ev ? atoi(ev) : 1; }
    const int es = p.out_f16 ? 2 : 4;
    p.o_tma = 0;
    p.o_pad = p.out_f16 ? 8 : (((p.hd + 4) % 8 == 4) ? 4 : 8);            // row pitch = odd number of 16-byte units
    p.o_rows = p.N <= 64 ? p.N : 16;
    const int pitch_b = (p.hd + p.o_pad) * es;
    if (o_tma_on && ((uintptr_t)out & 15) == 0 && (p.hd * es) % 16 == 0 && (out_pitch * es) % 16 == 0 && (p.N <= 64 || p.N % 16 == 0) &&
        pitch_b % 16 == 0 && (pitch_b / 16) % 2 == 1 && p.hd + p.o_pad <= 256 && (p.o_rows * pitch_b) % 128 == 0 && kRows * pitch_b <= (bytesQ + bytesK > bytesP ? bytesQ + bytesK : bytesP)) {
      typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
      static EncodeTiledFn enc_o = nullptr;
      if (!enc_o) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
          enc_o = reinterpret_cast<EncodeTiledFn>(fp);
      }
      if (enc_o) {
        const cuuint64_t R = (cuuint64_t)p.total_windows * p.N;
        cuuint64_t odim[3] = {(cuuint64_t)p.hd, (cuuint64_t)heads, R};
        cuuint64_t ostr[2] = {(cuuint64_t)p.hd * es, (cuuint64_t)out_pitch * es};
        cuuint32_t obox[3] = {(cuuint32_t)(p.hd + p.o_pad), 1, (cuuint32_t)p.o_rows};
        cuuint32_t oestr[3] = {1, 1, 1};
        if (enc_o(&p.mapO, p.out_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, odim, ostr, obox, oestr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
          p.o_tma = 1;
      }
    }
  }
  const int smem = p.offBar + 64 + 4 * kMaxParts * kRows * 4 + (kRows + 256) * 4 + 2 * 256 * 4;      // + row / key source tables, key coordinates
  if (smem > 227 * 1024) return 3;
  typedef void (*KernFn)(AttnParams);
  // [threads per row: 2, 4][head dim]
  static const KernFn kerns[2][5] = {
      {window_attention_tc_kernel<0, 2>, window_attention_tc_kernel<48, 2>, window_attention_tc_kernel<84, 2>, window_attention_tc_kernel<28, 2>,
       window_attention_tc_kernel<44, 2>},
      {window_attention_tc_kernel<0, 4>, window_attention_tc_kernel<48, 4>, window_attention_tc_kernel<84, 4>, window_attention_tc_kernel<28, 4>,
       window_attention_tc_kernel<44, 4>}};
  // four threads per row for windows larger than the 64 keys of an 8x8 window: their single resident CTA per SM (165 KB of operands)
  // needs the extra warps to hide its latencies (measured, Base 1080p: 12x12 global launch 214 -> 185 us; the 8x8 local launch, two
  // CTAs per SM, is 7 % slower with four); ATMVFI_ATTN_PARTS overrides
  static int parts_env = -1;
  if (parts_env < 0) { const char* ev = getenv("ATMVFI_ATTN_PARTS"); parts_env = ev ? atoi(ev) : 0; }
  const int parts = parts_env == 2 || parts_env == 4 ? parts_env : (p.N > 64 ? 4 : 2);
  const int pi = parts == 4 ? 1 : 0;
  const int ki = p.hd == 48 ? 1 : (p.hd == 84 ? 2 : (p.hd == 28 ? 3 : (p.hd == 44 ? 4 : 0)));
  static int configured_of_device[ATMVFI_MAX_DEVICES] = {0};      // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= ATMVFI_MAX_DEVICES) return 3;
  int& configured = configured_of_device[dev];
  if (configured < smem) {
    const int want = smem > 100 * 1024 ? 227 * 1024 : 100 * 1024;
    for (int i = 0; i < 10; ++i) {
      cudaError_t e = cudaFuncSetAttribute(kerns[i / 5][i % 5], cudaFuncAttributeMaxDynamicSharedMemorySize, want);
      if (e != cudaSuccess) {
        atmvfi_set_error("window_attention(tf32): cannot reserve shared memory: %s", cudaGetErrorString(e));
        return 1;
      }
    }
    configured = want;
    // (measured: forcing cudaSharedmemCarveoutMaxShared makes this kernel 15 % SLOWER - two CTAs already fit and the
    // gathers of Q/K/V profit from the L1 that the default split leaves)
  }
  const int64_t wgroups = (p.virt_windows + p.wpi - 1) / p.wpi;
  const int64_t items = wgroups * heads * p.mtiles;
  if (items <= 0) return 0;
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)items);
    cfg.blockDim = dim3(kRows * parts);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = atmvfi_pdl_enabled() ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kerns[pi][ki], p);
    if (le != cudaSuccess) {
      atmvfi_set_error("window_attention(tf32): launch failed: %s", cudaGetErrorString(le));
      return 1;
    }
  }
  return 0;
}
