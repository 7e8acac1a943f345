// Fused tail of the transformer Mlp (attention.py:74-85, 118-123; block residual attention.py:333):
//
//     out = x + fc2( GELU( DWConv3x3(h) + b_dw ) ) + b_fc2          h = fc1(norm2(x)), the 4C-wide hidden map
//
// The hidden activation GELU(DWConv(h)) is never written to HBM: it is produced tile by tile in shared memory, directly in the
// K-major SWIZZLE_128B layout of the A operand of a tcgen05 GEMM, by CUDA-core "converter" warps that sit between the TMA
// producer and the MMA issuer.  Per 128-pixel tile (16 x 8 pixels) and per K chunk (32 fp32 / 64 fp16 hidden channels):
//
//   warp 3       TMA: hidden box {chunk, 18, 10} with a one-pixel halo (zero-filled outside the image = the conv's padding, SWIZZLE_128B)
//                plus the ten rows {chunk, 10} of the depth-wise taps and bias -> raw ring (3 slots); L2 prefetch of the boxes five
//                steps ahead, so that the short ring only has to cover the L2 latency
//   warp 0       TMA: this CTA's half of the fc2 weight tile -> B ring (2 slots)
//   warps 4-19   workers in two groups of eight that take alternate steps: 3x3 depth-wise taps (packed fp32 FFMA2, in the order of
//                the stand-alone kernel), bias, GELU, rounding to the operand type, 16-byte stores into the swizzled A slot (3 slots),
//                fence.proxy.async, arrive.  The first twelve also run the epilogue of the previous tile right after their group's
//                first conversion of the next one: residual box by TMA into the warp's staging tile, accumulators from TMEM (+ bias
//                + residual) written back in place and handed to a TMA store
//   warp 1       (leader CTA) tcgen05.mma cta_group::2, M = 256 (both CTAs' pixel tiles), N = the WHOLE fc2 width (<= 384, issued
//                as 256 + rest) - so the hidden map is converted exactly once per pixel; each CTA holds half of the weight tile
//
// One accumulator stage (N columns of the 512): the epilogue of a tile is exposed.  The conversion (20 instructions per hidden
// element on the CUDA cores, 2 MUFU) bounds the kernel, not the tensor pipe or HBM.  History on B200, Base 1080p local grid
// (2 x 136 x 240 tokens, 384 <- 1536), tf32: the two stand-alone launches 177 + 225 us; first fused version 435 us (cluster-scope
// release / acquire on every step: ~1000 cycles each); 316 us without them; 224 us with two converter groups, three A slots and the
// raw slot handed back before the GELU half of a step (ncu: no pipe above 50 %, the converters were latency-bound at 0.45 IPC).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int kTW = 16, kTH = 8;                         // pixel tile = 128 GEMM rows; row r = y * 16 + x
constexpr int kHaloW = kTW + 2, kHaloH = kTH + 2;
constexpr int kRawBoxBytes = kHaloW * kHaloH * 128;      // 23040
constexpr int kRawSlotBytes = kRawBoxBytes + 10 * 256;   // + taps / bias rows (fp32: 128 B per row for 32 channels, 256 B for 64)
constexpr int kRawSlots = 3, kASlots = 3, kBSlots = 2;
constexpr int kABytes = 128 * 128;
constexpr int kMaxBSlotBytes = 192 * 128;                // half of a 384-row weight tile
constexpr int kWorkerWarps = 16, kGroupWarps = 8;        // two groups of converter warps
constexpr int kEpiWarps = 12;                            // the first twelve workers also run the epilogue (one 4 KB staging tile each)
constexpr int kEpiBufBytes = 4096;
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kASlots * kABytes;                       // 49152
constexpr int kOffEpi = kOffB + kBSlots * kMaxBSlotBytes;              // 98304
constexpr int kOffRaw = kOffEpi + kEpiWarps * kEpiBufBytes;            // 147456
constexpr int kOffBar = kOffRaw + kRawSlots * kRawSlotBytes;           // 224256
constexpr int kSmemBytes = kOffBar + 512;                              // 224768 of the 232448 available
constexpr int kThreads = 128 + 32 * kWorkerWarps;                      // 640
constexpr int kPrefetchAhead = 5;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kRawSlotBytes % 1024 == 0 && kOffRaw % 1024 == 0, "the swizzled hidden boxes start on 1024-byte boundaries");

struct MtParams {
  CUtensorMap mapRaw, mapW10, mapB1, mapB2, mapRes, mapOut;
  const float* bias;                   // fc2 bias, padded with zeros to n_tiles * block_n floats
  int B, H, W, C;
  int nk;                              // K steps = hidden channels / chunk
  int tiles_x, tiles_y, m_tiles, n_tiles, total_ctiles;
  int block_n, n1, n2;                 // columns per CTA tile; MMA split n1 (<= 256) + n2
  int row0;                            // first output row of every image (row window, include/atmvfi.h "ROW WINDOWS")
  int round;                           // tf32: round the produced operand / the output to TF32
  unsigned long long* prof;            // debug (ATMVFI_MT_PROF = 1 + CTA index): cycle counters of one CTA, see atmvfi_mlp_tail_prof_read
  int prof_cta;
};

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_u32(b)), "r"(n)); }
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_u32(b)) : "memory"); }
// arrive on the same barrier of CTA 0 of the pair.  (Plain release: the data the leader's MMA reads from this CTA's shared memory was
// published to the async proxy by fence.proxy.async before; a cluster-scope release here costs ~1000 cycles per step.)
__device__ __forceinline__ void bar_arrive_leader(uint64_t* b) {
  asm volatile(
      "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, 0;\nmbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(sm_u32(b))
      : "memory");
}
// Waits carry a suspend-time hint: the hardware parks the thread until the phase completes (or the hint, in ns, expires) instead
// of spinning.  ncu showed the idle roles (eight epilogue warps, the MMA and producer warps) executing 28 M try_wait / branch
// instructions per launch without it - issue slots taken from the converter warps that share their schedulers.
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(sm_u32(b)), "r"(parity), "r"(20000u)
                 : "memory");
}
__device__ __forceinline__ void bar_wait_cluster(uint64_t* b, uint32_t parity) { bar_wait(b, parity); }      // (arrivals from the peer CTA)
__device__ __forceinline__ bool elect1() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;          // shared::cluster address of the same offset in CTA 0 of the pair

__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(sm_u32(dst)),
               "l"(m), "r"(sm_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma2(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sm_u32(dst)), "l"(m),
               "r"(sm_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// weight half of this CTA, completion counted on the LEADER's barrier
__device__ __forceinline__ void tma2_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   sm_u32(dst)),
               "l"(m), "r"(sm_u32(bar) & kLeaderMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch4(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(sm_u32(src)), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void mma_pair_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_pair_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit_pair(uint64_t* bar) {      // arrives on `bar` in both CTAs once the MMAs issued so far retire
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(sm_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint64_t sdesc128(uint32_t saddr) {     // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_pair(int n, bool f16) {  // M = 256 across the pair, fp32 accumulate, A and B K-major
  uint32_t d = 1u << 4;
  if (!f16) d |= (2u << 7) | (2u << 10);
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(256 >> 4) << 24;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- packed fp32 arithmetic (sm_100 FFMA2 / FMUL2: two IEEE fp32 operations per instruction; the converters are issue-bound) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
struct F4 { f32x2 lo, hi; };                 // four channels as two packed pairs
__device__ __forceinline__ F4 f4_of(const float4& v) { F4 r; r.lo = pk2(v.x, v.y); r.hi = pk2(v.z, v.w); return r; }
__device__ __forceinline__ void fma4p(const F4& a, const F4& w, F4& c) { c.lo = fma2(a.lo, w.lo, c.lo); c.hi = fma2(a.hi, w.hi, c.hi); }
// gelu_fast (common.cuh) on a packed pair: the same fp32 operations in the same order, so the bits are those of the scalar version
__device__ __forceinline__ f32x2 gelu_fast2(f32x2 v) {
  const f32x2 x2 = mul2(v, pk2(0.70710678118654752440f, 0.70710678118654752440f));
  const f32x2 y2 = mul2(v, pk2(0.84932180028801904272f, 0.84932180028801904272f));
  const f32x2 yn2 = mul2(v, pk2(-0.84932180028801904272f, -0.84932180028801904272f));
  float x0, x1, z0, z1, t0, t1, e0, e1;
  upk2(x2, x0, x1);
  upk2(mul2(yn2, y2), z0, z1);                           // -(y * y)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(0.3275911f, fabsf(x0), 1.f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(0.3275911f, fabsf(x1), 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  const f32x2 t = pk2(t0, t1);
  // -(p t) with negated coefficients: fma(-p, e, 1) == fma(p', e, 1), p' = -p exactly
  f32x2 p = fma2(pk2(-1.061405429f, -1.061405429f), t, pk2(1.453152027f, 1.453152027f));
  p = fma2(p, t, pk2(-1.421413741f, -1.421413741f));
  p = fma2(p, t, pk2(0.284496736f, 0.284496736f));
  p = fma2(p, t, pk2(-0.254829592f, -0.254829592f));
  p = mul2(p, t);
  float r0, r1;
  upk2(fma2(p, pk2(e0, e1), pk2(1.f, 1.f)), r0, r1);
  const f32x2 r = pk2(copysignf(r0, x0), copysignf(r1, x1));
  const f32x2 h = mul2(v, pk2(0.5f, 0.5f));
  return fma2(h, r, h);
}

__device__ __forceinline__ void tile_of(const MtParams& p, int ct, int rank, int& n_tile, int& b, int& oy0, int& ox0) {
  n_tile = ct % p.n_tiles;
  int mt = (ct / p.n_tiles) * 2 + rank;
  const int tx = mt % p.tiles_x;
  mt /= p.tiles_x;
  const int ty = mt % p.tiles_y;
  b = mt / p.tiles_y;                        // >= B for the phantom tile of an odd tile count: TMA zero-fills / clips everything
  oy0 = p.row0 + ty * kTH;
  ox0 = tx * kTW;
}

// One K step of the conversion for one warp: DWConv 3x3 + bias + GELU of its 4-channel group(s) over the 16 x 8 pixel tile, from the
// swizzled halo box `raw` (followed by the ten tap / bias rows) into the K-major SWIZZLE_128B A tile `at`.  The raw slot is handed
// back (one arrival per warp on `raw_empty`) as soon as its values are in registers, before the GELU half of the step.
template <typename T>
__device__ __forceinline__ void convert_step(const uint8_t* raw, uint8_t* at, int cq, int x, int yh, int lane, bool rnd, uint64_t* raw_empty) {
  constexpr bool kF16 = Act<T>::kHalf;
  constexpr int kCh = kF16 ? 64 : 32;
  constexpr int kSub = kF16 ? 2 : 1;
  const float* w10 = reinterpret_cast<const float*>(raw + kRawBoxBytes);
#pragma unroll
  for (int sub = 0; sub < kSub; ++sub) {
    const int cg = cq + 8 * sub;                                    // 4-channel group inside the chunk
    const int unit = kF16 ? (cg >> 1) : cg, inner = kF16 ? ((cg & 1) << 3) : 0;   // 16-byte unit of a 128-byte pixel row, offset inside it
    F4 k[9], bz;
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = f4_of(*reinterpret_cast<const float4*>(w10 + t * kCh + 4 * cg));
    bz = f4_of(*reinterpret_cast<const float4*>(w10 + 9 * kCh + 4 * cg));
    F4 acc[4] = {bz, bz, bz, bz};
    auto px = [&](int pi) -> F4 { return f4_of(Act<T>::lds4(raw + pi * 128 + ((unit ^ (pi & 7)) << 4) + inner)); };
    const int pi0 = (4 * yh) * kHaloW + x;
#pragma unroll
    for (int i = 0; i < 6; ++i) {                                   // input rows 4 yh + i of the halo box
      const F4 l = px(pi0 + i * kHaloW), m = px(pi0 + i * kHaloW + 1), r = px(pi0 + i * kHaloW + 2);
      // taps in row-major order per output row (bias, top row, middle row, bottom row): the order of the stand-alone kernel
      if (i >= 2) { fma4p(l, k[6], acc[i - 2]); fma4p(m, k[7], acc[i - 2]); fma4p(r, k[8], acc[i - 2]); }
      if (i >= 1 && i <= 4) { fma4p(l, k[3], acc[i - 1]); fma4p(m, k[4], acc[i - 1]); fma4p(r, k[5], acc[i - 1]); }
      if (i <= 3) { fma4p(l, k[0], acc[i]); fma4p(m, k[1], acc[i]); fma4p(r, k[2], acc[i]); }
    }
    if (sub == kSub - 1) {
      // every value of the raw slot this warp needs is in registers: hand the slot back before the GELU half of the step
      // (two groups hold two of the three slots otherwise, leaving a single load in flight)
      __syncwarp();
      if (lane == 0) bar_arrive(raw_empty);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int row = (4 * yh + o) * kTW + x;
      float4 v;
      upk2(gelu_fast2(acc[o].lo), v.x, v.y);
      upk2(gelu_fast2(acc[o].hi), v.z, v.w);
      uint8_t* dst = at + row * 128 + ((unit ^ (row & 7)) << 4) + inner;
      if (kF16) *reinterpret_cast<uint2*>(dst) = Act<__half>::pack(v);
      else *reinterpret_cast<float4*>(dst) = round_tf32_if(v, rnd);
    }
  }
}

// Epilogue arithmetic of one 32-column chunk for one warp: lane = output pixel; its residual row sits in the staging tile `ebuf`
// (swizzled like the tensor map: SWIZZLE_128B for fp32 rows, SWIZZLE_64B for fp16 rows) and is replaced in place by
// accumulator + bias + residual, ready for the TMA store.
template <typename T>
__device__ __forceinline__ void epi_chunk(uint8_t* ebuf, const uint32_t (&r)[32], const float* bias, int lane, bool rnd) {
  constexpr bool kF16 = Act<T>::kHalf;
  constexpr int rb = 32 * (int)sizeof(T);
  const uint32_t swz = (uint32_t)((lane * rb) >> 7) & (uint32_t)((rb >> 4) - 1);
  uint8_t* rowp = ebuf + lane * rb;
  if (kF16) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4* cell = reinterpret_cast<uint4*>(rowp + (((uint32_t)c ^ swz) << 4));
      const uint4 rv = *cell;
      const float4 bz0 = __ldg(reinterpret_cast<const float4*>(bias + 8 * c)), bz1 = __ldg(reinterpret_cast<const float4*>(bias + 8 * c + 4));
      const float4 ra = Act<__half>::unpack(make_uint2(rv.x, rv.y)), rb4 = Act<__half>::unpack(make_uint2(rv.z, rv.w));
      const float4 va = make_float4(__uint_as_float(r[8 * c]) + bz0.x + ra.x, __uint_as_float(r[8 * c + 1]) + bz0.y + ra.y,
                                    __uint_as_float(r[8 * c + 2]) + bz0.z + ra.z, __uint_as_float(r[8 * c + 3]) + bz0.w + ra.w);
      const float4 vb = make_float4(__uint_as_float(r[8 * c + 4]) + bz1.x + rb4.x, __uint_as_float(r[8 * c + 5]) + bz1.y + rb4.y,
                                    __uint_as_float(r[8 * c + 6]) + bz1.z + rb4.z, __uint_as_float(r[8 * c + 7]) + bz1.w + rb4.w);
      const uint2 pa = Act<__half>::pack(va), pb = Act<__half>::pack(vb);
      *cell = make_uint4(pa.x, pa.y, pb.x, pb.y);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4* cell = reinterpret_cast<float4*>(rowp + (((uint32_t)c ^ swz) << 4));
      const float4 rv = *cell;
      const float4 bz = __ldg(reinterpret_cast<const float4*>(bias + 4 * c));
      *cell = round_tf32_if(make_float4(__uint_as_float(r[4 * c]) + bz.x + rv.x, __uint_as_float(r[4 * c + 1]) + bz.y + rv.y,
                                        __uint_as_float(r[4 * c + 2]) + bz.z + rv.z, __uint_as_float(r[4 * c + 3]) + bz.w + rv.w),
                            rnd);
    }
  }
}

// debug timers: CTA 0 only, one thread per role
#define MT_T0() const long long t0__ = prof_on ? clock64() : 0
#define MT_ADD(slot) do { if (prof_on) atomicAdd(&p.prof[slot], (unsigned long long)(clock64() - t0__)); } while (0)

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) mlp_tail_kernel(const __grid_constant__ MtParams p) {
  constexpr bool kF16 = Act<T>::kHalf;
  constexpr int kCh = kF16 ? 64 : 32;                      // hidden channels per K step (one 128-byte swizzle row)
  constexpr int kSub = kF16 ? 2 : 1;                       // 4-channel groups a converter thread handles per step
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (sm_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* rawFull = bars;                    // [3]
  uint64_t* rawEmpty = rawFull + kRawSlots;    // [3]
  uint64_t* aFull = rawEmpty + kRawSlots;      // [2]  leader's: both CTAs' converters
  uint64_t* aEmpty = aFull + kASlots;          // [2]
  uint64_t* bFull = aEmpty + kASlots;          // [2]  leader's: both halves of the weight tile
  uint64_t* bEmpty = bFull + kBSlots;          // [2]
  uint64_t* tFull = bEmpty + kBSlots;          // [1]
  uint64_t* tEmpty = tFull + 1;                // [1]  leader's: both CTAs' epilogues
  uint64_t* resFull = tEmpty + 1;              // [16 warps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resFull + kWorkerWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cta_rank();
  const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  unsigned long long gt0 = 0;
  if (p.prof && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
  const bool prof_on = p.prof != nullptr && (int)blockIdx.x == p.prof_cta && lane == 0 && warp <= 4;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRawSlots; ++s) { bar_init(&rawFull[s], 1); bar_init(&rawEmpty[s], kGroupWarps); }
    for (int s = 0; s < kASlots; ++s) { bar_init(&aFull[s], 2 * kGroupWarps); bar_init(&aEmpty[s], 1); }
    for (int s = 0; s < kBSlots; ++s) { bar_init(&bFull[s], 1); bar_init(&bEmpty[s], 1); }
    bar_init(tFull, 1);
    bar_init(tEmpty, 2 * kEpiWarps);
    for (int s = 0; s < kWorkerWarps; ++s) bar_init(&resFull[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sm_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();                // programmatic dependent launch (common.cuh): nothing above touched global memory
  pdl_wait();

  const int b_rows1 = p.n1 >> 1, b_rows2 = p.n2 >> 1;       // weight rows of the two MMAs held by one CTA
  const uint32_t b_tx = (uint32_t)p.block_n * 128u;         // both halves

  if (warp == 3) {
    // ======================================= TMA producer: hidden boxes + taps ===================
    int rs = 0;
    uint32_t rph = 0;
    // L2 prefetch cursor: runs kPrefetchAhead steps ahead of the loads (across tiles)
    int pf_ct = cluster_id, pf_g = 0, pf_lead = 0;
    auto prefetch_one = [&]() {
      if (pf_ct >= p.total_ctiles) return;
      int n_tile, b, oy0, ox0;
      tile_of(p, pf_ct, rank, n_tile, b, oy0, ox0);
      if (elect1() && b < p.B) tma_prefetch4(&p.mapRaw, pf_g * kCh, ox0 - 1, oy0 - 1, b);
      __syncwarp();
      if (++pf_g == p.nk) { pf_g = 0; pf_ct += num_clusters; }
    };
    for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters) {
      int n_tile, b, oy0, ox0;
      tile_of(p, ct, rank, n_tile, b, oy0, ox0);
      for (int g = 0; g < p.nk; ++g) {
        while (pf_lead < kPrefetchAhead + 1) { prefetch_one(); ++pf_lead; }
        --pf_lead;
        { MT_T0(); bar_wait(&rawEmpty[rs], rph ^ 1); MT_ADD(0); }
        if (elect1()) {
          uint8_t* dst = smem + kOffRaw + rs * kRawSlotBytes;
          bar_expect(&rawFull[rs], (uint32_t)(kRawBoxBytes + 10 * kCh * 4));
          tma4(dst, &p.mapRaw, &rawFull[rs], g * kCh, ox0 - 1, oy0 - 1, b);
          tma2(dst + kRawBoxBytes, &p.mapW10, &rawFull[rs], g * kCh, 0);
        }
        __syncwarp();
        if (++rs == kRawSlots) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ======================================= TMA producer: this CTA's half of the fc2 weight tiles
    int bs = 0;
    uint32_t bph = 0;
    for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters) {
      const int n0 = (ct % p.n_tiles) * p.block_n;
      for (int g = 0; g < p.nk; ++g) {
        { MT_T0(); bar_wait(&bEmpty[bs], bph ^ 1); MT_ADD(1); }
        if (elect1()) {
          uint8_t* dst = smem + kOffB + bs * kMaxBSlotBytes;
          if (rank == 0) bar_expect(&bFull[bs], b_tx);
          tma2_pair(dst, &p.mapB1, &bFull[bs], g * kCh, n0 + rank * b_rows1);
          if (p.n2) tma2_pair(dst + b_rows1 * 128, &p.mapB2, &bFull[bs], g * kCh, n0 + p.n1 + rank * b_rows2);
        }
        __syncwarp();
        if (++bs == kBSlots) { bs = 0; bph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer (leader CTA) ============================
    if (rank == 0) {
      const uint32_t id1 = idesc_pair(p.n1, kF16), id2 = idesc_pair(p.n2 ? p.n2 : 16, kF16);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, tcount = 0;
      for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters, ++tcount) {
        { MT_T0(); if (tcount) bar_wait_cluster(tEmpty, (tcount - 1) & 1); MT_ADD(2); }   // both epilogues have drained the accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int g = 0; g < p.nk; ++g) {
          { MT_T0(); bar_wait_cluster(&aFull[as], aph); MT_ADD(3); }
          { MT_T0(); bar_wait(&bFull[bs], bph); MT_ADD(4); }
          MT_T0();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ad = sdesc128(sm_u32(smem + kOffA + as * kABytes));
          const uint64_t bd1 = sdesc128(sm_u32(smem + kOffB + bs * kMaxBSlotBytes));
          const uint64_t bd2 = bd1 + (uint64_t)((b_rows1 * 128) >> 4);
          if (elect1()) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t acc = (g == 0 && j == 0) ? 0u : 1u;
              if (kF16) {
                mma_pair_f16(tmem_base, ad + 2 * j, bd1 + 2 * j, id1, acc);
                if (p.n2) mma_pair_f16(tmem_base + p.n1, ad + 2 * j, bd2 + 2 * j, id2, acc);
              } else {
                mma_pair_tf32(tmem_base, ad + 2 * j, bd1 + 2 * j, id1, acc);
                if (p.n2) mma_pair_tf32(tmem_base + p.n1, ad + 2 * j, bd2 + 2 * j, id2, acc);
              }
            }
            commit_pair(&aEmpty[as]);
            commit_pair(&bEmpty[bs]);
            if (g == p.nk - 1) commit_pair(tFull);
          }
          __syncwarp();
          MT_ADD(5);
          if (++as == kASlots) { as = 0; aph ^= 1; }
          if (++bs == kBSlots) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================= workers: converters + epilogue ======================
    // Sixteen warps in two groups of eight; group (s & 1) converts step s of the CTA's running step count into A slot (s % 3), so
    // that four converter warps per scheduler hide each other's latencies (two were latency-bound at 0.45 instructions per cycle).
    // In a group a warp owns one 4-channel set of the chunk (fp16: two), its lanes are 16 columns x 2 row halves: the taps are one
    // broadcast read for the warp, and with the SWIZZLE_128B layout of the box (16-byte unit ^ pixel index) the 16 pixels a half-
    // warp touches sit in different bank groups.  The epilogue of a tile is shared by all sixteen warps (TMEM lane quarter warp & 3,
    // every fourth 32-column chunk) and runs right after the group's first conversion of the NEXT tile, whose MMAs are waiting for
    // the accumulator anyway.
    const int wk = warp - 4;
    const int grp = wk >> 3, cq = wk & 7, x = lane & 15, yh = lane >> 4;
    const int q = warp & 3, eidx = wk >> 2;                 // epilogue (wk < 12): TMEM lane quarter (tile rows 2q, 2q + 1), chunks eidx, eidx + 3, ...
    const bool epi_warp = wk < kEpiWarps;
    uint8_t* const ebuf = smem + kOffEpi + wk * kEpiBufBytes;
    uint64_t* const rbar = resFull + wk;
    constexpr int es = (int)sizeof(T);
    constexpr int rb = 32 * es;                             // bytes per staged row: 128 (fp32) / 64 (fp16)
    const uint32_t swz = (uint32_t)((lane * rb) >> 7) & (uint32_t)((rb >> 4) - 1);
    uint32_t rpar = 0;

    // epilogue state of the tile whose accumulators are pending
    int e_n0 = 0, e_b = 0, e_cy = 0, e_ox0 = 0, e_mine = 0;
    auto load_res = [&](int j) {                            // j-th chunk of this warp -> the staging tile (lane 0)
      bar_expect(rbar, (uint32_t)(32 * rb));
      tma4(ebuf, &p.mapRes, rbar, e_n0 + (eidx + 3 * j) * 32, e_ox0, e_cy, e_b);
    };
    auto epilogue_begin = [&](int ct) {                     // tile ct just finished its conversions: set up, prefetch the first residual box
      int n_tile, oy0;
      tile_of(p, ct, rank, n_tile, e_b, oy0, e_ox0);
      e_n0 = n_tile * p.block_n;
      int nch = (min(p.block_n, p.C - e_n0) + 31) >> 5;     // 32-column chunks of this tile that hold real channels
      e_mine = (epi_warp && nch > eidx) ? (nch - eidx + 2) / 3 : 0;
      e_cy = oy0 + 2 * q;
      if (lane == 0 && e_mine > 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's stores have left the staging tile
        load_res(0);
        for (int j = 1; j < e_mine; ++j)                                   // later boxes: into L2 now, into the tile when it is free
          if (e_b < p.B) tma_prefetch4(&p.mapRes, e_n0 + (eidx + 3 * j) * 32, e_ox0, e_cy, e_b);
      }
      __syncwarp();
    };
    auto epilogue_run = [&](uint32_t tcount) {
      if (!epi_warp) return;
      { MT_T0(); bar_wait(tFull, tcount & 1); MT_ADD(9); }
      MT_T0();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int j = 0; j < e_mine; ++j) {
        const int u = eidx + 3 * j;
        const int co0 = e_n0 + u * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * 32), r);
        bar_wait(rbar, rpar);
        rpar ^= 1;
        epi_chunk<T>(ebuf, r, p.bias + co0, lane, p.round != 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store4(&p.mapOut, ebuf, co0, e_ox0, e_cy, e_b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (j + 1 < e_mine) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the box has left the staging tile
            load_res(j + 1);
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) bar_arrive(tEmpty); else bar_arrive_leader(tEmpty);
      }
      MT_ADD(10);
    };

    int rs = 0, as = 0;
    uint32_t rph = 0, aph = 0, gs = 0, tcount = 0;
    int prev_ct = -1;
    for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters, ++tcount) {
      bool first = true;                                    // this group's first conversion of the tile is still to come
      for (int g = 0; g < p.nk; ++g, ++gs) {
        if ((int)(gs & 1) == grp) {
          { MT_T0(); bar_wait(&rawFull[rs], rph); MT_ADD(6); }
          { MT_T0(); bar_wait(&aEmpty[as], aph ^ 1); MT_ADD(7); }
          MT_T0();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          convert_step<T>(smem + kOffRaw + rs * kRawSlotBytes, smem + kOffA + as * kABytes, cq, x, yh, lane, p.round != 0, &rawEmpty[rs]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05.mma
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) bar_arrive(&aFull[as]); else bar_arrive_leader(&aFull[as]);
          }
          MT_ADD(8);
          if (first) {
            first = false;
            if (prev_ct >= 0) epilogue_run(tcount - 1);     // the previous tile's accumulators: the MMAs of this tile wait for them
          }
        }
        if (++rs == kRawSlots) { rs = 0; rph ^= 1; }
        if (++as == kASlots) { as = 0; aph ^= 1; }
      }
      epilogue_begin(ct);
      prev_ct = ct;
    }
    if (prev_ct >= 0) epilogue_run(tcount - 1);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // stores complete before the CTA may exit
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                          // no CTA exits while the peer may still signal it or read its shared memory
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
  if (p.prof && threadIdx.x == 0 && blockIdx.x < 160) {
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    p.prof[16 + 2 * blockIdx.x] = gt0;
    p.prof[17 + 2 * blockIdx.x] = gt1;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn mt_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

unsigned long long* g_mt_prof = nullptr;

}  // namespace

// Debug aid: with ATMVFI_MT_PROF=1 CTA 0 of atmvfi_mlp_tail accumulates clock cycles per role: [0] producer waits for a free raw slot,
// [1] ... a free weight slot, [2] MMA warp waits for the drained accumulator, [3] ... for the converted A tile, [4] ... for the weight
// tile, [5] MMA issue, [6] converters wait for the raw box, [7] ... for a free A slot, [8] conversion, [9] epilogue waits for the
// accumulator, [10] epilogue.  Reads and clears the counters; non-zero when profiling is off.
extern "C" int atmvfi_mlp_tail_prof_read(unsigned long long* out16) {
  if (!g_mt_prof) return 1;
  cudaMemcpy(out16, g_mt_prof, 336 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);      // [16 + 2 i], [17 + 2 i]: globaltimer at the start / end of CTA i (last launch)
  cudaMemset(g_mt_prof, 0, 16 * sizeof(unsigned long long));
  return 0;
}

extern "C" int atmvfi_mlp_tail(const void* hidden, int hid_pitch, int B, int H, int W, int Ch, const float* w10, const void* w_fc2, int w_rows,
                               const float* bias_fc2, const void* residual, int res_pitch, void* out, int out_pitch, int C, int precision,
                               int y0, int y1, void* stream) {
  ATMVFI_REQUIRE(precision == ATMVFI_TF32 || precision == ATMVFI_F16, "mlp_tail: precision must be ATMVFI_TF32 or ATMVFI_F16");
  const bool f16 = precision == ATMVFI_F16;
  const int es = f16 ? 2 : 4, chunk = f16 ? 64 : 32;
  ATMVFI_REQUIRE(hidden && w10 && w_fc2 && bias_fc2 && residual && out, "mlp_tail: null argument");
  ATMVFI_REQUIRE(B > 0 && H > 0 && W > 0 && Ch > 0 && Ch % chunk == 0, "mlp_tail: hidden width %d must be a multiple of %d", Ch, chunk);
  ATMVFI_REQUIRE(C > 0 && C % 32 == 0, "mlp_tail: output width %d must be a multiple of 32", C);
  ATMVFI_REQUIRE((((uintptr_t)hidden | (uintptr_t)w10 | (uintptr_t)w_fc2 | (uintptr_t)bias_fc2 | (uintptr_t)residual | (uintptr_t)out) & 15) == 0 &&
                     (hid_pitch * es) % 16 == 0 && (res_pitch * es) % 16 == 0 && (out_pitch * es) % 16 == 0,
                 "mlp_tail: operands must be 16-byte aligned with pitches of whole 16-byte units");
  int ny;
  ATMVFI_REQUIRE(row_window(H, y0, y1, &y0, &ny), "mlp_tail: bad row window [%d,%d) for H=%d", y0, y1, H);
  if (ny <= 0) return 0;
  EncodeTiledFn enc = mt_get_encode();
  ATMVFI_REQUIRE(enc != nullptr, "mlp_tail: cuTensorMapEncodeTiled not available");
  MtParams p;
  memset(&p, 0, sizeof(p));
  p.n_tiles = (C + 383) / 384;
  p.block_n = ((C + p.n_tiles - 1) / p.n_tiles + 31) / 32 * 32;
  p.n1 = p.block_n < 256 ? p.block_n : 256;
  p.n2 = p.block_n - p.n1;
  ATMVFI_REQUIRE(p.block_n <= 384 && p.n1 % 16 == 0 && p.n2 % 16 == 0, "mlp_tail: tile width %d unsupported", p.block_n);
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint32_t one4[4] = {1, 1, 1, 1};
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)hid_pitch * es, (cuuint64_t)hid_pitch * es * W, (cuuint64_t)hid_pitch * es * W * H};
    cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, 1};
    CUresult r = enc(&p.mapRaw, dt, 4, const_cast<void*>(hidden), gdim, gstr, box, one4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ATMVFI_REQUIRE(r == CUDA_SUCCESS, "mlp_tail: cuTensorMapEncodeTiled(hidden) failed with %d", (int)r);
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Ch, 10};
    cuuint64_t gstr[1] = {(cuuint64_t)Ch * 4};
    cuuint32_t box[2] = {(cuuint32_t)chunk, 10};
    CUresult r = enc(&p.mapW10, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w10), gdim, gstr, box, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ATMVFI_REQUIRE(r == CUDA_SUCCESS, "mlp_tail: cuTensorMapEncodeTiled(taps) failed with %d", (int)r);
  }
  for (int part = 0; part < 2; ++part) {
    const int rows = (part == 0 ? p.n1 : p.n2) / 2;
    if (!rows) continue;
    cuuint64_t gdim[2] = {(cuuint64_t)Ch, (cuuint64_t)w_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)Ch * es};
    cuuint32_t box[2] = {(cuuint32_t)chunk, (cuuint32_t)rows};
    CUresult r = enc(part == 0 ? &p.mapB1 : &p.mapB2, dt, 2, const_cast<void*>(w_fc2), gdim, gstr, box, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ATMVFI_REQUIRE(r == CUDA_SUCCESS, "mlp_tail: cuTensorMapEncodeTiled(weights %d) failed with %d", part, (int)r);
  }
  for (int which = 0; which < 2; ++which) {
    const void* ptr = which == 0 ? residual : out;
    const int pitch = which == 0 ? res_pitch : out_pitch;
    // y extent = end of the row window: the tiles of the last tile row hang over it and the TMA unit clips them (strides keep H)
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)(y0 + ny), (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * W, (cuuint64_t)pitch * es * W * H};
    cuuint32_t box[4] = {32, (cuuint32_t)kTW, 2, 1};
    CUresult r = enc(which == 0 ? &p.mapRes : &p.mapOut, dt, 4, const_cast<void*>(ptr), gdim, gstr, box, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, which == 0 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ATMVFI_REQUIRE(r == CUDA_SUCCESS, "mlp_tail: cuTensorMapEncodeTiled(%s) failed with %d", which == 0 ? "residual" : "out", (int)r);
  }
  p.bias = bias_fc2;
  p.B = B; p.H = H; p.W = W; p.C = C;
  p.nk = Ch / chunk;
  p.row0 = y0;
  p.tiles_x = cdiv(W, kTW); p.tiles_y = cdiv(ny, kTH);
  p.m_tiles = p.tiles_x * p.tiles_y * B;
  p.total_ctiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  p.round = (!f16 && atmvfi_output_rounding()) ? 1 : 0;
  {
    static int prof = -1;
    if (prof < 0) {
      const char* ev = getenv("ATMVFI_MT_PROF");
      prof = ev && atoi(ev) > 0 ? atoi(ev) : 0;
      if (prof) { cudaMalloc(&g_mt_prof, 336 * sizeof(unsigned long long)); cudaMemset(g_mt_prof, 0, 336 * sizeof(unsigned long long)); }
    }
    p.prof = prof ? g_mt_prof : nullptr;
    p.prof_cta = prof - 1;
  }

  typedef void (*KernFn)(MtParams);
  KernFn kern = f16 ? (KernFn)mlp_tail_kernel<__half> : (KernFn)mlp_tail_kernel<float>;
  static int sms_of_device[ATMVFI_MAX_DEVICES] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  ATMVFI_REQUIRE(dev >= 0 && dev < ATMVFI_MAX_DEVICES, "mlp_tail: device ordinal %d out of range", dev);
  if (!sms_of_device[dev]) {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(mlp_tail_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tail_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) {
      atmvfi_set_error("mlp_tail: cannot reserve %d B of shared memory: %s", kSmemBytes, cudaGetErrorString(e));
      return 1;
    }
    sms_of_device[dev] = n_sm;
  }
  int clusters = sms_of_device[dev] / 2;
  if (p.total_ctiles < clusters) clusters = p.total_ctiles;
  if (clusters <= 0) return 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = atmvfi_pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
  if (le != cudaSuccess) {
    atmvfi_set_error("mlp_tail: launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return 0;
}
