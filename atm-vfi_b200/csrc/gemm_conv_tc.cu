// TF32 implicit-GEMM convolution / linear layer on the 5th-generation tensor cores (sm_100a).
//
//   D[128 pixels x BLOCK_N channels] (fp32, TMEM) += A[128 x 32] (smem, K-major, SW128) * B[BLOCK_N x 32]^T
//
// * A is never materialised.  The NHWC source is a 4-D TMA tensor {C, W, H, B}; one box of 32 channels
//   x (TW x TH pixels) lands in shared memory as 128-byte rows with the 128-byte swizzle the UMMA
//   descriptor expects.  TMA zero-fills outside the image (= conv zero padding) and beyond the source's
//   channel count; stride-2/4 convolutions use the tensor map's element strides; channel concatenation of
//   up to 4 inputs is a loop over 4 tensor maps.
// * 3x3 / stride 1 layers ("halo" mode) fetch, per 32-channel chunk and per horizontal tap kx, ONE box
//   with a vertical halo {32, TW, TH+2}; the three vertical taps are UMMA descriptors offset by ky*TW rows
//   (a multiple of the 1024-byte swizzle atom), so the activation traffic from L2 is 3 boxes instead of 9.
// * B (weights) is packed on the host as [N_pad][K_tc] K-major fp32 (pre-rounded to TF32).  CTAs are
//   launched in clusters of 2 along M: each CTA fetches half of the B tile and TMA-multicasts it to both,
//   halving the weight traffic from L2 (the kernel is L2-bandwidth-bound otherwise).
// * Warp roles (CTA = 256 threads, persistent, one CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer
//   (one elected lane, tcgen05.mma cta_group::1 kind::tf32, M=128, N=BLOCK_N, K=8), warp 2 = TMEM allocator,
//   warps 4-7 = epilogue (tcgen05.ld -> smem transpose -> bias / residual / PReLU / row remap -> coalesced
//   global stores).
// * Separate smem rings for A (3 slots x 24 KB) and B (4 slots x 32 KB) with full/empty mbarriers, and
//   2 TMEM accumulator stages (512 columns) so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "gemm_epilogue.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kChunk = 32;                      // fp32 elements per K step = one 128-byte swizzle row
constexpr int kMaxASlots = 4;
constexpr int kMaxBSlots = 12;
constexpr int kMaxBlockN = 256;
constexpr int kDataBytes = 184 * 1024;          // A ring + B ring, carved per layer (see atmvfi_gemm_conv_tc)
constexpr int kDataBytes16 = 146 * 1024;        // ... when 16 epilogue warps need twice the staging area
constexpr int kMaxABoxBytes = 24 * 1024;        // up to 192 rows of 128 B (halo box; 320 rows = 40 KB in pair mode), 128 rows otherwise
// Epilogue warps: 8 (two per TMEM lane quarter, alternating column chunks) or, for layers whose K is so short that the
// accumulator drain + global stores bound the tile time (k2s2 transposed convs, 1x1 and 24/48-channel 3x3 layers), 16.
constexpr int kEpiPitch = 36;                                 // floats per staged row: 16-byte aligned, conflict-free for 128-bit access
// per epilogue warp: 32 x 36 floats staging + 32 x int2 row info, rounded up to 5 KB so that every warp's region starts on a 1024-byte
// boundary (the TMA-store epilogue stages 32 rows x <= 128 B there in the SWIZZLE_128B / 64B / 32B layouts of its tensor maps)
constexpr int kEpiWarpBytes = 5 * 1024;
static_assert(32 * kEpiPitch * 4 + 32 * 8 <= kEpiWarpBytes, "epilogue staging region");
// shared memory: [A / B rings][epilogue staging, one region per warp][barriers, 512 B]
constexpr int smem_bytes(int epi_warps) { return (epi_warps == 16 ? kDataBytes16 : kDataBytes) + epi_warps * kEpiWarpBytes + 512; }
constexpr int threads_for(int epi_warps) { return 128 + 32 * epi_warps; }

struct TcPlan {                                 // host-side, produced by atmvfi_gemm_conv_plan
  CUtensorMap mapA[ATMVFI_MAX_SRC];
  CUtensorMap mapB;
  int nsrc;
  int chunks[ATMVFI_MAX_SRC];                   // 32-channel chunks per source
  int ntaps, ksize, stride, dil, pad;
  int TW, TH, tiles_x, tiles_y, B;
  int block_n, n_tiles, cq_pad;
  int Hout, Wout;
  int halo;                                     // 1: 3x3 stride-1 layer, A boxes carry a vertical halo
  int pair;                                     // 1: two vertically stacked 128-pixel tiles per CTA step (N <= 128)
  int cluster;                                  // CTAs per cluster (B multicast), 1 or 2
  int row0, row1;                               // row window of the GEMM grid: tiles cover output rows [row0, row1)
  int x3;                                       // 1: 3xTF32 (fp32-tolerance) datapath, weights packed as hi | lo chunk pairs
  int f16, chunk;                               // fp16 operands; elements per 128-byte K row (32 fp32 / 64 fp16)
  // TMA-store epilogue: output tensor maps (32-channel boxes + the 16-channel tail of block_n), one pair per destination
  CUtensorMap mapOut, mapOutTail, mapOut2, mapOut2Tail;
  int st_ok, st_bx, st_by, st_shuffle;
  int grp;                                      // grouped main loop: 0 off, 1 halo (3 vertical taps per weight slot), 2 chunk pairs
  uint32_t magic;
};
constexpr uint32_t kPlanMagic = 0xA7B20007u;

struct TcParams {
  CUtensorMap mapA[ATMVFI_MAX_SRC];
  CUtensorMap mapB;
  int nsrc;
  int chunks[ATMVFI_MAX_SRC];
  int last_mmas[ATMVFI_MAX_SRC];               // K=8 MMAs needed by the last (partial) 32-channel chunk of each source
  int ntaps, ksize, stride, dil, pad;
  int TW, TH, tiles_x, tiles_y, B;
  int block_n, n_tiles, cq_pad;
  int halo, a_bytes, sum_chunks, m_tiles;
  int th_super;                                 // rows of output covered by one CTA tile (TH, or 2*TH in pair mode)
  int a_slots, a_slot_bytes, b_slots, b_slot_bytes;   // smem rings: A at offset 0, (3xTF32: the A-lo ring,) B right after
  int a_lo_off, b_off;                          // byte offsets of the A-lo ring (3xTF32 only) and of the B ring
  int bar_off;                                  // barriers: behind the data rings and the epilogue staging regions
  int epi_off;                                  // epilogue staging regions (1024-byte aligned), right behind the data rings
  // TMA-store epilogue (kEpi == 4): the stored tile of one warp and one 32-column chunk is a box {st_w channels, st_bx, st_by} of the output
  CUtensorMap mapOut, mapOutTail, mapOut2, mapOut2Tail;
  int st_bx, st_by;                             // pixels of a warp's 32 rows along x / y (st_bx * st_by == 32)
  int st_shuffle;                               // 1: ConvTranspose k2 s2 - the box walks the output with element stride 2 from (2x + dx, 2y + dy)
  // Grouped main loop (fewer, larger pipeline steps: the MMA warp pays ~0.35 us of barrier / commit latency per step, which
  // bounded every layer with few MMAs per step).  grp 1 (3x3 stride-1 halo layers): one weight slot holds the 3 vertical taps of a
  // (chunk, kx) group, fetched by ONE 3-D TMA box.  grp 2 (1x1 / strided layers): a step covers TWO K chunks - two activation
  // boxes in one slot, their two weight tiles fetched by one 3-D box.  0: one weight tile per step (3xTF32, full-halo mode).
  int grp;
  int dbg;                                      // ATMVFI_TC_DEBUG_EPI (timing experiments only): 1 no TMA store, 2 no staging either, 3 no accumulator read
  int total_ctiles;                             // cluster tiles: ceil(m_tiles / cluster) * n_tiles
  int row0, row1;                               // output rows [row0, row1) of every image (row window)
  EpiParams epi;
};

// ------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// non-blocking probe (try_wait may suspend the thread until a time-out when the phase is not complete)
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// bulk tensor store shared -> global (the epilogue's output path): clipped to the tensor's extent by the TMA unit
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // smem may be reused
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // stores are complete

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// cta_group::2 variants: the two CTAs of a cluster drive one M = 256 MMA.  TMA completions of BOTH CTAs are counted on
// the LEADER's mbarrier (shared::cluster address with the peer bit cleared), commits are multicast to both CTAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_2cta(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the same barrier in CTA rank 0 of the cluster
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// One lane of a converged warp; the warp keeps executing uniformly, which lets ptxas hold descriptors and
// coordinates in uniform registers (a whole role under `if (lane == 0)` compiles to ELECT/R2UR loops per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 (fp16 operands, fp32 accumulate): K = 16 per instruction, i.e. the same 32 bytes of a 128-byte swizzle row as kind::tf32
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool kF16, int kCS>
__device__ __forceinline__ void tc_mma_any(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (kF16) {
    if (kCS == 2) tc_mma_f16_2cta(tmem_d, adesc, bdesc, idesc, accumulate); else tc_mma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    if (kCS == 2) tc_mma_tf32_2cta(tmem_d, adesc, bdesc, idesc, accumulate); else tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle: 8-row groups of 1024 B (SBO = 1024), LBO unused, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (ignored for swizzled K-major), bits [16,30)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;              // stride byte offset (between 8-row groups), bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version, bits [46,48)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B, bits [61,64)
  return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc_tf32(int n, int m = kBlockM) {
  uint32_t d = 0;
  d |= 1u << 4;                 // D format = F32
  d |= 2u << 7;                 // A format = TF32
  d |= 2u << 10;                // B format = TF32
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// kind::f16 with fp16 A and B (format code 0), fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc_f16(int n, int m = kBlockM) {
  uint32_t d = 0;
  d |= 1u << 4;                 // D format = F32; A / B format fields stay 0 = F16
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// cluster tile -> (n_tile, image, tile origin) for this CTA; m tiles beyond the last one are phantoms whose
// coordinates fall outside the tensor (TMA zero-fills, the epilogue stores nothing).
__device__ __forceinline__ void tile_coords(const TcParams& p, int ctile, int cs, int rank, int& n_tile, int& b, int& oy0, int& ox0) {
  n_tile = ctile % p.n_tiles;
  int mt = (ctile / p.n_tiles) * cs + rank;
  int tx = mt % p.tiles_x;
  mt /= p.tiles_x;
  int ty = mt % p.tiles_y;
  b = mt / p.tiles_y;                      // may be >= B for a phantom tile
  oy0 = p.row0 + ty * p.th_super;
  ox0 = tx * p.TW;
}

// kEpi: 0 = generic epilogue, 1 = plain layers (see "fast epilogue" below), 2 = generic with the residual rows prefetched,
// 3 = the fast epilogue writing the head-major q | k | v^T layout (its own instantiation: carrying that code in variant 1 cost
// every plain layer ~5 %)
//
// kX3 - "3xTF32": fp32-tolerance results on the tensor cores.  Every product a*b is issued as three kind::tf32 MMAs into the same
// TMEM accumulator: a_hi*b_lo + a_lo*b_hi + a_hi*b_hi with x_hi = tf32(x), x_lo = tf32(x - x_hi).  Weights arrive split from the
// host (pack_tc_x3: each 32-channel chunk is a hi row block followed by a lo row block).  Activations stay plain fp32 in HBM: the
// tensor core itself truncates the raw box to a_hi, and warps 2-3 ("converters") write a_lo = rna_tf32(a - trunc(a)) for every
// landed box into a second shared-memory ring with the same swizzled layout before the MMA warp may touch the slot.
//
// kF16 - precision ATMVFI_F16: sources / weights are fp16 (64 channels per 128-byte K row, kind::f16 MMAs, K = 16 each), the
// epilogue stores fp16 (or fp32 for the few fp32 consumers) and reads fp16 residuals.  Everything byte-based is unchanged.
template <int kHalo, int kCS, bool kPair, int kEpi, int kEW = 8, bool kX3 = false, bool kF16 = false>
__global__ void __launch_bounds__(threads_for(kEW), 1) gemm_conv_tc_kernel(const __grid_constant__ TcParams p) {
  static_assert(!kX3 || (!kPair && kHalo != 2 && kEW == 8), "3xTF32 supports the plain and halo box modes with 8 epilogue warps");
  static_assert(!(kX3 && kF16), "3xTF32 and fp16 are different datapaths");
  constexpr int kCh = kF16 ? 2 * kChunk : kChunk;          // elements per 128-byte K row
  constexpr bool kFastEpi = kEpi == 1 || kEpi == 3;
  constexpr bool kQkv = kEpi == 3;
  constexpr bool kResHoist = kEpi == 2;
  constexpr int kEpiWarps = kEW;
  extern __shared__ __align__(1024) uint8_t smem[];      // SWIZZLE_128B atoms need a 1024-byte aligned base
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  const int kEpiOff = p.epi_off;
  uint64_t* fullA = bars;                                  // [kMaxASlots]
  uint64_t* emptyA = fullA + kMaxASlots;                   // [kMaxASlots]
  uint64_t* fullB = emptyA + kMaxASlots;                   // [kMaxBSlots]
  uint64_t* emptyB = fullB + kMaxBSlots;                   // [kMaxBSlots]
  uint64_t* tfull = emptyB + kMaxBSlots;                   // [2]
  uint64_t* tempty = tfull + 2;                            // [2]
  uint64_t* convA = tempty + 3;                            // [kMaxASlots] 3xTF32: a_lo of the slot is written (both CTAs of a pair)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(convA + kMaxASlots);   // tempty[2] is a scratch barrier for experiments

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int cs = kCS;
  const int rank = kCS > 1 ? (int)cluster_ctarank() : 0;
  const int num_clusters = gridDim.x / cs, cluster_id = blockIdx.x / cs;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxASlots; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); mbar_init(&convA[s], 2 * cs); }
    for (int s = 0; s < kMaxBSlots; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiWarps * cs); }   // leader's tempty collects both CTAs' epilogues
    mbar_init(&tempty[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {   // whole warp: allocate all 512 TMEM columns (one CTA per SM); in 2-CTA mode both CTAs of the pair allocate
    if (cs == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();          // peers' barriers must exist before any multicast arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch (common.cuh): everything above touched no global memory
  pdl_launch_dependents();
  pdl_wait();

  // K is walked as "A groups" (one activation box in smem) x "steps" (one weight tile, 4 MMAs each):
  //   halo mode : group = (source, chunk, kx), steps ky = 0..2 reuse the box at row offset ky*TW
  //   otherwise : group = (tap, source, chunk), a single step
  constexpr int steps_per_group = kHalo == 2 ? 9 : (kHalo == 1 ? 3 : 1);
  const int b_rows = p.block_n / cs;                         // rows of the B tile this CTA fetches
  const int kASlots = p.a_slots, kBSlots = p.b_slots;
  uint8_t* const ringA = smem;
  uint8_t* const ringAlo = smem + p.a_lo_off;
  uint8_t* const ringB = smem + p.b_off;

  if (warp == 0 && !kX3 && kHalo != 2 && p.grp) {
    // ======================================= TMA producer, grouped steps =========================
    int as_ = 0, bs_ = 0;
    uint32_t aph_ = 0, bph_ = 0;
    const int grp_tiles = kHalo == 1 ? 3 : 2;                              // weight tiles per slot
    for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters) {
      int n_tile, b, oy0, ox0;
      tile_coords(p, ct, cs, rank, n_tile, b, oy0, ox0);
      const int n0 = n_tile * p.block_n + rank * b_rows;
      auto load_b = [&](int k0, int k2) {                                  // 3-D box {chunk, b_rows, grp_tiles} -> slot bs_
        mbar_wait(&emptyB[bs_], bph_ ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&fullB[bs_], p.block_n * 128 * grp_tiles);
          uint8_t* dst = ringB + bs_ * p.b_slot_bytes;
          if (cs == 2) tma_load_3d_2cta(dst, &p.mapB, &fullB[bs_], k0, n0, k2);
          else tma_load_3d(dst, &p.mapB, &fullB[bs_], k0, n0, k2);
        }
        __syncwarp();
        if (++bs_ == kBSlots) { bs_ = 0; bph_ ^= 1; }
      };
      if (kHalo == 1) {
        int cbase = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const CUtensorMap* mapA = s == 0 ? &p.mapA[0] : (s == 1 ? &p.mapA[1] : (s == 2 ? &p.mapA[2] : &p.mapA[3]));
          for (int c = 0; c < p.chunks[s]; ++c)
            for (int kx_i = 0; kx_i < 3; ++kx_i) {
              mbar_wait(&emptyA[as_], aph_ ^ 1);
              if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&fullA[as_], p.a_bytes * cs);
                if (cs == 2) tma_load_4d_2cta(ringA + as_ * p.a_slot_bytes, mapA, &fullA[as_], c * kCh, ox0 + kx_i - 1, oy0 - 1, b);
                else tma_load_4d(ringA + as_ * p.a_slot_bytes, mapA, &fullA[as_], c * kCh, ox0 + kx_i - 1, oy0 - 1, b);
              }
              __syncwarp();
              if (++as_ == kASlots) { as_ = 0; aph_ ^= 1; }
              load_b((kx_i * p.sum_chunks + cbase + c) * kCh, 0);          // the three vertical taps of (chunk, kx)
            }
          cbase += p.chunks[s];
        }
      } else {
        for (int tap_o = 0; tap_o < p.ntaps; ++tap_o) {
          const int ky = p.ksize == 3 ? tap_o / 3 : 0, kx = p.ksize == 3 ? tap_o % 3 : 0;
          const int iy = oy0 * p.stride + ky * p.dil - p.pad, ix = ox0 * p.stride + kx * p.dil - p.pad;
          int s = 0, c = 0;
          for (int gc = 0; gc < p.sum_chunks; gc += 2) {
            const int npair = min(2, p.sum_chunks - gc);
            mbar_wait(&emptyA[as_], aph_ ^ 1);
            if (elect_one() && rank == 0) mbar_expect_tx(&fullA[as_], p.a_bytes * cs * npair);
            __syncwarp();
            for (int j = 0; j < npair; ++j) {
              const CUtensorMap* mapA = s == 0 ? &p.mapA[0] : (s == 1 ? &p.mapA[1] : (s == 2 ? &p.mapA[2] : &p.mapA[3]));
              if (elect_one()) {
                uint8_t* dst = ringA + as_ * p.a_slot_bytes + j * p.a_bytes;
                if (cs == 2) tma_load_4d_2cta(dst, mapA, &fullA[as_], c * kCh, ix, iy, b);
                else tma_load_4d(dst, mapA, &fullA[as_], c * kCh, ix, iy, b);
              }
              __syncwarp();
              if (++c == p.chunks[s]) { c = 0; ++s; }
            }
            if (++as_ == kASlots) { as_ = 0; aph_ ^= 1; }
            load_b(0, tap_o * p.sum_chunks + gc);                          // the weight tiles of the two chunks
          }
        }
      }
    }
  } else if (warp == 0) {
    // ======================================= TMA producer =======================================
    {
      int as_ = 0, bs_ = 0;
      uint32_t aph_ = 0, bph_ = 0;                            // ring slot + phase bit, advanced without div/mod
      for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters) {
        int n_tile, b, oy0, ox0;
        tile_coords(p, ct, cs, rank, n_tile, b, oy0, ox0);
        const int outer = kHalo ? 1 : p.ntaps;
        for (int tap_o = 0; tap_o < outer; ++tap_o) {
          int cbase = 0;
          for (int s = 0; s < p.nsrc; ++s) {
            const CUtensorMap* mapA = s == 0 ? &p.mapA[0] : (s == 1 ? &p.mapA[1] : (s == 2 ? &p.mapA[2] : &p.mapA[3]));
            for (int c = 0; c < p.chunks[s]; ++c) {
              constexpr int inner = kHalo == 1 ? 3 : 1;
              for (int kx_i = 0; kx_i < inner; ++kx_i) {
                int ix, iy;
                if (kHalo) {
                  ix = ox0 + kx_i - 1;
                  iy = oy0 - 1;
                } else {
                  const int ky = p.ksize == 3 ? tap_o / 3 : 0, kx = p.ksize == 3 ? tap_o % 3 : 0;
                  iy = oy0 * p.stride + ky * p.dil - p.pad;
                  ix = ox0 * p.stride + kx * p.dil - p.pad;
                }
                mbar_wait(&emptyA[as_], aph_ ^ 1);
                if (elect_one()) {
                  if (kX3) {          // every CTA's box completes on its OWN barrier: its converter warps wait there
                    mbar_expect_tx(&fullA[as_], p.a_bytes);
                    tma_load_4d(ringA + as_ * p.a_slot_bytes, mapA, &fullA[as_], c * kCh, ix, iy, b);
                  } else {
                    if (rank == 0) mbar_expect_tx(&fullA[as_], p.a_bytes * cs);          // leader arms for both CTAs' boxes
                    if (cs == 2) tma_load_4d_2cta(ringA + as_ * p.a_slot_bytes, mapA, &fullA[as_], c * kCh, ix, iy, b);
                    else tma_load_4d(ringA + as_ * p.a_slot_bytes, mapA, &fullA[as_], c * kCh, ix, iy, b);
                  }
                }
                __syncwarp();
                for (int st = 0; st < steps_per_group; ++st) {
                  const int tap = kHalo == 2 ? st : (kHalo == 1 ? st * 3 + kx_i : tap_o);
                  const int kb = tap * p.sum_chunks + cbase + c;
                  constexpr int kParts = kX3 ? 2 : 1;          // 3xTF32: the hi and the lo tile of the chunk, one ring slot each
#pragma unroll
                  for (int part = 0; part < kParts; ++part) {
                    mbar_wait(&emptyB[bs_], bph_ ^ 1);
                    if (elect_one()) {
                      if (rank == 0) mbar_expect_tx(&fullB[bs_], p.block_n * 128);         // both halves of the weight tile
                      uint8_t* dst = ringB + bs_ * p.b_slot_bytes;
                      const int kcol = (kb * kParts + part) * kCh;
                      if (cs == 2)      // this CTA holds columns [rank*N/2, +N/2) of the weight tile
                        tma_load_2d_2cta(dst, &p.mapB, &fullB[bs_], kcol, n_tile * p.block_n + rank * b_rows);
                      else
                        tma_load_2d(dst, &p.mapB, &fullB[bs_], kcol, n_tile * p.block_n);
                    }
                    __syncwarp();
                    if (++bs_ == kBSlots) { bs_ = 0; bph_ ^= 1; }
                  }
                }
                if (++as_ == kASlots) { as_ = 0; aph_ ^= 1; }
              }
            }
            cbase += p.chunks[s];
          }
        }
      }
    }
  } else if (warp == 1 && kX3) {
    // ======================================= MMA issuer, 3xTF32 ==================================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_tf32(p.block_n, cs == 2 ? 256 : kBlockM);
      const int groups = (kHalo == 1 ? 3 : p.ntaps) * p.sum_chunks;
      const uint32_t a_step = kHalo == 1 ? (uint32_t)(p.TW * 128) >> 4 : 0;
      const uint32_t lo_delta = (uint32_t)p.a_lo_off >> 4;                        // descriptor distance raw box -> its a_lo copy
      uint32_t tcount = 0;
      int as_ = 0, bs_ = 0;
      uint32_t aph_ = 0, bph_ = 0;
      for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters, ++tcount) {
        // The tensor core adds into its fp32 accumulator with TRUNCATION (measured: the error against exact accumulation grows by
        // ~1.3e-8 of the accumulator per K = 8 MMA).  The two small cross terms therefore go to their OWN accumulator (columns
        // [256, 512)): 1/3 of the roundings on the large sum, and the small sum's roundings are 2^-11 times smaller.  The two TMEM
        // halves are one tile, so this mode has a single accumulator stage; the epilogue adds them in fp32.
        const uint32_t as = 0, aph = tcount & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base, tmem_c = tmem_base + kMaxBlockN;
        uint32_t first = 1;
        int g = 0;
        const int outer = kHalo ? 1 : p.ntaps;
        constexpr int inner = kHalo == 1 ? 3 : 1;
        for (int tap_o = 0; tap_o < outer; ++tap_o)
        for (int s = 0; s < p.nsrc; ++s)
        for (int c = 0; c < p.chunks[s]; ++c) {
          const int nmma = (c == p.chunks[s] - 1) ? p.last_mmas[s] : 4;
        for (int kx_i = 0; kx_i < inner; ++kx_i, ++g) {
          mbar_wait(&convA[as_], aph_);                                           // raw boxes landed AND a_lo written, in both CTAs
          tc_fence_after();
          const uint64_t adesc0 = make_smem_desc(smem_u32(ringA + as_ * p.a_slot_bytes));
          for (int st = 0; st < steps_per_group; ++st) {
            int bs_lo = bs_ + 1;
            uint32_t bph_lo = bph_;
            if (bs_lo == kBSlots) { bs_lo = 0; bph_lo ^= 1; }
            mbar_wait(&fullB[bs_], bph_);
            mbar_wait(&fullB[bs_lo], bph_lo);
            tc_fence_after();
            const bool last_step = st == steps_per_group - 1;
            const bool last_of_tile = last_step && g == groups - 1;
            const uint64_t a_hi = adesc0 + (uint64_t)(st * a_step), a_lo = a_hi + lo_delta;
            const uint64_t b_hi = make_smem_desc(smem_u32(ringB + bs_ * p.b_slot_bytes)), b_lo = make_smem_desc(smem_u32(ringB + bs_lo * p.b_slot_bytes));
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < nmma) {
                  const uint64_t o = (uint64_t)(j * 2);
                  const uint32_t acc = (first && j == 0) ? 0u : 1u;
                  if (cs == 2) {
                    tc_mma_tf32_2cta(tmem_c, a_hi + o, b_lo + o, idesc, acc);
                    tc_mma_tf32_2cta(tmem_c, a_lo + o, b_hi + o, idesc, 1u);
                    tc_mma_tf32_2cta(tmem_d, a_hi + o, b_hi + o, idesc, acc);
                  } else {
                    tc_mma_tf32(tmem_c, a_hi + o, b_lo + o, idesc, acc);
                    tc_mma_tf32(tmem_c, a_lo + o, b_hi + o, idesc, 1u);
                    tc_mma_tf32(tmem_d, a_hi + o, b_hi + o, idesc, acc);
                  }
                }
              if (cs == 2) {
                tc_commit_2cta(&emptyB[bs_]);
                tc_commit_2cta(&emptyB[bs_lo]);
                if (last_step) tc_commit_2cta(&emptyA[as_]);
                if (last_of_tile) tc_commit_2cta(&tfull[as]);
              } else {
                tc_commit(&emptyB[bs_]);
                tc_commit(&emptyB[bs_lo]);
                if (last_step) tc_commit(&emptyA[as_]);
                if (last_of_tile) tc_commit(&tfull[as]);
              }
            }
            __syncwarp();
            first = 0;
            bs_ = bs_lo + 1; bph_ = bph_lo;
            if (bs_ == kBSlots) { bs_ = 0; bph_ ^= 1; }
          }
          if (++as_ == kASlots) { as_ = 0; aph_ ^= 1; }
        }
        }
      }
    }
  } else if (kX3 && (warp == 2 || warp == 3)) {
    // ======================================= a_lo converters (3xTF32) ===========================
    // Two warps, half a box each: a_lo = rna_tf32(a - trunc_tf32(a)), written at the same offset of the a_lo ring (identical
    // swizzle), published to the async proxy, then one arrival per warp on the LEADER's convA barrier.
    const int half = warp - 2;
    const int groups = (kHalo == 1 ? 3 : p.ntaps) * p.sum_chunks;
    const int vec_per_half = p.a_bytes >> 5;                                      // 16-byte vectors in half a box
    int as_ = 0;
    uint32_t aph_ = 0;
    for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters) {
      for (int g = 0; g < groups; ++g) {
        mbar_wait(&fullA[as_], aph_);
        const float4* src = reinterpret_cast<const float4*>(ringA + as_ * p.a_slot_bytes) + half * vec_per_half;
        float4* dst = reinterpret_cast<float4*>(ringAlo + as_ * p.a_slot_bytes) + half * vec_per_half;
        for (int v = lane; v < vec_per_half; v += 32) {
          const float4 a = src[v];
          float4 lo;
          lo.x = round_tf32_if(a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u), true);
          lo.y = round_tf32_if(a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u), true);
          lo.z = round_tf32_if(a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u), true);
          lo.w = round_tf32_if(a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u), true);
          dst[v] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) {
          if (cs == 2) mbar_arrive_leader(&convA[as_]); else mbar_arrive(&convA[as_]);
        }
        if (++as_ == kASlots) { as_ = 0; aph_ ^= 1; }
      }
    }
  } else if (warp == 1 && !kX3 && kHalo != 2 && p.grp) {
    // ======================================= MMA issuer, grouped steps ===========================
    if (rank == 0) {
      const uint32_t idesc = kF16 ? make_idesc_f16(p.block_n, cs == 2 ? 256 : kBlockM) : make_idesc_tf32(p.block_n, cs == 2 ? 256 : kBlockM);
      const uint32_t a_step = kHalo == 1 ? (uint32_t)(p.TW * 128) >> 4 : (uint32_t)p.a_bytes >> 4;   // next vertical tap / next chunk's box
      const uint32_t b_step = (uint32_t)(b_rows * 128) >> 4;                                          // next weight tile of the slot
      const int ngroups = kHalo == 1 ? 3 * p.sum_chunks : p.ntaps * ((p.sum_chunks + 1) >> 1);
      uint32_t tcount = 0;
      int as_ = 0, bs_ = 0;
      uint32_t aph_ = 0, bph_ = 0;
      for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters, ++tcount) {
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kMaxBlockN;
        mbar_wait(&fullA[as_], aph_);
        mbar_wait(&fullB[bs_], bph_);
        int s = 0, c = 0, sub = 0;                 // running (source, chunk); halo: sub = kx of the group
        for (int g = 0; g < ngroups; ++g) {
          int nm0, nm1 = 0, nparts;
          if (kHalo == 1) {
            nm0 = (c == p.chunks[s] - 1) ? p.last_mmas[s] : 4;
            nparts = 3;
          } else {
            nm0 = (c == p.chunks[s] - 1) ? p.last_mmas[s] : 4;
            int s1 = s, c1 = c + 1;
            if (c1 == p.chunks[s1]) { c1 = 0; ++s1; }
            const int gc = g % ((p.sum_chunks + 1) >> 1);
            nparts = (2 * gc + 1 < p.sum_chunks) ? 2 : 1;
            if (nparts == 2) nm1 = (c1 == p.chunks[s1] - 1) ? p.last_mmas[s1] : 4;
          }
          int as_n = as_ + 1, bs_n = bs_ + 1;
          uint32_t aph_n = aph_, bph_n = bph_;
          if (as_n == kASlots) { as_n = 0; aph_n ^= 1; }
          if (bs_n == kBSlots) { bs_n = 0; bph_n ^= 1; }
          const bool last = g == ngroups - 1;
          uint32_t okA = 1, okB = 1;
          if (!last) { okA = mbar_test_wait(&fullA[as_n], aph_n); okB = mbar_test_wait(&fullB[bs_n], bph_n); }
          tc_fence_after();
          const uint64_t adesc0 = make_smem_desc(smem_u32(ringA + as_ * p.a_slot_bytes));
          const uint64_t bdesc0 = make_smem_desc(smem_u32(ringB + bs_ * p.b_slot_bytes));
          if (elect_one()) {
#pragma unroll
            for (int part = 0; part < 3; ++part)
              if (part < nparts) {
                const int nm = (kHalo == 1 || part == 0) ? nm0 : nm1;
                const uint64_t adesc = adesc0 + (uint64_t)(part * a_step), bdesc = bdesc0 + (uint64_t)(part * b_step);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < nm) tc_mma_any<kF16, kCS>(tmem_d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (g == 0 && part == 0 && j == 0) ? 0u : 1u);
                if (kPair) {
                  const uint64_t adesc2 = adesc + (uint64_t)((kBlockM * 128) >> 4);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (j < nm) tc_mma_any<kF16, kCS>(tmem_d + 128, adesc2 + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (g == 0 && part == 0 && j == 0) ? 0u : 1u);
                }
              }
            if (cs == 2) {
              tc_commit_2cta(&emptyB[bs_]);
              tc_commit_2cta(&emptyA[as_]);
              if (last) tc_commit_2cta(&tfull[as]);
            } else {
              tc_commit(&emptyB[bs_]);
              tc_commit(&emptyA[as_]);
              if (last) tc_commit(&tfull[as]);
            }
          }
          __syncwarp();
          // advance the running chunk
          if (kHalo == 1) {
            if (++sub == 3) { sub = 0; if (++c == p.chunks[s]) { c = 0; ++s; } }
          } else {
            for (int j = 0; j < nparts; ++j)
              if (++c == p.chunks[s]) { c = 0; ++s; }
            if (s >= p.nsrc) { s = 0; c = 0; }     // next tap walks the sources again
          }
          as_ = as_n; aph_ = aph_n; bs_ = bs_n; bph_ = bph_n;
          if (!okA) mbar_wait(&fullA[as_], aph_);
          if (!okB) mbar_wait(&fullB[bs_], bph_);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =========================================
    if (rank == 0) {   // 2-CTA mode: the leader issues M = 256 MMAs that read both CTAs' shared memory
      const uint32_t idesc = kF16 ? make_idesc_f16(p.block_n, cs == 2 ? 256 : kBlockM) : make_idesc_tf32(p.block_n, cs == 2 ? 256 : kBlockM);
      const int groups = (kHalo == 2 ? 1 : (kHalo == 1 ? 3 : p.ntaps)) * p.sum_chunks;
      const uint32_t a_step = kHalo == 1 ? (uint32_t)(p.TW * 128) >> 4 : 0;      // descriptor units of 16 B per vertical tap
      const uint32_t a_sbo = kHalo == 2 ? (uint32_t)((p.TW + 2) * 128) : 1024u;   // full-halo box: tile rows are TW+2 pixels apart
      uint32_t tcount = 0;
      int as_ = 0, bs_ = 0;
      uint32_t aph_ = 0, bph_ = 0;
      for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters, ++tcount) {
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kMaxBlockN;
        // The issuing warp stalls ~100 cycles on every mbarrier probe and the tensor pipe drains meanwhile, so
        // the state of the NEXT step's barriers is probed before this step's MMAs are issued and only
        // re-checked (normally already true) afterwards.
        uint32_t first = 1;
        mbar_wait(&fullA[as_], aph_);
        mbar_wait(&fullB[bs_], bph_);
        int g = 0;
        const int outer = kHalo ? 1 : p.ntaps;
        const int inner = kHalo == 1 ? 3 : 1;
        for (int tap_o = 0; tap_o < outer; ++tap_o)
        for (int s = 0; s < p.nsrc; ++s)
        for (int c = 0; c < p.chunks[s]; ++c) {
          const int nmma = (c == p.chunks[s] - 1) ? p.last_mmas[s] : 4;     // skip all-zero K slices of a partial chunk
        for (int kx_i = 0; kx_i < inner; ++kx_i, ++g) {
          const uint64_t adesc0 = make_smem_desc(smem_u32(ringA + as_ * p.a_slot_bytes), a_sbo);
          int as_n = as_ + 1;
          uint32_t aph_n = aph_;
          if (as_n == kASlots) { as_n = 0; aph_n ^= 1; }
          for (int st = 0; st < steps_per_group; ++st) {
            int bs_n = bs_ + 1;
            uint32_t bph_n = bph_;
            if (bs_n == kBSlots) { bs_n = 0; bph_n ^= 1; }
            const bool last_step = st == steps_per_group - 1;
            const bool last_of_tile = last_step && g == groups - 1;
            uint32_t okB = 1, okA = 1;
            if (!last_of_tile) {
              okB = mbar_test_wait(&fullB[bs_n], bph_n);
              if (last_step) okA = mbar_test_wait(&fullA[as_n], aph_n);
            }
            tc_fence_after();
            const uint64_t adesc = adesc0 + (kHalo == 2 ? (uint64_t)((((st / 3) * (p.TW + 2) + (st % 3)) * 128) >> 4) : (uint64_t)(st * a_step));
            const uint64_t bdesc = make_smem_desc(smem_u32(ringB + bs_ * p.b_slot_bytes));
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)   // up to 4 x (K = 8 tf32 / 16 fp16 = 32 bytes) inside the 128-byte swizzle row
                if (j < nmma) tc_mma_any<kF16, kCS>(tmem_d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (first && j == 0) ? 0u : 1u);
              if (kPair) {                  // second 128-pixel tile of the pair: next TH rows of the same box, same weights
                const uint64_t adesc2 = adesc + (uint64_t)((kBlockM * 128) >> 4);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < nmma) tc_mma_any<kF16, kCS>(tmem_d + 128, adesc2 + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), idesc, (first && j == 0) ? 0u : 1u);
              }
              if (cs == 2) {
                tc_commit_2cta(&emptyB[bs_]);
                if (last_step) tc_commit_2cta(&emptyA[as_]);   // frees the activation boxes of both CTAs once the MMAs retire
                if (last_of_tile) tc_commit_2cta(&tfull[as]);  // accumulators complete in both CTAs
              } else {
                tc_commit(&emptyB[bs_]);
                if (last_step) tc_commit(&emptyA[as_]);
                if (last_of_tile) tc_commit(&tfull[as]);
              }
            }
            __syncwarp();
            first = 0;
            bs_ = bs_n; bph_ = bph_n;
            if (!okB) mbar_wait(&fullB[bs_], bph_);
            if (last_step && !okA) mbar_wait(&fullA[as_n], aph_n);
          }
          as_ = as_n; aph_ = aph_n;
        }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue ===========================================
    // Warp w drains TMEM lanes [32*(w%4), +32) (= 32 output pixels); the two warps of a quarter alternate over the
    // 32-column chunks.  Accumulators pass through a padded per-warp smem tile so that global traffic is
    // row-coalesced: 8 lanes x float4 cover 128 contiguous bytes of one output row; bias / PReLU slopes are
    // fetched once per lane per chunk.
    const int ew = warp - 4;                                  // 0..kEW-1
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const int half = ew >> 2;                                 // which of the kEW/4 interleaved chunk sequences this warp takes
    constexpr int kSeq = kEW / 4;
    float* stage = reinterpret_cast<float*>(smem + kEpiOff + ew * kEpiWarpBytes);
    int2* s_row = reinterpret_cast<int2*>(smem + kEpiOff + ew * kEpiWarpBytes + 32 * kEpiPitch * 4);   // (orow, m) per row
    const int m_local = q * 32 + lane;
    const int th = m_local / p.TW, tw = m_local % p.TW;
    const int l8 = lane & 7, rsub = lane >> 3, col = 4 * l8;
    const EpiParams& e = p.epi;
    const bool rnd = e.round != 0;
    uint32_t tcount = 0;
    for (int ct = cluster_id; ct < p.total_ctiles; ct += num_clusters, ++tcount) {
      const uint32_t as = kX3 ? 0 : (tcount & 1), aph = kX3 ? (tcount & 1) : ((tcount >> 1) & 1);   // 3xTF32: one stage = main + correction halves
      int n_tile, b, oy0, ox0;
      tile_coords(p, ct, cs, rank, n_tile, b, oy0, ox0);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const int nchunks = (p.block_n + 31) >> 5;
      const int nunits = kPair ? 2 * nchunks : nchunks;       // (sub-tile, 32-column chunk) units, split between the two warps of a quarter
      int last_q = -1, last_t = -1;
      int64_t m = 0;
      bool row_ok = false;
      if (kEpi == 4) {
        // ---------------- TMA-store epilogue ----------------
        // A TMEM lane is an output pixel, so after tcgen05.ld every lane holds 32 consecutive channels of ITS pixel: bias / PReLU /
        // narrowing happen in registers, the lane writes its 32-channel row into the warp's staging tile (swizzled like the
        // output tensor map: conflict-free 16-byte stores) and one elected lane hands the 32-pixel x 32-channel box to the TMA
        // unit.  No per-element address arithmetic, no bounds checks (the TMA unit clips the box to the tensor: image border,
        // row window, channels beyond Cout, phantom tiles), no global store instructions.  ConvTranspose k2 s2: the same box
        // walks the output with element stride 2.
        uint8_t* const sbuf = smem + kEpiOff + ew * kEpiWarpBytes;
        const int es = (kF16 && e.out_half) ? 2 : 4;
        const int qx = (q * 32) % p.TW, qy = (q * 32) / p.TW;          // where this warp's 32 pixels sit inside the 128-pixel tile
        for (int u = half; u < nunits; u += kSeq) {
          const int t = kPair ? u / nchunks : 0;
          const int c0 = (kPair ? u - t * nchunks : u) * 32;
          const int n0 = n_tile * p.block_n + c0;
          int sq = 0, co0 = n0;
          if (p.st_shuffle) { sq = n0 / p.cq_pad; co0 = n0 - sq * p.cq_pad; }
          if (co0 >= e.Cout || sq > 3) continue;
          const int width = min(32, p.block_n - c0);                    // 32, or a 16-column tail (block_n is a multiple of 16)
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * kMaxBlockN + t * 128;
          uint32_t r[32];
          __syncwarp();
          if (p.dbg >= 3) continue;
          tc_ld32(trow + c0, r);
          tc_wait_ld();
          int cx = ox0 + qx, cy = oy0 + t * p.TH + qy;
          if (p.st_shuffle) { cx = 2 * cx + (sq & 1); cy = 2 * cy + (sq >> 1); }
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {          // bias / slope arrays are padded to a multiple of 32 floats by the caller
            const float4 bz = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + co0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[c] = __uint_as_float(r[c]) + bz.x; v[c + 1] = __uint_as_float(r[c + 1]) + bz.y;
            v[c + 2] = __uint_as_float(r[c + 2]) + bz.z; v[c + 3] = __uint_as_float(r[c + 3]) + bz.w;
            if (e.prelu) {
              const float4 sl = __ldg(reinterpret_cast<const float4*>(e.prelu + co0 + c));
              v[c] = v[c] > 0.f ? v[c] : v[c] * sl.x; v[c + 1] = v[c + 1] > 0.f ? v[c + 1] : v[c + 1] * sl.y;
              v[c + 2] = v[c + 2] > 0.f ? v[c + 2] : v[c + 2] * sl.z; v[c + 3] = v[c + 3] > 0.f ? v[c + 3] : v[c + 3] * sl.w;
            }
          }
          if (kF16 && e.head32 && co0 + 31 >= e.head32_c0) {             // fp32 copy of the motion channels: a handful of scalars per pixel
            const int oy = oy0 + t * p.TH + th, ox = ox0 + tw;
            if (b < p.B && oy < p.row1 && ox < e.Wout) {
              float* hp = e.head32 + (((int64_t)b * e.Hout + oy) * e.Wout + ox) * e.head32_pitch;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (co0 + j >= e.head32_c0 && co0 + j < e.Cout) hp[co0 + j - e.head32_c0] = v[j];
            }
          }
          // one pass per destination: the output, then the PReLU'd second output of the decoder levels
          const int npass = e.out2 ? 2 : 1;
          for (int pass = 0; pass < npass; ++pass) {
            if (pass == 1) {
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                const float4 sl = __ldg(reinterpret_cast<const float4*>(e.prelu2 + co0 + c));
                v[c] = v[c] > 0.f ? v[c] : v[c] * sl.x; v[c + 1] = v[c + 1] > 0.f ? v[c + 1] : v[c + 1] * sl.y;
                v[c + 2] = v[c + 2] > 0.f ? v[c + 2] : v[c + 2] * sl.z; v[c + 3] = v[c + 3] > 0.f ? v[c + 3] : v[c + 3] * sl.w;
              }
            }
            const int pes = (pass == 1 && kF16) ? 2 : es;                 // out2 is always an activation map
            const int rb = width * pes;                                   // bytes per staged row: 128 / 64 / 32
            const uint32_t swz = (uint32_t)((lane * rb) >> 7) & (uint32_t)((rb >> 4) - 1);   // SWIZZLE_<rb>B: 16-byte chunk index ^ address bits [7, ..)
            if (p.dbg >= 2) { if (v[0] == 1.2345e-30f) sbuf[lane] = 1; continue; }
            if (lane == 0) tma_store_wait_read();                         // the previous box has left the staging tile
            __syncwarp();
            uint8_t* rowp = sbuf + lane * rb;
            if (pes == 2) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c * 8 < width) {
                  uint4 w;
                  *reinterpret_cast<__half2*>(&w.x) = __floats2half2_rn(v[8 * c], v[8 * c + 1]);
                  *reinterpret_cast<__half2*>(&w.y) = __floats2half2_rn(v[8 * c + 2], v[8 * c + 3]);
                  *reinterpret_cast<__half2*>(&w.z) = __floats2half2_rn(v[8 * c + 4], v[8 * c + 5]);
                  *reinterpret_cast<__half2*>(&w.w) = __floats2half2_rn(v[8 * c + 6], v[8 * c + 7]);
                  *reinterpret_cast<uint4*>(rowp + (((uint32_t)c ^ swz) << 4)) = w;
                }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                if (c * 4 < width)
                  *reinterpret_cast<float4*>(rowp + (((uint32_t)c ^ swz) << 4)) =
                      round_tf32_if(make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]), rnd);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && p.dbg < 1) {
              const CUtensorMap* mp = pass == 0 ? (width == 32 ? &p.mapOut : &p.mapOutTail) : (width == 32 ? &p.mapOut2 : &p.mapOut2Tail);
              tma_store_4d(mp, sbuf, co0, cx, cy, b);
              tma_store_commit();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (cs == 2) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]);
        }
        continue;
      }
      for (int u = half; u < nunits; u += kSeq) {
        const int t = kPair ? u / nchunks : 0;                // sub-tile of the pair (warp-uniform)
        const int c0 = (kPair ? u - t * nchunks : u) * 32;
        if (t != last_t) {
          const int oy = oy0 + t * p.TH + th, ox = ox0 + tw;
          row_ok = b < p.B && oy < p.row1 && ox < e.Wout;
          m = ((int64_t)b * e.Hout + oy) * e.Wout + ox;
          last_t = t;
          last_q = -1;
        }
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * kMaxBlockN + t * 128;
        const int n0 = n_tile * p.block_n + c0;               // first GEMM column of this chunk (warp-uniform)
        int sq = 0, co0 = n0;
        if (e.out_mode == ATMVFI_OUT_SHUFFLE2) { sq = n0 / p.cq_pad; co0 = n0 - sq * p.cq_pad; }
        if (co0 >= (kFastEpi ? e.cout4 : e.Cout) || sq > 3) continue;                // padding columns: nothing to store (uniform branch)
        uint32_t r[32];
        __syncwarp();
        tc_ld32(trow + c0, r);
        tc_wait_ld();
        if (kX3) {              // + the cross-term accumulator
          uint32_t rc[32];
          tc_ld32(trow + kMaxBlockN + c0, rc);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rc[j]));
        }
        if (kQkv && n0 >= 2 * e.qkv_C) {
          // V^T[h][d][r]: consecutive GEMM rows are contiguous for a fixed column, and a TMEM lane IS a row - store straight
          // from the accumulator registers, 32 rows x 4 bytes = one 128-byte line per column (no shared-memory transpose)
          if (row_ok) {
            const int nv = min(min(32, p.block_n - c0), e.Cout - n0);
            float* vb = e.out + 2 * (int64_t)e.qkv_C * e.qkv_R + (int64_t)(n0 - 2 * e.qkv_C) * e.qkv_R + m;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv) vb[(int64_t)j * e.qkv_R] = round_tf32_if(__uint_as_float(r[j]) + (e.bias ? __ldg(e.bias + n0 + j) : 0.f), rnd);
          }
          continue;
        }
        if (!kFastEpi && sq != last_q) {
          const int64_t orow = row_ok ? epi_out_row(e, m, sq) : -1;
          s_row[lane] = make_int2((int)orow, (int)m);
          last_q = sq;
        }
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          *reinterpret_cast<float4*>(&stage[lane * kEpiPitch + c]) =
              make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        __syncwarp();
        const int nvalid = min(min(32, p.block_n - c0), e.Cout - co0);   // block_n may end inside this 32-column chunk
        const bool full4 = col + 4 <= nvalid;
        if (kFastEpi) {
          // layers without residual and with pixel-major output: rows are addressed arithmetically (TW is a power of two), nothing
          // but the accumulator tile is read from shared memory.  Channel counts that are not multiples of 4 (101, 197, 389 ...)
          // are stored as whole 4-channel vectors up to e.cout4: the pad lanes of the map receive zeros (zero weight rows, zero-
          // padded bias).  Optional second PReLU'd output and (fp16 mode) fp32 copy of the motion channels.
          const int nvalid_f = min(min(32, p.block_n - c0), e.cout4 - co0);
          if (col < nvalid_f) {
            float4 bz4 = make_float4(0.f, 0.f, 0.f, 0.f), sl4 = make_float4(1.f, 1.f, 1.f, 1.f), sl24 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (e.bias) bz4 = __ldg(reinterpret_cast<const float4*>(e.bias + co0 + col));
            const bool act = e.prelu != nullptr;
            if (act) sl4 = __ldg(reinterpret_cast<const float4*>(e.prelu + co0 + col));
            const bool dual = !kQkv && e.out2 != nullptr;
            if (dual) sl24 = __ldg(reinterpret_cast<const float4*>(e.prelu2 + co0 + col));
            const int tw_mask = p.TW - 1, tw_shift = 31 - __clz(p.TW);
            const int oyb = oy0 + (kPair ? t * p.TH : 0);
            // QKV_HEADS (q / k columns; whole-v chunks never get here): row pitch = head dim, column offset = head plane.
            // In the one chunk that straddles the k | v boundary the v lanes store their 4 columns one by one.
            constexpr bool qkv = kQkv;
            const bool qkv_v = qkv && co0 + col >= 2 * e.qkv_C;
            float* obase = qkv ? e.out + (qkv_v ? 0 : epi_qkv_offset(e, 0, co0 + col)) : e.out + co0 + col;
            const int64_t opitch = qkv ? e.qkv_hd : e.out_pitch;
            const bool head = kF16 && !kQkv && e.head32 != nullptr && co0 + col + 3 >= e.head32_c0;
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
              const int row = rr * 4 + rsub, ml = q * 32 + row;
              const int oy = oyb + (ml >> tw_shift), ox = ox0 + (ml & tw_mask);
              if (b < p.B && oy < p.row1 && ox < e.Wout) {
                const float4 a4 = *reinterpret_cast<const float4*>(&stage[row * kEpiPitch + col]);
                float4 v = make_float4(a4.x + bz4.x, a4.y + bz4.y, a4.z + bz4.z, a4.w + bz4.w);
                if (act) {
                  v.x = v.x > 0.f ? v.x : v.x * sl4.x; v.y = v.y > 0.f ? v.y : v.y * sl4.y;
                  v.z = v.z > 0.f ? v.z : v.z * sl4.z; v.w = v.w > 0.f ? v.w : v.w * sl4.w;
                }
                const int64_t mrow = ((int64_t)b * e.Hout + oy) * e.Wout + ox;
                if (head) {            // fp32 copy of the motion channels, before any narrowing
                  const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int co = co0 + col + k;
                    if (co >= e.head32_c0 && co < e.Cout) e.head32[mrow * e.head32_pitch + (co - e.head32_c0)] = vv[k];
                  }
                }
                if (dual) {
                  float4 w2 = make_float4(v.x > 0.f ? v.x : v.x * sl24.x, v.y > 0.f ? v.y : v.y * sl24.y,
                                          v.z > 0.f ? v.z : v.z * sl24.z, v.w > 0.f ? v.w : v.w * sl24.w);
                  if (kF16) Act<__half>::st4(reinterpret_cast<__half*>(e.out2) + mrow * e.out2_pitch + co0 + col, w2);
                  else *reinterpret_cast<float4*>(e.out2 + mrow * e.out2_pitch + co0 + col) = round_tf32_if(w2, rnd);
                }
                if (kF16 && e.out_half) {         // fp16 map (never the q|k|v layout, which stays fp32)
                  Act<__half>::st4(reinterpret_cast<__half*>(e.out) + mrow * e.out_pitch + co0 + col, v);
                  continue;
                }
                v = round_tf32_if(v, rnd);
                if (qkv_v) {
                  const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) e.out[epi_qkv_offset(e, mrow, co0 + col + k)] = vv[k];
                } else {
                  *reinterpret_cast<float4*>(obase + mrow * opitch) = v;
                }
              }
            }
          }
          continue;
        }
        float bz[4] = {0.f, 0.f, 0.f, 0.f}, sl[4] = {1.f, 1.f, 1.f, 1.f}, sl2[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (col + k < nvalid) {
            if (e.bias) bz[k] = __ldg(e.bias + co0 + col + k);
            if (e.prelu) sl[k] = __ldg(e.prelu + co0 + col + k);
            if (e.out2) sl2[k] = __ldg(e.prelu2 + co0 + col + k);
          }
        if (col < nvalid) {
          // residual rows of the 8 output rows this lane serves: issued back to back so that 8 loads are in flight per
          // lane (the layers with a residual - attention proj, fc2 - are bandwidth-bound in this epilogue)
          float4 res4[kResHoist ? 8 : 1];
          const bool res_vec = kResHoist && full4;
          if (res_vec) {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
              const int2 ri = s_row[rr * 4 + rsub];
              const int64_t roff = (int64_t)ri.y * e.res_pitch + co0 + col;
              res4[rr] = ri.x < 0 ? make_float4(0.f, 0.f, 0.f, 0.f)
                                  : (kF16 ? Act<__half>::ld4(reinterpret_cast<const __half*>(e.residual) + roff) : __ldg(reinterpret_cast<const float4*>(e.residual + roff)));
            }
          }
          auto emit_row = [&](const int rr) {
            const int row = rr * 4 + rsub;
            const int2 ri = s_row[row];
            if (ri.x < 0) return;
            const float4 a4 = *reinterpret_cast<const float4*>(&stage[row * kEpiPitch + col]);
            float v[4] = {a4.x + bz[0], a4.y + bz[1], a4.z + bz[2], a4.w + bz[3]};
            if (e.residual) {
              const int64_t roff = (int64_t)ri.y * e.res_pitch + co0 + col;
              const float* rs = e.residual + roff;
              const __half* rsh = reinterpret_cast<const __half*>(e.residual) + roff;
              if (res_vec) {
                const float4 t = res4[kResHoist ? rr : 0];
                v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
              } else if (full4) {
                const float4 t = kF16 ? Act<__half>::ld4(rsh) : __ldg(reinterpret_cast<const float4*>(rs));
                v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (col + k < nvalid) v[k] += kF16 ? Act<__half>::ld(rsh + k) : __ldg(rs + k);
              }
            }
            if (e.prelu) {
#pragma unroll
              for (int k = 0; k < 4; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * sl[k];
            }
            float w[4];
            if (e.out2) {
#pragma unroll
              for (int k = 0; k < 4; ++k) w[k] = round_tf32_if(v[k] > 0.f ? v[k] : v[k] * sl2[k], rnd);
            }
            if (kF16) {
              if (e.head32) {         // fp32 copy of the motion channels (flows + occlusion logit), before any narrowing
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (col + k < nvalid && co0 + col + k >= e.head32_c0)
                    e.head32[(int64_t)ri.x * e.head32_pitch + (co0 + col + k - e.head32_c0)] = v[k];
              }
              if (e.out_half) {
                __half* o1 = reinterpret_cast<__half*>(e.out) + (int64_t)ri.x * e.out_pitch + co0 + col;
                if (full4) {
                  Act<__half>::st4(o1, make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (col + k < nvalid) o1[k] = __float2half_rn(v[k]);
                }
              }
              if (e.out2) {
                __half* o2 = reinterpret_cast<__half*>(e.out2) + (int64_t)ri.x * e.out2_pitch + co0 + col;
                if (full4) {
                  Act<__half>::st4(o2, make_float4(w[0], w[1], w[2], w[3]));
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (col + k < nvalid) o2[k] = __float2half_rn(w[k]);
                }
              }
              if (e.out_half) return;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = round_tf32_if(v[k], rnd);
            float* o1 = e.out + (int64_t)ri.x * e.out_pitch + co0 + col;
            if (full4) {
              *reinterpret_cast<float4*>(o1) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (col + k < nvalid) o1[k] = v[k];
            }
            if (e.out2 && !kF16) {
              float* o2 = e.out2 + (int64_t)ri.x * e.out2_pitch + co0 + col;
              if (full4) {
                *reinterpret_cast<float4*>(o2) = make_float4(w[0], w[1], w[2], w[3]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (col + k < nvalid) o2[k] = w[k];
              }
            }
          };
          if (kResHoist) {      // full unroll: res4[rr] must stay in registers
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) emit_row(rr);
          } else {
#pragma unroll 4
            for (int rr = 0; rr < 8; ++rr) emit_row(rr);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                         // accumulator stage drained (the leader's MMA warp waits for it)
        if (cs == 2) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]);
      }
    }
    if (kEpi == 4 && lane == 0) tma_store_wait_all();          // bulk stores of this warp are complete before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();          // no CTA may exit while a peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    if (cs == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------- kernel tables
typedef void (*TcKernelFn)(TcParams);

// Instantiations of one operand type.  tbl 0: [fast epilogue][0 plain, 1 halo, 2 full halo, 3 paired halo][cluster]; tbl 1: residual
// prefetch [cluster]; tbl 2: 16 epilogue warps [fast][0 plain, 1 paired halo][cluster]; tbl 3: head-major q|k|v [16 warps][cluster].
template <bool kF16>
struct TcKernels {
  static TcKernelFn get(int tbl, int a, int b, int c) {
    static const TcKernelFn table[2][4][2] = {
        {{gemm_conv_tc_kernel<0, 1, false, 0, 8, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 0, 8, false, kF16>},
         {gemm_conv_tc_kernel<1, 1, false, 0, 8, false, kF16>, gemm_conv_tc_kernel<1, 2, false, 0, 8, false, kF16>},
         {gemm_conv_tc_kernel<kF16 ? 1 : 2, 1, false, 0, 8, false, kF16>, gemm_conv_tc_kernel<kF16 ? 1 : 2, 2, false, 0, 8, false, kF16>},
         {gemm_conv_tc_kernel<1, 1, true, 0, 8, false, kF16>, gemm_conv_tc_kernel<1, 2, true, 0, 8, false, kF16>}},
        {{gemm_conv_tc_kernel<0, 1, false, 1, 8, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 1, 8, false, kF16>},
         {gemm_conv_tc_kernel<1, 1, false, 1, 8, false, kF16>, gemm_conv_tc_kernel<1, 2, false, 1, 8, false, kF16>},
         {gemm_conv_tc_kernel<kF16 ? 1 : 2, 1, false, 1, 8, false, kF16>, gemm_conv_tc_kernel<kF16 ? 1 : 2, 2, false, 1, 8, false, kF16>},
         {gemm_conv_tc_kernel<1, 1, true, 1, 8, false, kF16>, gemm_conv_tc_kernel<1, 2, true, 1, 8, false, kF16>}}};
    static const TcKernelFn res_table[2] = {gemm_conv_tc_kernel<0, 1, false, 2, 8, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 2, 8, false, kF16>};
    static const TcKernelFn res_table16[2] = {gemm_conv_tc_kernel<0, 1, false, 2, 16, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 2, 16, false, kF16>};
    static const TcKernelFn table16[2][2][2] = {
        {{gemm_conv_tc_kernel<0, 1, false, 0, 16, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 0, 16, false, kF16>},
         {gemm_conv_tc_kernel<1, 1, true, 0, 16, false, kF16>, gemm_conv_tc_kernel<1, 2, true, 0, 16, false, kF16>}},
        {{gemm_conv_tc_kernel<0, 1, false, 1, 16, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 1, 16, false, kF16>},
         {gemm_conv_tc_kernel<1, 1, true, 1, 16, false, kF16>, gemm_conv_tc_kernel<1, 2, true, 1, 16, false, kF16>}}};
    static const TcKernelFn qkv_table[2][2] = {{gemm_conv_tc_kernel<0, 1, false, 3, 8, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 3, 8, false, kF16>},
                                               {gemm_conv_tc_kernel<0, 1, false, 3, 16, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 3, 16, false, kF16>}};
    // TMA-store epilogue: tbl 5 [0 plain, 1 halo, 2 paired halo][cluster] with 8 warps, tbl 6 [0 plain, 1 paired halo][cluster] with 16
    static const TcKernelFn st_table[3][2] = {
        {gemm_conv_tc_kernel<0, 1, false, 4, 8, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 4, 8, false, kF16>},
        {gemm_conv_tc_kernel<1, 1, false, 4, 8, false, kF16>, gemm_conv_tc_kernel<1, 2, false, 4, 8, false, kF16>},
        {gemm_conv_tc_kernel<1, 1, true, 4, 8, false, kF16>, gemm_conv_tc_kernel<1, 2, true, 4, 8, false, kF16>}};
    static const TcKernelFn st_table16[2][2] = {
        {gemm_conv_tc_kernel<0, 1, false, 4, 16, false, kF16>, gemm_conv_tc_kernel<0, 2, false, 4, 16, false, kF16>},
        {gemm_conv_tc_kernel<1, 1, true, 4, 16, false, kF16>, gemm_conv_tc_kernel<1, 2, true, 4, 16, false, kF16>}};
    switch (tbl) {
      case 5: return st_table[a][c];
      case 6: return st_table16[a][c];
      case 0: return table[a][b][c];
      case 1: return res_table[c];
      case 2: return table16[a][b][c];
      case 4: return res_table16[c];
      default: return qkv_table[a][c];
    }
  }
  // opt in to > 48 KB of dynamic shared memory for every instantiation (per device)
  static cudaError_t configure() {
    for (int a = 0; a < 2; ++a)
      for (int c = 0; c < 2; ++c) {
        for (int b = 0; b < 4; ++b) {
          cudaError_t e = cudaFuncSetAttribute(get(0, a, b, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(8));
          if (e != cudaSuccess) return e;
        }
        for (int b = 0; b < 2; ++b) {
          cudaError_t e = cudaFuncSetAttribute(get(2, a, b, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(16));
          if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaFuncSetAttribute(get(3, a, 0, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(a ? 16 : 8));
        if (e != cudaSuccess) return e;
        for (int v = 0; v < 3; ++v) {
          e = cudaFuncSetAttribute(get(5, v, 0, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(8));
          if (e != cudaSuccess) return e;
          if (v < 2) e = cudaFuncSetAttribute(get(6, v, 0, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(16));
          if (e != cudaSuccess) return e;
        }
        if (a == 0) {
          e = cudaFuncSetAttribute(get(1, 0, 0, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(8));
          if (e != cudaSuccess) return e;
          e = cudaFuncSetAttribute(get(4, 0, 0, c), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(16));
          if (e != cudaSuccess) return e;
        }
      }
    return cudaSuccess;
  }
};

}  // namespace

#ifdef ATMVFI_TC_F16_TU
// The fp16 instantiations live in their own translation unit (gemm_conv_tc_f16.cu) so that the two halves compile in parallel.
TcKernelFn atmvfi_tc_f16_kernel(int tbl, int a, int b, int c) { return TcKernels<true>::get(tbl, a, b, c); }
cudaError_t atmvfi_tc_f16_configure() { return TcKernels<true>::configure(); }
#else
TcKernelFn atmvfi_tc_f16_kernel(int tbl, int a, int b, int c);
cudaError_t atmvfi_tc_f16_configure();

namespace {

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// Layout contract shared with atmvfi/pack.py (pack_tc): see tc_layout() there.
extern "C" int atmvfi_gemm_conv_plan_bytes(void) { return (int)sizeof(TcPlan); }

extern "C" int atmvfi_gemm_conv_plan(const atmvfi_gemm_conv_desc* d, void* plan_host) {
  ATMVFI_REQUIRE(d && plan_host, "gemm_conv_plan: null argument");
  ATMVFI_REQUIRE(((uintptr_t)plan_host & 63) == 0, "gemm_conv_plan: plan buffer must be 64-byte aligned");
  EncodeTiledFn enc = get_encode();
  ATMVFI_REQUIRE(enc != nullptr, "gemm_conv_plan: cuTensorMapEncodeTiled not available (no CUDA driver?)");
  TcPlan* pl = reinterpret_cast<TcPlan*>(plan_host);
  memset(pl, 0, sizeof(TcPlan));
  pl->nsrc = d->nsrc;
  pl->ksize = d->ksize; pl->ntaps = d->ksize * d->ksize; pl->stride = d->stride; pl->dil = d->dil;
  pl->pad = d->dil * (d->ksize - 1) / 2;
  pl->B = d->B; pl->Hout = d->Hout; pl->Wout = d->Wout;
  {
    int ny;
    ATMVFI_REQUIRE(row_window(d->Hout, d->row_begin, d->row_end, &pl->row0, &ny), "gemm_conv(tf32): bad row window [%d,%d) for Hout=%d",
                   d->row_begin, d->row_end, d->Hout);
    pl->row1 = pl->row0 + ny;
  }
  const int Hwin = pl->row1 - pl->row0;          // rows this launch produces
  ATMVFI_REQUIRE(d->stride == 1 || d->stride == 2 || d->stride == 4, "gemm_conv(tf32): stride %d unsupported", d->stride);

  // 3x3 stride-1 layers fetch activation boxes with a vertical halo and reuse them for the 3 vertical taps
  static int halo_ok = -1;
  if (halo_ok < 0) { const char* ev = getenv("ATMVFI_TC_HALO"); halo_ok = ev ? atoi(ev) : 1; }
  pl->x3 = d->precision == ATMVFI_TF32X3 ? 1 : 0;
  pl->f16 = d->precision == ATMVFI_F16 ? 1 : 0;
  pl->chunk = pl->f16 ? 2 * kChunk : kChunk;
  const int es = pl->f16 ? 2 : 4;                 // bytes per operand element
  const int pitch_align = 16 / es;
  pl->halo = (halo_ok && d->ksize == 3 && d->stride == 1 && d->dil == 1) ? (pl->x3 ? 1 : halo_ok) : 0;   // 1: 3 boxes / chunk, 2: one full-halo box
  // pixel tile TW x TH = 128: least padding waste, then squarest.  Element-strided boxes are capped at 256
  // per dimension; halo boxes need TW % 8 == 0 (vertical taps = whole swizzle atoms) and (TH+2)*TW <= 192 rows.
  int best_tw = 0;
  int64_t best_cost = -1;
  const int cands[5] = {16, 32, 8, 64, 128};
  for (int i = 0; i < 5; ++i) {
    int tw = cands[i], th = 128 / tw;
    if (tw * d->stride > 256 || th * d->stride > 256) continue;
    if (pl->halo == 1 && (th + 2) * tw * 128 > kMaxABoxBytes) continue;
    if (pl->halo == 2 && tw != 8) continue;          // full-halo box: every tile row must be one 8-row swizzle group
    int64_t cost = (int64_t)cdiv(d->Wout, tw) * cdiv(Hwin, th);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_tw = tw; }
  }
  ATMVFI_REQUIRE(best_tw > 0, "gemm_conv(tf32): no tile shape for stride %d", d->stride);
  pl->TW = best_tw; pl->TH = 128 / best_tw;
  pl->pair = 0;
  const bool shuffle = d->out_mode == ATMVFI_OUT_SHUFFLE2;
  pl->cq_pad = round_up_i(d->Cout, 32);           // ConvTranspose: each of the 4 column blocks starts on a 32-column chunk
  const int n_need = shuffle ? 4 * pl->cq_pad : round_up_i(d->Cout, 16);
  pl->n_tiles = cdiv(n_need, kMaxBlockN);
  pl->block_n = round_up_i(cdiv(n_need, pl->n_tiles), shuffle ? 32 : 16);
  {
    // thin 3x3 layers (N <= 128): two vertically stacked pixel tiles per CTA step share each weight tile and one
    // activation box with a common halo -> weight traffic and issue overhead per pixel halve
    static int pair_ok = -1;
    if (pair_ok < 0) { const char* ev = getenv("ATMVFI_TC_PAIR"); pair_ok = ev ? atoi(ev) : 1; }
    pl->pair = (pair_ok && !pl->x3 && pl->halo == 1 && pl->block_n <= 128 && Hwin > pl->TH && (2 * pl->TH + 2) * pl->TW * 128 <= 40 * 1024) ? 1 : 0;
  }
  {
    // Wave balance of small grids (the 1/8- and 1/16-resolution layers): the kernel is persistent, one CTA per SM in clusters of two,
    // so a layer with few (M tile, N tile) pairs runs ceil(tiles / 74) rounds and the last one can be mostly empty (the global motion
    // head: 96 cluster tiles = 2 rounds at 65 %).  Halving the N tile (the packed weight layout does not change: the same rows, twice
    // the tiles) doubles the tiles; it is taken when the rounds x tile-width product drops by more than the cost of the doubled
    // activation traffic.  The margin matters: taking every halving that merely ties on rounds x width made local_motion_mlp 20 % and
    // the 128-channel U-Net layers 50 % slower (half-width tiles lose pairing and re-read every activation box).
    // ATMVFI_TC_BALANCE=0 disables it.
    static int balance = -1;
    if (balance < 0) { const char* ev = getenv("ATMVFI_TC_BALANCE"); balance = ev ? atoi(ev) : 1; }
    static int sms_cached = 0;
    if (!sms_cached) {
      int dev = 0;
      cudaGetDevice(&dev);
      if (cudaDeviceGetAttribute(&sms_cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms_cached <= 0) sms_cached = 148;
    }
    if (balance && !shuffle && !pl->x3 && pl->block_n % 32 == 0 && pl->block_n >= 128 && d->out_mode != ATMVFI_OUT_QKV_HEADS) {
      const int clusters = sms_cached / 2;
      const int64_t mt_now = (int64_t)cdiv(d->Wout, pl->TW) * cdiv(Hwin, pl->TH * (pl->pair ? 2 : 1)) * d->B;
      const int64_t mt_half = (int64_t)cdiv(d->Wout, pl->TW) * cdiv(Hwin, pl->TH) * d->B;           // the halved tile never pairs
      const int64_t ct_now = (mt_now + 1) / 2 * pl->n_tiles, ct_half = (mt_half + 1) / 2 * (2 * pl->n_tiles);
      const double cost_now = (double)cdiv(ct_now, clusters) * pl->block_n * (pl->pair ? 2 : 1);
      const double cost_half = (double)cdiv(ct_half, clusters) * (pl->block_n / 2) * 1.12;
      if (ct_now <= 8 * clusters && cost_half < 0.92 * cost_now) {
        pl->block_n /= 2;
        pl->n_tiles *= 2;
        pl->pair = 0;
      }
    }
  }
  pl->tiles_x = cdiv(d->Wout, pl->TW);
  pl->tiles_y = cdiv(Hwin, pl->TH * (pl->pair ? 2 : 1));
  {
    // clusters of 2 CTAs along M share each weight tile through TMA multicast
    static int forced = -1;
    if (forced < 0) { const char* ev = getenv("ATMVFI_TC_CLUSTER"); forced = ev ? atoi(ev) : 0; }
    const int64_t m_tiles = (int64_t)pl->tiles_x * pl->tiles_y * d->B;
    pl->cluster = forced ? forced : (m_tiles >= 2 ? 2 : 1);
    ATMVFI_REQUIRE(pl->cluster == 1 || pl->cluster == 2, "gemm_conv(tf32): cluster size %d unsupported", pl->cluster);
  }

  const int n_pad = pl->n_tiles * pl->block_n;

  int ktc = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    pl->chunks[s] = cdiv(d->src[s].C, pl->chunk);
    ktc += pl->chunks[s] * pl->chunk;
  }
  ktc *= pl->ntaps;
  if (pl->x3) ktc *= 2;                           // every 32-channel chunk is stored as a hi block followed by a lo block
  ATMVFI_REQUIRE(d->ldw == ktc, "gemm_conv(tf32): packed weight row length %d != expected %d", d->ldw, ktc);

  for (int s = 0; s < d->nsrc; ++s) {
    const atmvfi_src& sr = d->src[s];
    ATMVFI_REQUIRE(((uintptr_t)sr.ptr & 15) == 0 && sr.pitch % pitch_align == 0, "gemm_conv(tensor cores): source %d must be 16-byte aligned with a pitch of whole 16-byte units", s);
    cuuint64_t gdim[4] = {(cuuint64_t)sr.C, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
    cuuint64_t gstr[3] = {(cuuint64_t)sr.pitch * es, (cuuint64_t)sr.pitch * es * d->Win, (cuuint64_t)sr.pitch * es * d->Win * d->Hin};
    cuuint32_t box[4] = {(cuuint32_t)pl->chunk, (cuuint32_t)((pl->halo == 2 ? pl->TW + 2 : pl->TW) * d->stride),
                         (cuuint32_t)((pl->halo ? (pl->pair ? 2 : 1) * pl->TH + 2 : pl->TH) * d->stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    CUresult r = enc(&pl->mapA[s], pl->f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(sr.ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ATMVFI_REQUIRE(r == CUDA_SUCCESS, "gemm_conv(tf32): cuTensorMapEncodeTiled(A%d) failed with %d (C=%d W=%d H=%d B=%d pitch=%d stride=%d)", s,
                   (int)r, sr.C, d->Win, d->Hin, d->B, sr.pitch, d->stride);
  }
  {
    ATMVFI_REQUIRE(((uintptr_t)d->weight & 15) == 0, "gemm_conv(tf32): weights must be 16-byte aligned");
    // Grouped main loop (see TcParams::grp): decided here because it fixes the shape of the weight tensor map.  The weight ring
    // must still hold two slots next to the activation ring in the smaller (16-epilogue-warp) shared-memory budget.
    static int grp_ok = -1;
    if (grp_ok < 0) { const char* ev = getenv("ATMVFI_TC_GROUP"); grp_ok = ev ? atoi(ev) : 1; }
    pl->grp = 0;
    const int sum_chunks = ktc / (pl->ntaps * pl->chunk * (pl->x3 ? 2 : 1));
    const int tile_bytes = (pl->block_n / pl->cluster) * 128;
    if (grp_ok && !pl->x3 && pl->halo == 1) {
      const int a_bytes = ((pl->pair ? 2 : 1) * pl->TH + 2) * pl->TW * 128;
      const int a_ring = 3 * ((a_bytes + 1023) / 1024 * 1024);
      if ((kDataBytes - a_ring) / (3 * tile_bytes) >= 2) pl->grp = 1;
    } else if (grp_ok >= 2 && !pl->x3 && pl->halo == 0 && sum_chunks >= 2) {
      // (measured on B200: chunk pairs do NOT pay for the 1x1 layers - fc1 92 -> 103 us, qkv 126 -> 130 us: the shallower activation
      // ring costs more than the saved barrier round trips; kept behind ATMVFI_TC_GROUP=2)
      const int a_ring = 2 * (2 * kBlockM * 128);              // two slots of two boxes (the 16-warp budget)
      if ((kDataBytes16 - a_ring) / (2 * tile_bytes) >= 2) pl->grp = 2;
    }
    const CUtensorMapDataType dt = pl->f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r;
    if (pl->grp == 1) {
      // [N_pad][K_tc] seen as {K inside one ky plane (kx, chunk, channel), n, ky}: box {chunk, rows, 3} = the vertical taps of (chunk, kx)
      const cuuint64_t S = (cuuint64_t)sum_chunks * pl->chunk;
      cuuint64_t gdim[3] = {3 * S, (cuuint64_t)n_pad, 3};
      cuuint64_t gstr[2] = {(cuuint64_t)ktc * es, 3 * S * es};
      cuuint32_t box[3] = {(cuuint32_t)pl->chunk, (cuuint32_t)(pl->block_n / pl->cluster), 3};
      cuuint32_t estr[3] = {1, 1, 1};
      r = enc(&pl->mapB, dt, 3, const_cast<float*>(d->weight), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (pl->grp == 2) {
      // {channel inside a chunk, n, chunk index over the whole K}: box {chunk, rows, 2} = the weight tiles of two consecutive chunks
      cuuint64_t gdim[3] = {(cuuint64_t)pl->chunk, (cuuint64_t)n_pad, (cuuint64_t)(ktc / pl->chunk)};
      cuuint64_t gstr[2] = {(cuuint64_t)ktc * es, (cuuint64_t)pl->chunk * es};
      cuuint32_t box[3] = {(cuuint32_t)pl->chunk, (cuuint32_t)(pl->block_n / pl->cluster), 2};
      cuuint32_t estr[3] = {1, 1, 1};
      r = enc(&pl->mapB, dt, 3, const_cast<float*>(d->weight), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gdim[2] = {(cuuint64_t)ktc, (cuuint64_t)n_pad};
      cuuint64_t gstr[1] = {(cuuint64_t)ktc * es};
      cuuint32_t box[2] = {(cuuint32_t)pl->chunk, (cuuint32_t)(pl->block_n / pl->cluster)};
      cuuint32_t estr[2] = {1, 1};
      r = enc(&pl->mapB, dt, 2, const_cast<float*>(d->weight), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    ATMVFI_REQUIRE(r == CUDA_SUCCESS, "gemm_conv(tf32): cuTensorMapEncodeTiled(B, group mode %d) failed with %d", pl->grp, (int)r);
  }
  ATMVFI_REQUIRE(((uintptr_t)d->out & 7) == 0 && d->out_pitch % 4 == 0 && (pl->f16 || ((uintptr_t)d->out & 15) == 0),
                 "gemm_conv(tensor cores): output must be 16-byte aligned (fp16 maps: 8-byte) with pitch %% 4 == 0");
  ATMVFI_REQUIRE(!d->out2 || (((uintptr_t)d->out2 & 7) == 0 && d->out2_pitch % 4 == 0 && (pl->f16 || ((uintptr_t)d->out2 & 15) == 0)),
                 "gemm_conv(tensor cores): out2 must be 16-byte aligned (fp16 maps: 8-byte)");
  {
    // TMA-store epilogue (see the kernel): layers without residual whose output is pixel-major or a k2 s2 ConvTranspose
    static int tma_store = -1;
    if (tma_store < 0) { const char* ev = getenv("ATMVFI_TC_TMASTORE"); tma_store = ev ? atoi(ev) : 1; }
    const bool shuf = d->out_mode == ATMVFI_OUT_SHUFFLE2;
    const int es_out = (pl->f16 && !d->out_f32) ? 2 : 4, es_out2 = pl->f16 ? 2 : 4;
    pl->st_ok = 0;
    if (tma_store && !pl->x3 && (d->out_mode == ATMVFI_OUT_PIXEL || shuf) && !d->residual && d->param_pad >= 32 &&
        ((uintptr_t)d->out & 15) == 0 && (d->out_pitch * es_out) % 16 == 0 &&
        (!d->out2 || (((uintptr_t)d->out2 & 15) == 0 && (d->out2_pitch * es_out2) % 16 == 0))) {
      pl->st_bx = pl->TW < 32 ? pl->TW : 32;
      pl->st_by = 32 / pl->st_bx;
      pl->st_shuffle = shuf ? 1 : 0;
      const int sc = shuf ? 2 : 1;
      auto encode_out = [&](CUtensorMap* map, void* ptr, int es, int pitch, int boxw) -> bool {
        cuuint64_t gdim[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wout * sc, (cuuint64_t)pl->row1 * sc, (cuuint64_t)d->B};
        cuuint64_t gstr[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * d->Wout * sc, (cuuint64_t)pitch * es * d->Wout * sc * d->Hout * sc};
        cuuint32_t box[4] = {(cuuint32_t)boxw, (cuuint32_t)(pl->st_bx * sc), (cuuint32_t)(pl->st_by * sc), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)sc, (cuuint32_t)sc, 1};
        const int rb = boxw * es;
        const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        return enc(map, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      };
      const bool tail = pl->block_n % 32 != 0;
      bool ok = encode_out(&pl->mapOut, d->out, es_out, d->out_pitch, 32) && (!tail || encode_out(&pl->mapOutTail, d->out, es_out, d->out_pitch, 16));
      if (ok && d->out2) ok = encode_out(&pl->mapOut2, d->out2, es_out2, d->out2_pitch, 32) && (!tail || encode_out(&pl->mapOut2Tail, d->out2, es_out2, d->out2_pitch, 16));
      pl->st_ok = ok ? 1 : 0;
    }
  }
  pl->magic = kPlanMagic;
  return 0;
}

int atmvfi_gemm_conv_tc(const atmvfi_gemm_conv_desc* d, cudaStream_t st) {
  const TcPlan* pl = reinterpret_cast<const TcPlan*>(d->tma_host);
  ATMVFI_REQUIRE(pl && pl->magic == kPlanMagic, "gemm_conv(tf32): missing plan (call atmvfi_gemm_conv_plan first)");
  typedef TcKernelFn KernelFn;
  const bool f16 = pl->f16 != 0;
  // kernel families: TcKernels<false> (tf32, this translation unit), atmvfi_tc_f16_kernel (fp16, gemm_conv_tc_f16.cu), 3xTF32 below
  auto lookup = [f16](int tbl, int a, int b, int c) -> KernelFn { return f16 ? atmvfi_tc_f16_kernel(tbl, a, b, c) : TcKernels<false>::get(tbl, a, b, c); };
  // fast epilogue: pixel-major output, no residual / second output / fp32 head copy, whole float4 columns, aligned bias and slopes
  // (channel counts that are not multiples of 4 qualify when the caller allows whole-vector stores into the map's pad lanes and passes
  // bias / slope arrays padded to a multiple of 4: pad_stores)
  const bool vec_ok = (d->Cout % 4 == 0 || d->pad_stores) && (((uintptr_t)d->bias | (uintptr_t)d->prelu | (uintptr_t)d->prelu2) & 15) == 0;
  const bool plain = !d->residual && !d->out2 && !(f16 && d->head32) && d->Cout % 4 == 0 && vec_ok;
  const int fast = (d->out_mode == ATMVFI_OUT_PIXEL && !d->residual && vec_ok) ? 1 : 0;
  const bool qkv_fast = d->out_mode == ATMVFI_OUT_QKV_HEADS && plain && !pl->halo && !pl->pair;
  ATMVFI_REQUIRE(d->out_mode != ATMVFI_OUT_QKV_HEADS || qkv_fast,
                 "gemm_conv(tf32): QKV_HEADS needs 16-byte aligned bias and Cout %% 4 == 0 (the generic epilogue does not carry this layout)");
  ATMVFI_REQUIRE(d->out_mode != ATMVFI_OUT_QKV_HEADS || !f16 || d->out_f32, "gemm_conv(f16): the head-major q|k|v output is fp32 (set out_f32)");
  // 3xTF32: [epilogue kind 0 generic / 1 fast / 2 residual prefetch][halo][cluster]
  static const KernelFn x3_table[3][2][2] = {
      {{gemm_conv_tc_kernel<0, 1, false, 0, 8, true>, gemm_conv_tc_kernel<0, 2, false, 0, 8, true>},
       {gemm_conv_tc_kernel<1, 1, false, 0, 8, true>, gemm_conv_tc_kernel<1, 2, false, 0, 8, true>}},
      {{gemm_conv_tc_kernel<0, 1, false, 1, 8, true>, gemm_conv_tc_kernel<0, 2, false, 1, 8, true>},
       {gemm_conv_tc_kernel<1, 1, false, 1, 8, true>, gemm_conv_tc_kernel<1, 2, false, 1, 8, true>}},
      {{gemm_conv_tc_kernel<0, 1, false, 2, 8, true>, gemm_conv_tc_kernel<0, 2, false, 2, 8, true>},
       {gemm_conv_tc_kernel<0, 1, false, 2, 8, true>, gemm_conv_tc_kernel<0, 2, false, 2, 8, true>}}};
  ATMVFI_REQUIRE(!pl->x3 || d->out_mode != ATMVFI_OUT_QKV_HEADS, "gemm_conv(3xtf32): the head-major q|k|v layout belongs to the tf32 attention path");
  const int res_align = f16 ? 8 : 4;             // residual rows are read as 4-element vectors: 16 B (fp32) / 8 B (fp16)
  const bool res_ok = d->residual && d->Cout % 4 == 0 && ((uintptr_t)d->residual & 15) == 0 && d->res_pitch % res_align == 0;
  KernelFn kern = lookup(0, fast, pl->pair ? 3 : pl->halo, pl->cluster - 1);
  int epi_warps = 8;
  if (pl->st_ok && pl->halo != 2) {
    // TMA-store epilogue; 16 warps for the layers whose tile time is the accumulator drain (short K, 1x1 / transposed)
    static int epi16 = -1;
    if (epi16 < 0) { const char* ev = getenv("ATMVFI_TC_EPI16"); epi16 = ev ? atoi(ev) : 1; }
    int ktc = 0;
    for (int s2 = 0; s2 < pl->nsrc; ++s2) ktc += pl->chunks[s2] * pl->chunk;
    ktc *= pl->ntaps;
    const bool can16 = pl->pair || !pl->halo;
    if (can16 && (epi16 == 2 || (epi16 == 1 && ktc <= 640 && pl->ksize == 1))) {
      kern = lookup(6, pl->pair ? 1 : 0, 0, pl->cluster - 1);
      epi_warps = 16;
    } else {
      kern = lookup(5, pl->pair ? 2 : pl->halo, 0, pl->cluster - 1);
    }
  } else if (pl->x3) {
    kern = x3_table[(res_ok && !pl->halo) ? 2 : fast][pl->halo][pl->cluster - 1];
  } else if (res_ok && !pl->halo && !pl->pair) {
    // linear layers with a residual (attention proj, Mlp fc2): generic epilogue with the residual rows prefetched.  Their tile
    // time is the epilogue (accumulator drain + residual reads + stores at memory latency), so 16 epilogue warps double the
    // loads / stores in flight per SM (ATMVFI_TC_RES16=0: 8 warps)
    static int res16 = -1;
    if (res16 < 0) { const char* ev = getenv("ATMVFI_TC_RES16"); res16 = ev ? atoi(ev) : 1; }
    if (res16) { kern = lookup(4, 0, 0, pl->cluster - 1); epi_warps = 16; }
    else kern = lookup(1, 0, 0, pl->cluster - 1);
  } else if (pl->pair || !pl->halo) {
    static int epi16 = -1;                    // ATMVFI_TC_EPI16: 0 never, 1 (default) short-K layers, 2 every eligible layer
    if (epi16 < 0) { const char* ev = getenv("ATMVFI_TC_EPI16"); epi16 = ev ? atoi(ev) : 1; }
    int ktc = 0;
    for (int s2 = 0; s2 < pl->nsrc; ++s2) ktc += pl->chunks[s2] * pl->chunk;
    ktc *= pl->ntaps;
    // 16 epilogue warps for layers whose K loop is shorter than the accumulator drain (tile time = epilogue time).
    // measured (Base 1080p): k2s2 transposed convs 539->380, 290->190, 277->209, 148->107 us; 3x3 layers lose (fewer
    // activation slots), so only 1x1 / transposed layers take this path by default
    if (epi16 == 2 || (epi16 == 1 && ktc <= 640 && pl->ksize == 1)) {
      kern = lookup(2, fast, pl->pair ? 1 : 0, pl->cluster - 1);
      epi_warps = 16;
    }
    if (qkv_fast) kern = lookup(3, epi_warps == 16 ? 1 : 0, 0, pl->cluster - 1);
  }
  // SM count and the opt-in to > 48 KB of dynamic shared memory are PER DEVICE: a process that runs models on several GPUs
  // (model.to another device, two models) configures each device the first time it launches there.
  static int sms_of_device[ATMVFI_MAX_DEVICES] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  ATMVFI_REQUIRE(dev >= 0 && dev < ATMVFI_MAX_DEVICES, "gemm_conv(tf32): device ordinal %d out of range", dev);
  if (!sms_of_device[dev]) {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = TcKernels<false>::configure();
    if (e == cudaSuccess) e = atmvfi_tc_f16_configure();
    for (int i = 0; i < 12 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(x3_table[i / 4][(i / 2) % 2][i % 2], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(8));
    if (e != cudaSuccess) {
      atmvfi_set_error("gemm_conv(tf32): cannot reserve %d B of shared memory: %s", smem_bytes(16), cudaGetErrorString(e));
      return 1;
    }
    sms_of_device[dev] = n_sm;
  }
  const int num_sms = sms_of_device[dev];
  TcParams p;
  memcpy(p.mapA, pl->mapA, sizeof(p.mapA));
  memcpy(&p.mapB, &pl->mapB, sizeof(p.mapB));
  p.nsrc = pl->nsrc;
  int ch = 0;
  for (int s = 0; s < ATMVFI_MAX_SRC; ++s) {
    p.chunks[s] = pl->chunks[s];
    ch += pl->chunks[s];
    const int rem = s < d->nsrc ? d->src[s].C % pl->chunk : 0;
    p.last_mmas[s] = rem ? (f16 ? (rem + 15) / 16 : (rem + 7) / 8) : 4;      // K per MMA: 8 (tf32) / 16 (fp16) elements = 32 bytes
  }
  p.ntaps = pl->ntaps; p.ksize = pl->ksize; p.stride = pl->stride; p.dil = pl->dil; p.pad = pl->pad;
  p.TW = pl->TW; p.TH = pl->TH; p.tiles_x = pl->tiles_x; p.tiles_y = pl->tiles_y; p.B = pl->B;
  p.block_n = pl->block_n; p.n_tiles = pl->n_tiles; p.cq_pad = pl->cq_pad;
  p.halo = pl->halo;
  p.a_bytes = (pl->halo == 2 ? (pl->TH + 2) * (pl->TW + 2) : (pl->halo ? ((pl->pair ? 2 : 1) * pl->TH + 2) * pl->TW : kBlockM)) * 128;
  p.th_super = pl->TH * (pl->pair ? 2 : 1);
  p.sum_chunks = ch;
  const int data_bytes = epi_warps == 16 ? kDataBytes16 : kDataBytes;
  p.epi_off = data_bytes;
  p.bar_off = data_bytes + epi_warps * kEpiWarpBytes;
  memcpy(&p.mapOut, &pl->mapOut, sizeof(CUtensorMap)); memcpy(&p.mapOutTail, &pl->mapOutTail, sizeof(CUtensorMap));
  memcpy(&p.mapOut2, &pl->mapOut2, sizeof(CUtensorMap)); memcpy(&p.mapOut2Tail, &pl->mapOut2Tail, sizeof(CUtensorMap));
  p.st_bx = pl->st_bx; p.st_by = pl->st_by; p.st_shuffle = pl->st_shuffle;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* ev = getenv("ATMVFI_TC_DEBUG_EPI"); dbg = ev ? atoi(ev) : 0; }
    p.dbg = dbg;
  }
  p.a_slots = pl->halo ? (epi_warps == 16 ? 2 : 3) : 4;     // 16-warp layers have short K loops: two (large, paired) boxes suffice
  if (pl->x3) p.a_slots = pl->halo ? 2 : 3;                 // every slot exists twice (raw box + its a_lo copy)
  p.grp = pl->grp;
  p.a_slot_bytes = (p.a_bytes + 1023) / 1024 * 1024;
  if (p.grp == 2) { p.a_slot_bytes = 2 * p.a_bytes; p.a_slots = epi_warps == 16 ? 2 : 3; }   // a slot = the boxes of two chunks
  p.a_lo_off = p.a_slots * p.a_slot_bytes;
  p.b_off = (pl->x3 ? 2 : 1) * p.a_slots * p.a_slot_bytes;
  p.b_slot_bytes = (pl->block_n / pl->cluster) * 128;        // 2-CTA mode: each CTA holds half of the weight tile
  if (p.grp == 1) p.b_slot_bytes *= 3;                       // a slot = the three vertical taps of (chunk, kx)
  if (p.grp == 2) p.b_slot_bytes *= 2;                       // ... the tiles of two chunks
  p.b_slots = (data_bytes - p.b_off) / p.b_slot_bytes;
  if (p.b_slots > kMaxBSlots) p.b_slots = kMaxBSlots;
  ATMVFI_REQUIRE(p.b_slots >= 2, "gemm_conv(tf32): shared memory rings too small (%d weight slots)", p.b_slots);
  p.m_tiles = pl->tiles_x * pl->tiles_y * pl->B;
  const int cs = pl->cluster;
  p.total_ctiles = ((p.m_tiles + cs - 1) / cs) * pl->n_tiles;
  p.row0 = pl->row0; p.row1 = pl->row1;
  p.epi = make_epi(d);
  if (p.total_ctiles <= 0) return 0;
  int clusters = num_sms / cs;
  if (p.total_ctiles < clusters) clusters = p.total_ctiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * cs));
  cfg.blockDim = dim3(threads_for(epi_warps));
  cfg.dynamicSmemBytes = smem_bytes(epi_warps);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = atmvfi_pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
  if (le != cudaSuccess) {
    atmvfi_set_error("gemm_conv(tf32): launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return 0;
}

#endif  // ATMVFI_TC_F16_TU
