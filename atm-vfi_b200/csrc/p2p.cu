// NVLink peer-to-peer row exchange for the spatial row-slab mode (one 4K frame pair across 2/4/8 GPUs,
// SURVEY.md section 8e).  One process per GPU; every rank allocates ONE arena with identical layout, exports
// it with CUDA IPC and maps its peers' arenas, so "the same buffer on rank p" is peer_base[p] + offset.
//
// Data path: the producer of halo rows PUSHES them into the consumer's copy of the buffer with plain 16-byte
// stores through the peer mapping (NVLink / NVSwitch), fences at system scope and raises a flag in the
// consumer's memory; the consumer's stream waits on its local flag.  One kernel per exchange site does all
// three (push my pieces, signal my destinations, wait for my sources), so a site costs one launch and the
// whole forward stays a single CUDA graph per rank.  Flags carry a monotonically increasing step number
// ("epoch") kept in device memory, so graph replays need no host-side argument patching.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

// A dead or badly skewed peer must not hang the GPU: every flag wait gives up after this long (default 4 s, set by
// atmvfi_p2p_set_timeout_ms / ATMVFI_P2P_TIMEOUT_MS; baked into a launch - and into a captured CUDA graph - when it is issued).
// A time-out sets the rank's sticky error word; from then on every exchange / step-begin kernel of that rank returns at once
// WITHOUT pushing rows or raising flags (a poisoned step must not write into peers that are still consuming the previous one),
// so its peers time out on their next wait and the failure spreads in bounded time instead of corrupting frames silently.
unsigned long long g_spin_timeout_ns = 0;
unsigned long long spin_timeout_ns() {
  if (!g_spin_timeout_ns) {
    const char* ev = getenv("ATMVFI_P2P_TIMEOUT_MS");
    const long ms = ev ? atol(ev) : 0;
    g_spin_timeout_ns = (unsigned long long)(ms > 0 ? ms : 4000) * 1000000ull;
  }
  return g_spin_timeout_ns;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until *flag has reached `epoch` (wrap-safe); on time-out record the failure and give up
__device__ __forceinline__ bool spin_until(const uint32_t* flag, uint32_t epoch, uint32_t* error_word, unsigned long long timeout_ns) {
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (globaltimer_ns() - t0 > timeout_ns) {
      atomicExch(error_word, 1u);
      return false;
    }
    __nanosleep(64);
  }
  return true;
}
__device__ __forceinline__ bool poisoned(const uint32_t* error_word) { return *reinterpret_cast<const volatile uint32_t*>(error_word) != 0; }

struct ExchangeParams {
  atmvfi_p2p_piece piece[ATMVFI_P2P_MAX_PIECES];
  int npieces;
  uint32_t* signal[ATMVFI_P2P_MAX_PEERS];        // flags in PEER memory that this rank raises
  int nsignal;
  const uint32_t* wait[ATMVFI_P2P_MAX_PEERS];    // flags in LOCAL memory raised by peers
  int nwait;
  const uint32_t* epoch;                         // local step counter
  uint32_t* counter;                             // local CTA-completion counter of this site (returns to 0)
  uint32_t* error_word;
  unsigned long long timeout_ns;
};

// grid-strided 16-byte copies; the last CTA to finish signals the destinations, then waits for the sources
__global__ void __launch_bounds__(256) p2p_exchange_kernel(const __grid_constant__ ExchangeParams p) {
  if (poisoned(p.error_word)) return;            // set by an earlier kernel of this stream only: uniform over the grid
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int i = 0; i < p.npieces; ++i) {
    const atmvfi_p2p_piece& pc = p.piece[i];
    const bool a16 = ((((uintptr_t)pc.src | (uintptr_t)pc.dst | pc.chunk_bytes | pc.chunk_stride) & 15) == 0);   // uniform
    if (a16) {
      const int64_t vec_per_chunk = (int64_t)(pc.chunk_bytes >> 4);
      const int64_t total = vec_per_chunk * pc.nchunks;
      for (int64_t v = tid; v < total; v += nthreads) {
        const int64_t ch = v / vec_per_chunk, off = v - ch * vec_per_chunk;
        const int4 val = *(reinterpret_cast<const int4*>(static_cast<const char*>(pc.src) + ch * pc.chunk_stride) + off);
        *(reinterpret_cast<int4*>(static_cast<char*>(pc.dst) + ch * pc.chunk_stride) + off) = val;
      }
    } else {   // small planar maps whose rows are not multiples of 16 bytes (e.g. 1/16-resolution flows)
      const int64_t w_per_chunk = (int64_t)(pc.chunk_bytes >> 2);
      const int64_t total = w_per_chunk * pc.nchunks;
      for (int64_t v = tid; v < total; v += nthreads) {
        const int64_t ch = v / w_per_chunk, off = v - ch * w_per_chunk;
        const uint32_t val = *(reinterpret_cast<const uint32_t*>(static_cast<const char*>(pc.src) + ch * pc.chunk_stride) + off);
        *(reinterpret_cast<uint32_t*>(static_cast<char*>(pc.dst) + ch * pc.chunk_stride) + off) = val;
      }
    }
  }
  __threadfence_system();                        // my stores are visible system-wide before anything I do next
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(p.counter, 1u);
    if (done == gridDim.x - 1) {                 // every CTA's stores are fenced and counted
      __threadfence_system();
      const uint32_t e = *p.epoch;
      for (int i = 0; i < p.nsignal; ++i) st_release_sys(p.signal[i], e);
      *p.counter = 0;
      for (int i = 0; i < p.nwait; ++i)
        if (!spin_until(p.wait[i], e, p.error_word, p.timeout_ns)) break;
    }
  }
}

// Start of a step: bump the epoch, tell every peer, wait until every peer has started the same step (which means it
// has finished the previous one: its halo rows may be overwritten).
__global__ void p2p_step_begin_kernel(uint32_t* epoch, ExchangeParams p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (poisoned(p.error_word)) return;
    const uint32_t e = *epoch + 1;
    *epoch = e;
    __threadfence_system();
    for (int i = 0; i < p.nsignal; ++i) st_release_sys(p.signal[i], e);
    for (int i = 0; i < p.nwait; ++i)
      if (!spin_until(p.wait[i], e, p.error_word, p.timeout_ns)) break;
  }
}

}  // namespace

extern "C" {

int atmvfi_arena_alloc(size_t bytes, void** ptr) {
  ATMVFI_REQUIRE(ptr != nullptr && bytes > 0, "arena_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) {
    atmvfi_set_error("arena_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int atmvfi_arena_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    atmvfi_set_error("arena_free: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int atmvfi_ipc_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == ATMVFI_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    atmvfi_set_error("ipc_export: %s", cudaGetErrorString(e));
    return 1;
  }
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

int atmvfi_ipc_open(const unsigned char* handle64, void** peer_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    atmvfi_set_error("ipc_open: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int atmvfi_ipc_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) {
    atmvfi_set_error("ipc_close: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int atmvfi_p2p_set_timeout_ms(int ms) {
  ATMVFI_REQUIRE(ms > 0, "p2p_set_timeout_ms: %d is not a positive number of milliseconds", ms);
  g_spin_timeout_ns = (unsigned long long)ms * 1000000ull;
  return 0;
}

int atmvfi_p2p_exchange(const atmvfi_p2p_piece* pieces, int npieces, uint32_t* const* signal_flags, int nsignal,
                        const uint32_t* const* wait_flags, int nwait, const uint32_t* epoch, uint32_t* counter,
                        uint32_t* error_word, void* stream) {
  ATMVFI_REQUIRE(npieces >= 0 && npieces <= ATMVFI_P2P_MAX_PIECES, "p2p_exchange: %d pieces (max %d)", npieces, ATMVFI_P2P_MAX_PIECES);
  ATMVFI_REQUIRE(nsignal >= 0 && nsignal <= ATMVFI_P2P_MAX_PEERS && nwait >= 0 && nwait <= ATMVFI_P2P_MAX_PEERS, "p2p_exchange: too many peers");
  ATMVFI_REQUIRE(epoch && counter && error_word, "p2p_exchange: null control pointer");
  ExchangeParams p;
  memset(&p, 0, sizeof(p));
  int64_t bytes = 0;
  for (int i = 0; i < npieces; ++i) {
    const atmvfi_p2p_piece& pc = pieces[i];
    ATMVFI_REQUIRE((((uintptr_t)pc.src | (uintptr_t)pc.dst | pc.chunk_bytes | pc.chunk_stride) & 3) == 0,
                   "p2p_exchange: piece %d is not 4-byte aligned", i);
    p.piece[i] = pc;
    bytes += (int64_t)pc.chunk_bytes * pc.nchunks;
  }
  p.npieces = npieces;
  for (int i = 0; i < nsignal; ++i) p.signal[i] = signal_flags[i];
  for (int i = 0; i < nwait; ++i) p.wait[i] = wait_flags[i];
  p.nsignal = nsignal; p.nwait = nwait; p.epoch = epoch; p.counter = counter; p.error_word = error_word;
  p.timeout_ns = spin_timeout_ns();
  // enough CTAs to keep the NVLink ports busy for large pushes, one CTA for flag-only sites
  int64_t ctas = (bytes + 65535) / 65536;
  if (ctas < 1) ctas = 1;
  if (ctas > 148) ctas = 148;
  p2p_exchange_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(p);
  ATMVFI_CHECK_LAUNCH("p2p_exchange");
  return 0;
}

int atmvfi_p2p_step_begin(uint32_t* epoch, uint32_t* const* signal_flags, int nsignal, const uint32_t* const* wait_flags, int nwait,
                          uint32_t* error_word, void* stream) {
  ATMVFI_REQUIRE(nsignal >= 0 && nsignal <= ATMVFI_P2P_MAX_PEERS && nwait >= 0 && nwait <= ATMVFI_P2P_MAX_PEERS, "p2p_step_begin: too many peers");
  ExchangeParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < nsignal; ++i) p.signal[i] = signal_flags[i];
  for (int i = 0; i < nwait; ++i) p.wait[i] = wait_flags[i];
  p.nsignal = nsignal; p.nwait = nwait; p.error_word = error_word;
  p.timeout_ns = spin_timeout_ns();
  p2p_step_begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(epoch, p);
  ATMVFI_CHECK_LAUNCH("p2p_step_begin");
  return 0;
}

}  // extern "C"
