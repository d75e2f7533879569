// One-shot NVLink/NVSwitch peer-to-peer all-gather for the contrastive head.
// Replaces diffdist.functional.all_gather + torch.distributed.barrier() in the reference
// (modules/util_module.py:180-190, modules/modeling.py:352-354): every rank writes its slab of L2-normalised
// embeddings (and later its per-row log-sum-exp) straight into all peers' buffers over NVLink and raises a
// per-peer epoch flag; nobody waits for a collective launch, a ring or a host barrier.
//
// Buffers are plain cudaMalloc allocations exported with cudaIpcGetMemHandle (set up once, off the hot path).
// Protocol per exchange e (monotone epoch counter, one signal pad per rank):
//   wait   consumed[self][p] >= e-1   for all p  (peer p finished reading what I wrote last time)
//   copy   my slab -> buffer of every peer (and my own) at slot `rank`
//   signal ready[p][self] = e         for all p  (st.release.sys after __threadfence_system)
//   wait   ready[self][p] >= e        for all p  (ld.acquire.sys)
// and, after the last reader of the gathered data, sc_p2p_release marks consumed[p][self] = e on all peers.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

extern void sc_count_launch(int n);

namespace {

constexpr int MAX_WORLD = 16;

unsigned long long p2p_timeout_ns() {
  static const unsigned long long v = [] {
    const char* e = getenv("SEGCLIP_P2P_TIMEOUT_S");
    const double s = e ? atof(e) : 1800.0;
    return (unsigned long long)((s > 0 ? s : 1800.0) * 1e9);
  }();
  return v;
}

struct PeerTable {
  void* buf[MAX_WORLD];        // peer data buffers (index = rank), same layout everywhere
  uint32_t* pad[MAX_WORLD];    // peer signal pads: [0..W) ready, [W..2W) consumed
};

SC_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SC_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
SC_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Waits for a peer's flag with exponential __nanosleep back-off.  Ranks legitimately drift apart by minutes (rank-0-only
// checkpoint + evaluation between epochs, main_task_align.py:484-490; data-loader start-up), so the dead-peer check is a
// WALL-CLOCK limit (SEGCLIP_P2P_TIMEOUT_S, default 1800 s -- the same order as the NCCL watchdog), not a poll count.
SC_DEVINL void spin_until(const uint32_t* p, uint32_t epoch, unsigned long long timeout_ns) {
  if ((int)(ld_acquire_sys(p) - epoch) >= 0) return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned ns = 32;
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(ns);
    if (ns < 4096) ns <<= 1;
    if (globaltimer_ns() - t0 > timeout_ns) {   // a peer died or the protocol is broken: fail loudly
      printf("segclip_b200 p2p: peer flag timeout after %llu s (want epoch %u, have %u)\n", timeout_ns / 1000000000ull, epoch,
             ld_acquire_sys(p));
      __trap();
    }
  }
}

struct Segments {
  const void* src[4];
  long off[4];     // byte offset inside every peer buffer
  long bytes[4];
  int n;
};

// Writes each segment of this rank into the same offset of every peer buffer (own buffer included).
__global__ void __launch_bounds__(256) p2p_allgather_kernel(PeerTable tab, Segments seg, int rank, int world, uint32_t epoch,
                                                             unsigned int* __restrict__ block_counter,
                                                             unsigned long long timeout_ns) {
  __shared__ bool last;
  uint32_t* mypad = tab.pad[rank];
  // block 0 waits until every peer has released what this rank wrote last time, then opens the gate for the others
  if (blockIdx.x == 0) {
    if (threadIdx.x < world) spin_until(mypad + world + threadIdx.x, epoch - 1, timeout_ns);
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(block_counter + 1, epoch);
  } else if (threadIdx.x == 0) {
    while ((int)(atomicAdd(block_counter + 1, 0u) - epoch) < 0) {}
  }
  __syncthreads();
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
  for (int p = 0; p < world; ++p) {
    char* base = (char*)tab.buf[(rank + p) % world];      // stagger the peers
    for (int s = 0; s < seg.n; ++s) {
      if (((seg.bytes[s] | seg.off[s] | (long)(uintptr_t)seg.src[s]) & 15) == 0) {
        const uint4* src = (const uint4*)seg.src[s];
        uint4* dst = (uint4*)(base + seg.off[s]);
        for (long i = tid; i < seg.bytes[s] / 16; i += nthr) dst[i] = src[i];
      } else {
        const uint32_t* src = (const uint32_t*)seg.src[s];
        uint32_t* dst = (uint32_t*)(base + seg.off[s]);
        for (long i = tid; i < seg.bytes[s] / 4; i += nthr) dst[i] = src[i];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(block_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *block_counter = 0;
  __threadfence_system();
  if (threadIdx.x < world) {
    st_release_sys(tab.pad[threadIdx.x] + rank, epoch);      // ready[p][rank] = epoch
    spin_until(mypad + threadIdx.x, epoch, timeout_ns);               // ready[self][p] >= epoch
  }
}

__global__ void p2p_release_kernel(PeerTable tab, int rank, int world, uint32_t epoch) {
  if (threadIdx.x < world) st_release_sys(tab.pad[threadIdx.x] + world + rank, epoch);   // consumed[p][rank] = epoch
}

}  // namespace

extern "C" {

// Allocates `bytes` of device memory + a zeroed signal pad and returns their IPC handles (64 bytes each).
int sc_p2p_alloc(int64_t bytes, void** buf, void** pad, void* buf_handle_out, void* pad_handle_out) {
  SC_CHECK_ARG(bytes > 0 && buf && pad && buf_handle_out && pad_handle_out, "sc_p2p_alloc: bad args");
  SC_CUDA(cudaMalloc(buf, (size_t)bytes));
  SC_CUDA(cudaMalloc(pad, 4096));
  SC_CUDA(cudaMemset(*pad, 0, 4096));
  SC_CUDA(cudaMemset(*buf, 0, (size_t)bytes));
  SC_CUDA(cudaDeviceSynchronize());
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  SC_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)buf_handle_out, *buf));
  SC_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)pad_handle_out, *pad));
  return SC_OK;
}

int sc_p2p_open(const void* handle, void** ptr) {
  SC_CHECK_ARG(handle && ptr, "sc_p2p_open: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  SC_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SC_OK;
}

int sc_p2p_close(void* ptr) {
  SC_CUDA(cudaIpcCloseMemHandle(ptr));
  return SC_OK;
}

int sc_p2p_free(void* buf, void* pad) {
  if (buf) SC_CUDA(cudaFree(buf));
  if (pad) SC_CUDA(cudaFree(pad));
  return SC_OK;
}

// peer_bufs / peer_pads: host arrays of `world` device pointers (index = rank; own entries included).
// Copies nseg (<= 4) segments srcs[i][0, nbytes[i]) to byte offset offs[i] of every peer buffer, then synchronises on
// the epoch flags.  Sizes/offsets must be multiples of 4 bytes.  scratch: device uint32[2], zero-initialised once.
int sc_p2p_allgather(int nseg, const void* const* srcs, const int64_t* nbytes, const int64_t* offs, void* const* peer_bufs,
                     void* const* peer_pads, int rank, int world, uint32_t epoch, void* scratch, void* stream) {
  SC_CHECK_ARG(srcs && nbytes && offs && peer_bufs && peer_pads && scratch, "sc_p2p_allgather: null pointer");
  SC_CHECK_ARG(nseg >= 1 && nseg <= 4, "sc_p2p_allgather: 1..4 segments");
  SC_CHECK_ARG(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "sc_p2p_allgather: bad rank/world");
  PeerTable tab;
  for (int i = 0; i < world; ++i) {
    tab.buf[i] = peer_bufs[i];
    tab.pad[i] = (uint32_t*)peer_pads[i];
  }
  Segments seg;
  seg.n = nseg;
  long total = 0;
  for (int i = 0; i < nseg; ++i) {
    SC_CHECK_ARG(nbytes[i] % 4 == 0 && offs[i] % 4 == 0 && ((uintptr_t)srcs[i] & 3) == 0, "sc_p2p_allgather: 4-byte alignment");
    seg.src[i] = srcs[i];
    seg.bytes[i] = nbytes[i];
    seg.off[i] = offs[i];
    total += nbytes[i];
  }
  int blocks = (int)((total / 16 + 255) / 256);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  sc_count_launch(1);
  p2p_allgather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tab, seg, rank, world, epoch, (unsigned int*)scratch,
                                                                       p2p_timeout_ns());
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_p2p_release(void* const* peer_pads, int rank, int world, uint32_t epoch, void* stream) {
  SC_CHECK_ARG(peer_pads && world >= 1 && world <= MAX_WORLD, "sc_p2p_release: bad args");
  PeerTable tab;
  for (int i = 0; i < world; ++i) tab.pad[i] = (uint32_t*)peer_pads[i];
  sc_count_launch(1);
  p2p_release_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(tab, rank, world, epoch);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

}  // extern "C"
