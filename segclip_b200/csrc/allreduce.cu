// Gradient all-reduce over NVSwitch multicast (NVLS), SURVEY 8(f) rank 2: replaces the NCCL ring all-reduce of the DDP
// buckets (main_task_align.py:251-252) by ONE two-shot kernel per bucket:
//
//   every rank owns 1/W of the bucket; for its slice it issues  multimem.ld_reduce.add.v4.f32  on the MULTICAST address
//   (the switch reads the W replicas, adds them in fp32 and returns the sum), scales by 1/W (DDP's mean) and writes the
//   result back with  multimem.st  (the switch broadcasts it into all W replicas).  Each byte crosses each GPU's links
//   once in each direction; there is no ring, no staging buffer, no reduction arithmetic on the SMs.
//
// The flat gradient buffer is symmetric memory (same offset on every rank, mapped once by the host side:
// segclip_b200/allreduce.py).  The kernel uses no shared memory and few registers, so its CTAs co-reside with the
// persistent tcgen05 GEMM CTAs (229 KB of smem, one per SM) -- NCCL's kernels cannot, which left the ring all-reduce
// interleaving between GEMM launches.  Cross-rank ordering: per-CTA flag barriers (st.release.sys / ld.acquire.sys on
// peer-mapped flag arrays, monotonically increasing epochs -> no resets) before the first load (all replicas of the
// bucket are final) and after the last store (all replicas hold the mean).
#include "common.cuh"

extern void sc_count_launch(int n);

namespace {

SC_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SC_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
SC_DEVINL unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// CTA `b` of every rank meets: thread t < world raises flag (b, rank) in peer t's array and waits for flag (b, t) in its own.
SC_DEVINL void rank_barrier(uint32_t* const* flags, int rank, int world, uint32_t epoch, unsigned long long timeout_ns, int* err) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    st_release_sys(flags[t] + (size_t)blockIdx.x * world + rank, epoch);
    const uint32_t* mine = flags[rank] + (size_t)blockIdx.x * world + t;
    const unsigned long long t0 = gtimer();
    // epochs only grow; the signed difference tolerates wrap-around
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      __nanosleep(64);
      if (gtimer() - t0 > timeout_ns) {        // a peer never arrived: report instead of hanging the GPU
        if (err) *err = 1;
        break;
      }
    }
  }
  __syncthreads();
}

constexpr int NVLS_THREADS = 256;      // 256 threads x ~64 registers fit beside a resident GEMM CTA (320 threads x <= 136 registers)
constexpr int NVLS_UNROLL = 8;         // 16-byte reductions in flight per thread: the loop is bound by NVLink round trips, not issue

__global__ void __launch_bounds__(NVLS_THREADS) nvls_allreduce_kernel(float* mc, size_t n4, int rank, int world, uint32_t* const* flags,
                                                            uint32_t epoch, float scale, unsigned long long timeout_ns, int* err) {
  rank_barrier(flags, rank, world, epoch, timeout_ns, err);
  const size_t per = (n4 + world - 1) / world;
  const size_t lo = (size_t)rank * per < n4 ? (size_t)rank * per : n4, hi = lo + per < n4 ? lo + per : n4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (NVLS_UNROLL - 1) * stride < hi; i += NVLS_UNROLL * stride) {
    float4 v[NVLS_UNROLL];
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                   : "l"(mc + 4 * (i + u * stride))
                   : "memory");
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * (i + u * stride)),
                   "f"(v[u].x * scale), "f"(v[u].y * scale), "f"(v[u].z * scale), "f"(v[u].w * scale)
                   : "memory");
  }
  for (; i < hi; i += stride) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc + 4 * i)
                 : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v.x * scale),
                 "f"(v.y * scale), "f"(v.z * scale), "f"(v.w * scale)
                 : "memory");
  }
  rank_barrier(flags, rank, world, epoch + 1, timeout_ns, err);
}

}  // namespace

extern "C" int sc_nvls_allreduce(void* multicast_ptr, int64_t count, float scale, int rank, int world, void* const* peer_flags_dev,
                                 uint32_t epoch, int blocks, int64_t timeout_ms, int32_t* err_flag, void* stream) {
  SC_CHECK_ARG(multicast_ptr && peer_flags_dev && count > 0 && world >= 1 && rank >= 0 && rank < world,
               "sc_nvls_allreduce: bad arguments");
  SC_CHECK_ARG(count % 4 == 0 && ((uintptr_t)multicast_ptr & 15) == 0, "sc_nvls_allreduce: the range must be 16-byte aligned");
  SC_CHECK_ARG(blocks >= 1 && blocks <= SC_NVLS_MAX_BLOCKS && world <= NVLS_THREADS, "sc_nvls_allreduce: 1..%d blocks", SC_NVLS_MAX_BLOCKS);
  sc_count_launch(1);
  nvls_allreduce_kernel<<<blocks, NVLS_THREADS, 0, (cudaStream_t)stream>>>((float*)multicast_ptr, (size_t)(count / 4), rank, world,
                                                                 (uint32_t* const*)peer_flags_dev, epoch, scale,
                                                                 (unsigned long long)timeout_ms * 1000000ull, err_flag);
  SC_LAUNCH_CHECK();
  return SC_OK;
}
