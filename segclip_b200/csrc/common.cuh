// Common device/host helpers for libsegclip_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/segclip_b200.h"

#define SC_DEVINL __device__ __forceinline__

// ---- error plumbing ---------------------------------------------------------------------------
void sc_set_error(const char* fmt, ...);

#define SC_CHECK_ARG(cond, ...)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      sc_set_error(__VA_ARGS__);                        \
      return SC_ERR_INVALID;                            \
    }                                                   \
  } while (0)

#define SC_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      sc_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SC_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define SC_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      sc_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SC_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

int sc_num_sms();   // SM count of the CURRENT device (cached per device)

// ---- programmatic dependent launch ------------------------------------------------------------------
// A step is ~600 back-to-back launches on one stream, most of them persistent kernels whose CTAs all retire within a few
// microseconds of each other.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization a kernel's CTAs are
// scheduled onto an SM as soon as the previous kernel's CTAs have left it: launch latency and the prologue (mbarrier
// init, TMEM allocation, tensor-map prefetch) run under the previous kernel's tail.  Every kernel launched this way
// executes pdl_launch_dependents() first (lets ITS successor in) and pdl_wait() -- which returns once the previous
// grid has completed and its writes are visible -- before its first global-memory access.  SC_PDL=0 = plain launches.
bool sc_pdl_enabled();
SC_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SC_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t sc_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = sc_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Once-per-(call site, device) latch for cudaFuncSetAttribute: the attribute is per device and entry points are called
// from several host threads (autograd runs backward on its own thread), so the latch is an atomic per-device bit mask.
// Two threads racing on the first call both set the (idempotent) attribute.
#include <atomic>
struct sc_device_once {
  std::atomic<unsigned long long> mask{0};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask.load(std::memory_order_acquire) & bit) return false;
    return true;
  }
  void done() {
    int dev = 0;
    cudaGetDevice(&dev);
    mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
};

// ---- dtype helpers ------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> struct sc_dtype_of;
template <> struct sc_dtype_of<float> { static constexpr int value = SC_F32; };
template <> struct sc_dtype_of<bf16> { static constexpr int value = SC_BF16; };

SC_DEVINL float to_f32(float v) { return v; }
SC_DEVINL float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> SC_DEVINL T from_f32(float v);
template <> SC_DEVINL float from_f32<float>(float v) { return v; }
template <> SC_DEVINL bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// load/store with runtime dtype (used by low-volume kernels only)
SC_DEVINL float ld_any(const void* p, long i, int dtype) {
  return dtype == SC_F32 ? ((const float*)p)[i] : __bfloat162float(((const bf16*)p)[i]);
}
SC_DEVINL void st_any(void* p, long i, int dtype, float v) {
  if (dtype == SC_F32) ((float*)p)[i] = v;
  else ((bf16*)p)[i] = __float2bfloat16_rn(v);
}

// ---- activations ----------------------------------------------------------------------------------
// QuickGELU x*sigmoid(1.702x) (reference modules/module_clip_util.py:134-136),
// exact erf GELU (nn.GELU default; reference modules/module_seg_vit.py:128, module_mae.py:151)
SC_DEVINL float act_fwd(float x, int act) {
  if (act == SC_ACT_QUICKGELU) return __fdividef(x, 1.0f + __expf(-1.702f * x));
  if (act == SC_ACT_GELU_ERF) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}
SC_DEVINL float act_grad(float x, int act) {
  if (act == SC_ACT_DERIV) return x;          // x already is the derivative (stored by the forward epilogue)
  if (act == SC_ACT_QUICKGELU) {
    float s = __fdividef(1.0f, 1.0f + __expf(-1.702f * x));
    return s * (1.0f + 1.702f * x * (1.0f - s));
  }
  if (act == SC_ACT_GELU_ERF) {
    float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
  }
  return 1.0f;
}

// ---- warp helpers -----------------------------------------------------------------------------------
SC_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
SC_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
