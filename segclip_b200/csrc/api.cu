// Library-level plumbing: error text, launch counter, sc_gemm dispatch.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void sc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_kind[SC_K_COUNT];

void sc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void sc_count_kernel(int kind, int n) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  if (kind >= 0 && kind < SC_K_COUNT) g_kind[kind].fetch_add(1, std::memory_order_relaxed);
}

int sc_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool sc_pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("SC_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Entry points may be called from a thread that has not touched the CUDA runtime yet (autograd's
// backward worker).  Driver-API calls (cuTensorMapEncodeTiled) need the primary context bound to the
// calling thread, and the thread's current device must be the stream's.
int sc_enter(cudaStream_t st) {
  static thread_local bool bound = false;
  int cur = -1;
  SC_CUDA(cudaGetDevice(&cur));
  if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread) {
    int dev = cur;
    if (cudaStreamGetDevice(st, &dev) == cudaSuccess && dev != cur) {
      SC_CUDA(cudaSetDevice(dev));
      bound = false;
    }
  }
  if (!bound) {
    SC_CUDA(cudaFree(0));
    bound = true;
  }
  return SC_OK;
}

int sc_gemm_tc(const sc_gemm_desc* d, cudaStream_t st);
int sc_gemm_simt(const sc_gemm_desc* d, cudaStream_t st);

extern "C" {

const char* sc_last_error(void) { return g_err; }
int sc_abi_version(void) { return SC_ABI_VERSION; }
long long sc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
long long sc_kernel_launches(int kind) {
  return (kind >= 0 && kind < SC_K_COUNT) ? g_kind[kind].load(std::memory_order_relaxed) : -1;
}

int sc_gemm(const sc_gemm_desc* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(d && d->A && d->B && d->C, "sc_gemm: null operand");
  SC_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "sc_gemm: bad shape %d %d %d", d->M, d->N, d->K);
  SC_CHECK_ARG(!(d->accumulate && d->c_dtype != SC_F32), "sc_gemm: accumulate needs fp32 C");
  {
    int rc = sc_enter(st);
    if (rc) return rc;
  }
  if (d->in_dtype == SC_BF16 && !d->force_simt) return sc_gemm_tc(d, st);
  SC_CHECK_ARG(!d->colsum_out, "sc_gemm: colsum_out is only available on the bf16 tensor-core path");
  SC_CHECK_ARG(d->in_dtype == SC_F32 || d->in_dtype == SC_BF16, "sc_gemm: bad in_dtype %d", d->in_dtype);
  return sc_gemm_simt(d, st);
}

}  // extern "C"
