// tcgen05 GEMM for sm_100a: C = epilogue(A * B^T), bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-7  epilogue   (tcgen05.ld TMEM -> registers -> fused epilogue -> global)
//   warp  8    TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier expect-tx ring)
//   warp  9    MMA issuer (single thread, tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
// Tile 128 x BN x 64 (BN = 256 or 128), 4/6-stage smem ring, two TMEM accumulators so the
// epilogue of tile i overlaps the main loop of tile i+1.  Operands may be K-major (row-major
// [rows, K]) or MN-major (row-major [K, rows]); the latter serves dgrad (B = W) and wgrad
// (A = dY, B = X) without materialising transposes.  Optional split-K accumulates with fp32
// atomics (wgrad: K = batch*tokens, output only a few dozen tiles).
#include <cuda.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "gemm_tc_common.cuh"

// ---- host helpers shared with gemm_tc2.cu ----
// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + dispatch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, stride;
  uint32_t b0, b1;
  int swizzle;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 && swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ k.d0;
    h = h * 1000003u ^ k.d1;
    h = h * 1000003u ^ k.stride;
    h = h * 1000003u ^ ((uint64_t)k.b0 << 32 | k.b1);
    h = h * 1000003u ^ (uint64_t)k.swizzle;
    return h;
  }
};

// bf16 2-D tensor map, 128B swizzle, zero fill.  dim0 is the contiguous dimension.
int sc_get_tensor_map(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                      CUtensorMap* out) {
  return sc_get_tensor_map_sw(ptr, dim0, dim1, stride_elems, box0, box1, 128, out);
}

// same with an explicit swizzle span in bytes (128 or 64; 64 is used by the TMA-store epilogue's 32 x 32 bf16 boxes)
int sc_get_tensor_map_sw(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                         int swizzle_bytes, CUtensorMap* out) {
  return sc_get_tensor_map_any(ptr, dim0, dim1, stride_elems, box0, box1, swizzle_bytes, 2, out);
}

// element size 2 = bf16, 4 = fp32 (fp32 residual / output boxes of the GEMM epilogue)
int sc_get_tensor_map_any(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                          int swizzle_bytes, int elem_bytes, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, dim0, dim1, stride_elems, box0, box1, swizzle_bytes | (elem_bytes << 16)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return SC_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    sc_set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return SC_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstride[1] = {stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sc_set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p dims=(%llu,%llu) stride=%llu box=(%u,%u)", (int)r, ptr,
                 (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)stride_elems, box0, box1);
    return SC_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 65536) cache.clear();
  cache[key] = *out;
  return SC_OK;
}

// bf16 3-D tensor map {dim0 (contiguous), dim1, dim2} with element strides for dims 1 and 2, box {box0, box1, 1}: the
// attention backward stores dQ / dK / dV tiles through it so that rows past a sample's L are clipped by the TMA itself.
int sc_get_tensor_map_3d(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1_elems,
                         uint64_t stride2_elems, uint32_t box0, uint32_t box1, int swizzle_bytes, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, dim0, dim1 | (dim2 << 32), stride1_elems | (stride2_elems << 32), box0, box1, swizzle_bytes | (3 << 24)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return SC_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    sc_set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return SC_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {dim0, dim1, dim2};
  cuuint64_t gstride[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sc_set_error("cuTensorMapEncodeTiled(3d) failed (%d): ptr=%p dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)", (int)r, ptr,
                 (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2, (unsigned long long)stride1_elems,
                 (unsigned long long)stride2_elems, box0, box1);
    return SC_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 65536) cache.clear();
  cache[key] = *out;
  return SC_OK;
}

// bf16 tensor map of rank 3 or 4 with explicit dims / element strides (of dims 1..) / box: head-padded attention tiles -- the
// head dimension is its own (innermost) tensor dimension, so a 64-wide box over a 48-wide head is zero-filled on loads and
// clipped on stores.  Not cached by pointer alone: the key folds every field.
int sc_get_tensor_map_nd(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* boxd,
                         int swizzle_bytes, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  uint64_t hd = 1469598103934665603ull, hs = 1469598103934665603ull;
  for (int i = 0; i < rank; ++i) hd = (hd ^ dims[i]) * 1099511628211ull;
  for (int i = 0; i + 1 < rank; ++i) hs = (hs ^ strides_elems[i]) * 1099511628211ull;
  uint32_t hb0 = 2166136261u, hb1 = (uint32_t)rank;
  for (int i = 0; i < rank; ++i) hb0 = (hb0 ^ boxd[i]) * 16777619u;
  MapKey key{ptr, hd, dims[0] | (dims[rank - 1] << 32), hs, hb0, hb1, swizzle_bytes | (5 << 24)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return SC_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc || rank < 3 || rank > 4) {
    sc_set_error("cuTensorMapEncodeTiled not available from the CUDA driver (or bad rank %d)", rank);
    return SC_ERR_CUDA;
  }
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; box[i] = boxd[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_elems[i] * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sc_set_error("cuTensorMapEncodeTiled(rank %d) failed (%d): ptr=%p dims=(%llu,%llu,%llu) box=(%u,%u,%u)", rank, (int)r, ptr,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], boxd[0], boxd[1], boxd[2]);
    return SC_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 65536) cache.clear();
  cache[key] = *out;
  return SC_OK;
}

// Picks the compile-time specialised epilogue (tc::EF_*) for a descriptor, or EF_GENERIC.
int sc_select_epilogue(const sc_gemm_desc* d, int splits) {
  using namespace tc;
  int ef = EF_GENERIC;
  const bool simple = d->alpha == 1.0f && !d->rowbias;
  const bool acc = d->accumulate && splits <= 1;
  // colsum_out is implemented by the bf16-output epilogues only (checked by the caller below)
  if (simple) {
    const bool bias = d->bias != nullptr, resid = d->residual != nullptr, c2 = d->C2 != nullptr, aux = d->mul_aux != nullptr;
    const bool out_bf16 = d->c_dtype == SC_BF16;
    if (splits > 1) {
      if (!bias && !resid && !c2 && !aux && d->act == SC_ACT_NONE) ef = EF_ATOMIC | EF_OUT_F32;
    } else if (!aux && !c2 && !resid && !bias && d->act == SC_ACT_NONE && !out_bf16) {
      ef = EF_OUT_F32 | (acc ? EF_ACCUM : 0);      // plain fp32 output (grouped-conv features, accumulated dgrads)
    } else if (acc) {
      // every other accumulating combination: generic epilogue
    } else if (!aux && !c2 && !resid && d->act == SC_ACT_NONE && out_bf16) {
      ef = bias ? EF_BIAS : 0;
    } else if (bias && c2 && d->c2_dtype == SC_BF16 && out_bf16 && d->act == SC_ACT_QUICKGELU && !resid && !aux) {
      ef = EF_BIAS | EF_QGELU | EF_C2 | (d->c2_is_act_grad ? EF_C2_DERIV : 0);
    } else if (bias && resid && d->residual_dtype == SC_F32 && !c2 && !aux && d->act == SC_ACT_NONE && !out_bf16) {
      ef = EF_BIAS | EF_RESID | EF_OUT_F32;
    } else if (bias && resid && d->residual_dtype == SC_BF16 && !c2 && !aux && d->act == SC_ACT_NONE && out_bf16) {
      ef = EF_BIAS | EF_RESID_BF;          // bf16 residual stream (2-CTA kernel; the 1-CTA kernel takes the generic epilogue)
    } else if (aux && d->mul_aux_dtype == SC_BF16 && d->mul_aux_act == SC_ACT_QUICKGELU && !bias && !resid && !c2 &&
               d->act == SC_ACT_NONE && out_bf16) {
      ef = EF_MULAUX_QGELU;
    } else if (aux && d->mul_aux_dtype == SC_BF16 && d->mul_aux_act == SC_ACT_DERIV && !bias && !resid && !c2 &&
               d->act == SC_ACT_NONE && out_bf16) {
      ef = EF_MULAUX_DERIV;
    } else if (bias && c2 && d->c2_dtype == SC_BF16 && out_bf16 && d->act == SC_ACT_GELU_ERF && !resid && !aux) {
      ef = EF_BIAS | EF_GELU | EF_C2 | (d->c2_is_act_grad ? EF_C2_DERIV : 0);
    } else if (aux && d->mul_aux_dtype == SC_BF16 && d->mul_aux_act == SC_ACT_GELU_ERF && !bias && !resid && !c2 &&
               d->act == SC_ACT_NONE && out_bf16) {
      ef = EF_MULAUX_GELU;
    }
  }
  return ef;
}

namespace {
using namespace tc;

template <int BN>
struct TileCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + NUM_EPI_WARPS * 4096;
};

template <int BN, bool A_MN, bool B_MN, int EF>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int tiles_m,
               int tiles_n, int splits, int kb_total, int kb_per_split, EpiParams ep) {
  using Cfg = TileCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // 128B-swizzle atoms need a 1024-byte aligned base
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  uint8_t* epi_stage = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256;   // 8 warps x 4 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == NUM_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();      // after the TMEM allocation: a successor CTA sharing this SM can never starve this one
  pdl_wait();                   // prologue above overlaps the previous kernel's tail (common.cuh)

  const int num_items = tiles_m * tiles_n * splits;

  if (warp == NUM_EPI_WARPS) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int nt = item % tiles_n;
        const int mt = (item / tiles_n) % tiles_m;
        const int sp = item / (tiles_n * tiles_m);
        const int m0 = mt * BM, n0 = nt * BN;
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if (!A_MN) {
            tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(&tmA, &full_bar[stage], sa + j * 8192, m0 + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(&tmB, &full_bar[stage], sb + j * 8192, n0 + j * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == NUM_EPI_WARPS + 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c=f32 (bit4), a=b=bf16 (bits 7,10), majors (15,16), N>>3 (17), M>>4 (24)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t a_kstep = A_MN ? (16 * 128) >> 4 : (16 * 2) >> 4;  // descriptor units of 16 B per UMMA_K
      constexpr uint32_t b_kstep = B_MN ? (16 * 128) >> 4 : (16 * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int sp = item / (tiles_n * tiles_m);
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc<A_MN>(smem_u32(smem_a + stage * Cfg::A_BYTES));
          const uint64_t db = make_smem_desc<B_MN>(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            tcgen05_mma_f16(tmem_d, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int half = warp >> 2;         // column half of the tile
    const uint32_t stage = smem_u32(epi_stage) + warp * 4096;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int nt = item % tiles_n;
      const int mt = (item / tiles_n) % tiles_m;
      const int nbase = nt * BN + half * (BN / 2);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * (BN / 2);
      const int mrow0 = mt * BM + quarter * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 2 / 32; ++c) {
        float v[32];
        float4 b4, pre[8];
        const int n0 = nbase + c * 32;
        epi_prefetch<EF>(ep, lane, mrow0, n0, b4, pre);     // global reads first: latency overlaps the TMEM load
        tmem_ld32(taddr + c * 32, v);
        epi_finish<EF>(ep, v, stage, lane, mrow0, n0, b4, pre);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == NUM_EPI_WARPS + 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int BN, bool A_MN, bool B_MN, int EF>
int launch(const sc_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb, int splits, cudaStream_t st) {
  using Cfg = TileCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EF>;
  static sc_device_once once;  // per template instantiation
  if (once.first()) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    once.done();
  }
  const int tiles_m = ceil_div(d->M, BM), tiles_n = ceil_div(d->N, BN);
  const int kb_total = ceil_div(d->K, BK);
  int kb_per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kb_per);  // no empty split
  sc_gemm_desc dd = *d;
  dd.split_k = splits;
  EpiParams ep = make_epi(&dd);
  const int items = tiles_m * tiles_n * splits;
  const int grid = items < sc_num_sms() ? items : sc_num_sms();
  SC_CUDA(sc_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), Cfg::SMEM_BYTES, st, ta, tb, tiles_m, tiles_n, splits, kb_total,
                        kb_per, ep));
  return SC_OK;
}

}  // namespace

extern void sc_count_kernel(int kind, int n);
int sc_gemm_tc2(const sc_gemm_desc* d, cudaStream_t st);

// set by the 2-CTA launcher when the per-head row dots (sc_gemm_desc.dot_out) were fused into the epilogue
bool& sc_gemm_dot_fused() {
  static thread_local bool f = false;
  return f;
}
int sc_attn_delta(const void* o, const void* d_o, long bs, long rs, int B, int H, int L, float* delta, cudaStream_t st);
static int sc_gemm_tc_impl(const sc_gemm_desc* d, cudaStream_t st);

// Returns SC_ERR_UNSUPPORTED when the problem does not meet the TMA alignment rules (caller falls
// back to the FMA kernel only in fp32 mode; in bf16 mode this is an error).
int sc_gemm_tc(const sc_gemm_desc* d, cudaStream_t st) {
  if (!d->dot_out) return sc_gemm_tc_impl(d, st);
  if (!(d->dot_aux && d->dot_L > 0 && d->M % d->dot_L == 0 && d->N % 64 == 0 && d->c_dtype == SC_BF16 && !d->accumulate)) {
    sc_set_error("sc_gemm: dot_out needs dot_aux, dot_L dividing M, N %% 64 == 0 and a bf16 C");
    return SC_ERR_INVALID;
  }
  sc_gemm_dot_fused() = false;
  int rc = sc_gemm_tc_impl(d, st);
  if (rc) return rc;
  if (sc_gemm_dot_fused()) return SC_OK;
  // kernels without the fused epilogue (1-CTA tiles, other operand layouts): one pass over C and dot_aux after the GEMM
  return sc_attn_delta(d->dot_aux, d->C, (long)d->dot_L * d->ldc, d->ldc, d->M / d->dot_L, d->N / 64, d->dot_L, d->dot_out, st);
}

static int sc_gemm_tc_impl(const sc_gemm_desc* d, cudaStream_t st) {
  const bool a_mn = d->trans_a != 0, b_mn = d->trans_b != 0;
  auto aligned16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!(d->lda % 8 == 0 && d->ldb % 8 == 0 && aligned16(d->A) && aligned16(d->B) && d->N % 8 == 0 &&
        d->ldc % 8 == 0 && aligned16(d->C) && (!d->C2 || aligned16(d->C2)) &&
        (!d->bias || aligned16(d->bias)) && (!d->residual || (aligned16(d->residual) && d->ldr % (d->residual_dtype == SC_BF16 ? 8 : 4) == 0)) &&
        (!d->rowbias || (aligned16(d->rowbias) && d->ld_rowbias % 4 == 0)) && (!d->mul_aux || aligned16(d->mul_aux)))) {
    sc_set_error("sc_gemm(bf16): operands must be 16-byte aligned with leading dimensions multiple of 8 "
                 "(M=%d N=%d K=%d lda=%lld ldb=%lld ldc=%lld)", d->M, d->N, d->K, (long long)d->lda,
                 (long long)d->ldb, (long long)d->ldc);
    return SC_ERR_UNSUPPORTED;
  }
  {
    static const int use_2cta = [] { const char* e = getenv("SC_GEMM_2CTA"); return e ? atoi(e) : 1; }();
    if (use_2cta) {
      const int rc2 = sc_gemm_tc2(d, st);
      if (rc2 != SC_ERR_UNSUPPORTED) return rc2;
    }
  }
  const int waste256 = ceil_div(d->N, 256) * 256 - d->N, waste128 = ceil_div(d->N, 128) * 128 - d->N;
  const int BN = (waste256 <= waste128) ? 256 : 128;

  const int kb_total = ceil_div(d->K, BK);
  int splits = d->split_k;
  if (splits < 0) splits = sc_pick_splits(ceil_div(d->M, BM) * ceil_div(d->N, BN), kb_total, sc_num_sms());
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  if (splits > 1 && !(d->accumulate && d->c_dtype == SC_F32 && !d->C2)) {
    sc_set_error("sc_gemm: split_k > 1 requires accumulate=1 into an fp32 C and no C2");
    return SC_ERR_INVALID;
  }

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = sc_get_tensor_map(d->A, d->K, d->M, d->lda, BK, BM, &ta);
  else rc = sc_get_tensor_map(d->A, d->M, d->K, d->lda, 64, BK, &ta);
  if (rc) return rc;
  if (!b_mn) rc = sc_get_tensor_map(d->B, d->K, d->N, d->ldb, BK, BN, &tb);
  else rc = sc_get_tensor_map(d->B, d->N, d->K, d->ldb, 64, BK, &tb);
  if (rc) return rc;

  const int ef = sc_select_epilogue(d, splits);
  if (d->colsum_out && (ef == EF_GENERIC || (ef & (EF_OUT_F32 | EF_ATOMIC | EF_RESID)) != 0 || ((uintptr_t)d->colsum_out & 15) != 0)) {
    sc_set_error("sc_gemm: colsum_out needs a bf16-output specialised epilogue and a 16-byte aligned pointer");
    return SC_ERR_UNSUPPORTED;
  }
  sc_count_kernel(SC_K_GEMM_TC1, 1);
#define SC_L(BN_, A_, B_, EF_) return launch<BN_, A_, B_, EF_>(d, ta, tb, splits, st);
#define SC_DISPATCH(BN_)                                                                     \
  if (!a_mn && !b_mn) {                                                                      \
    if (ef == EF_BIAS) SC_L(BN_, false, false, EF_BIAS)                                      \
    if (ef == 0) SC_L(BN_, false, false, 0)                                                  \
    if (ef == (EF_BIAS | EF_QGELU | EF_C2)) SC_L(BN_, false, false, EF_BIAS | EF_QGELU | EF_C2) \
    if (ef == (EF_BIAS | EF_QGELU | EF_C2 | EF_C2_DERIV)) SC_L(BN_, false, false, EF_BIAS | EF_QGELU | EF_C2 | EF_C2_DERIV) \
    if (ef == (EF_BIAS | EF_GELU | EF_C2 | EF_C2_DERIV)) SC_L(BN_, false, false, EF_BIAS | EF_GELU | EF_C2 | EF_C2_DERIV) \
    if (ef == (EF_BIAS | EF_RESID | EF_OUT_F32)) SC_L(BN_, false, false, EF_BIAS | EF_RESID | EF_OUT_F32) \
    if (ef == (EF_BIAS | EF_GELU | EF_C2)) SC_L(BN_, false, false, EF_BIAS | EF_GELU | EF_C2)  \
    if (ef == EF_OUT_F32) SC_L(BN_, false, false, EF_OUT_F32)                                \
    SC_L(BN_, false, false, EF_GENERIC)                                                      \
  }                                                                                          \
  if (!a_mn && b_mn) {                                                                       \
    if (ef == 0) SC_L(BN_, false, true, 0)                                                   \
    if (ef == EF_MULAUX_QGELU) SC_L(BN_, false, true, EF_MULAUX_QGELU)                       \
    if (ef == EF_MULAUX_GELU) SC_L(BN_, false, true, EF_MULAUX_GELU)                         \
    if (ef == EF_MULAUX_DERIV) SC_L(BN_, false, true, EF_MULAUX_DERIV)                       \
    if (ef == EF_OUT_F32) SC_L(BN_, false, true, EF_OUT_F32)                                 \
    if (ef == (EF_OUT_F32 | EF_ACCUM)) SC_L(BN_, false, true, EF_OUT_F32 | EF_ACCUM)         \
    SC_L(BN_, false, true, EF_GENERIC)                                                       \
  }                                                                                          \
  if (a_mn && b_mn) {                                                                        \
    if (ef == (EF_ATOMIC | EF_OUT_F32)) SC_L(BN_, true, true, EF_ATOMIC | EF_OUT_F32)        \
    SC_L(BN_, true, true, EF_GENERIC)                                                        \
  }                                                                                          \
  SC_L(BN_, true, false, EF_GENERIC)
  if (BN == 256) { SC_DISPATCH(256) }
  SC_DISPATCH(128)
#undef SC_DISPATCH
#undef SC_L
}
