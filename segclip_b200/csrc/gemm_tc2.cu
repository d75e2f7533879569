// 2-CTA tcgen05 GEMM (cta_group::2): a pair of SMs computes one 256 x 256 tile.
//
// Each CTA of the cluster stages its 128 rows of A and its 128-row half of B (32 KB per stage instead of 48 KB for a
// 128 x 256 tile), so L2 -> SM operand traffic per FLOP drops by 1.5x and the smem ring is 6 deep.  The leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16); every CTA keeps its own 128 x 256 accumulator in its TMEM
// (two buffers) and runs its own epilogue.  Synchronisation:
//   full[stage]  (leader's): both CTAs' TMA loads complete_tx on the leader's barrier (peer bit cleared)
//   empty[stage] (both)    : tcgen05.commit ... multicast::cluster 0b11 from the leader's MMA thread
//   tmem_full[acc] (both)  : same multicast commit after the last k-block
//   tmem_empty[acc] (leader's, 2 x EPI2_WARPS arrivals): epilogue warps of both CTAs arrive through the cluster window
// Everything else (operand layouts, split-K, fused epilogues) is identical to gemm_tc.cu.
#include "gemm_tc_common.cuh"

extern void sc_count_kernel(int kind, int n);

#ifdef SC_GEMM_TRACE
// debug build (-DSC_GEMM_TRACE, tools/trace_gemm.py): globaltimer at the start and the end of every CTA of the last launch --
// how far apart the statically scheduled CTA pairs finish
__device__ unsigned long long g_gemm_trace[2][256];
SC_DEVINL unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

namespace {
using namespace tc;

#ifndef SC_TC2_EPI_WARPS
#define SC_TC2_EPI_WARPS 8
#endif
// 8 epilogue warps (4 TMEM lane quarters x 2 column halves); 16 were measured on B200 and change nothing: the epilogue is
// bound by L1TEX data-pipe wavefronts, not by latency (see epi_finish_tma).  Each warp owns 8 KB of staging (four 2 KB TMA
// boxes, or one 4 KB fp32 transpose patch), which costs one stage of the operand ring (5 x 32 KB).
constexpr int EPI2_WARPS = SC_TC2_EPI_WARPS;
constexpr int EPI2_COLS = 256 / (EPI2_WARPS / 4);      // accumulator columns per warp
constexpr int THREADS2 = (EPI2_WARPS + 2) * 32;
constexpr int STAGES2 = 5;
constexpr int EPI2_STAGE_BYTES = 8192;   // per epilogue warp
constexpr int A2_BYTES = 128 * BK * 2;
constexpr int B2_BYTES = 128 * BK * 2;
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + EPI2_WARPS * EPI2_STAGE_BYTES + 256 + EPI2_WARPS * 4 * 8;   // + input-tile mbarriers
#ifndef SC_TC2_MAX_SHIFT_BMN
#define SC_TC2_MAX_SHIFT_BMN 2      // column slices of an MN-major B operand: 2^shift per tile (64 columns = 32 per CTA at 2)
#endif
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair

SC_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SC_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// (the *_e wrappers are called by a whole warp under uniform control flow; one elected lane executes -- see
// tcgen05_mma_f16_e in gemm_tc_common.cuh)
SC_DEVINL void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(leader_bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
SC_DEVINL void tcgen05_mma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
SC_DEVINL void tcgen05_commit2(uint64_t* bar) {   // arrives on the same barrier of both CTAs
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
SC_DEVINL void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

template <bool A_MN, bool B_MN, int EF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                const __grid_constant__ CUtensorMap tmX, int tiles_m, int tiles_n,
                int splits, int kb_total, int kb_per_split, int full_items, int sub_shift, EpiParams ep) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES2 * A2_BYTES;
  uint8_t* epi_stage = smem + STAGES2 * STAGE2_BYTES;                       // 1024-byte aligned (swizzled TMA boxes)
  uint64_t* bars = (uint64_t*)(epi_stage + EPI2_WARPS * EPI2_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES2;
  uint64_t* tmem_full = bars + 2 * STAGES2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  uint64_t* aux_bar = (uint64_t*)(epi_stage + EPI2_WARPS * EPI2_STAGE_BYTES + 256);     // [epilogue warp][chunk]
  constexpr bool kAux = EpiAuxTma<EF>::value;      // bf16 input tile (residual / activation-gradient operand): 4 boxes of 2 KB
  constexpr bool kRes = EpiResTma<EF>::value;      // fp32 residual tile: 2 boxes of 4 KB
  constexpr bool kIn = kAux || kRes;
  constexpr int NBOX = kRes ? 2 : 4;
  constexpr uint32_t BOXB = kRes ? 4096u : 2048u;
  static_assert(!kIn || EPI2_COLS == 128, "input-tile epilogue: four chunks per warp and tile");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EPI2_WARPS);
    }
    if constexpr (kIn) {
      for (int i = 0; i < EPI2_WARPS * 4; ++i) mbar_init(&aux_bar[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == EPI2_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Work items.  The first `full_items` are whole 256 x 256 tiles (x split-K slices); the tiles of an incomplete last wave
  // (splits == 1 only) are cut into 2^sub_shift column slices so that the tail is spread over all CTA pairs instead of
  // running on a few of them at full-tile cost (text tower: 154 tiles on 74 pairs = 3 waves for 2.08 of work -> 2.25).
  const int num_items = full_items + ((tiles_m * tiles_n * splits - full_items) << sub_shift);
  struct Item { int mt, nt, sp, ncol0, ncols; };
  auto decode = [&](int item) {
    Item t;
    int tile = item, q = 0;
    t.ncols = 256;
    if (item >= full_items) {
      const int j = item - full_items;
      tile = full_items + (j >> sub_shift);
      q = j & ((1 << sub_shift) - 1);
      t.ncols = 256 >> sub_shift;
    }
    t.nt = tile % tiles_n;
    t.mt = (tile / tiles_n) % tiles_m;
    t.sp = tile / (tiles_n * tiles_m);
    t.ncol0 = t.nt * 256 + q * t.ncols;
    return t;
  };
  pdl_launch_dependents();      // after the TMEM allocation: a successor CTA sharing this SM can never starve this one
  pdl_wait();                   // everything above ran under the previous kernel's tail; global memory from here on
#ifdef SC_GEMM_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 256) g_gemm_trace[0][blockIdx.x] = gtimer();
#endif

  if (warp == EPI2_WARPS) {
    // =============================== TMA producer (both CTAs; whole warp, elected issue) ===============================
    {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < num_items; item += npairs) {
        const Item t = decode(item);
        // each CTA stages its half of the item's columns (a 128-row box whatever the slice width: a narrow slice reads ahead)
        const int m0 = t.mt * 256 + rank * 128, n0 = t.ncol0 + rank * (t.ncols >> 1);
        const int kb0 = t.sp * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        // (prefetching the next tile's operands into L2 from here was measured on B200: -15 % on the K = 768 shapes)
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx_e(&full_bar[stage], 2 * STAGE2_BYTES);
          uint8_t* sa = smem_a + stage * A2_BYTES;
          uint8_t* sb = smem_b + stage * B2_BYTES;
          if (!A_MN) {
            tma_load_2d_2sm(&tmA, &full_bar[stage], sa, kb * BK, m0);
          } else {
            tma_load_2d_2sm(&tmA, &full_bar[stage], sa, m0, kb * BK);
            tma_load_2d_2sm(&tmA, &full_bar[stage], sa + 8192, m0 + 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(&tmB, &full_bar[stage], sb, kb * BK, n0);
          } else {
            tma_load_2d_2sm(&tmB, &full_bar[stage], sb, n0, kb * BK);
            tma_load_2d_2sm(&tmB, &full_bar[stage], sb + 8192, n0 + 64, kb * BK);
          }
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI2_WARPS + 1) {
    // =============================== MMA issuer (leader CTA only; whole warp, elected issue) ===============================
    if (rank == 0) {
      constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                  ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t a_kstep = A_MN ? (16 * 128) >> 4 : (16 * 2) >> 4;
      constexpr uint32_t b_kstep = B_MN ? (16 * 128) >> 4 : (16 * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = pair; item < num_items; item += npairs, ++it) {
        const Item t = decode(item);
        const uint32_t idesc = idesc0 | ((uint32_t)(t.ncols >> 3) << 17);      // N of the pair's MMA = slice width
        const int kb0 = t.sp * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc<A_MN>(smem_u32(smem_a + stage * A2_BYTES));
          const uint64_t db = make_smem_desc<B_MN>(smem_u32(smem_b + stage * B2_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tcgen05_mma2_f16(tmem_d, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc,
                             (kb > kb0 || k > 0) ? 1u : 0u);
          tcgen05_commit2(&empty_bar[stage]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit2(&tmem_full[acc]);
      }
    }
  } else {
    // =============================== epilogue (both CTAs, own 128 rows) ===============================
    const int quarter = warp & 3;
    const int cgrp = warp >> 2;
    const uint32_t stage = smem_u32(epi_stage) + warp * EPI2_STAGE_BYTES;
    uint32_t g = 0;
    uint32_t aux_phase = 0;          // bit c: parity of this warp's input-tile barrier c
    float dotacc = 0.f;              // EF_ROWDOT: running per-head row dot of this lane's row
    int it = 0;
    for (int item = pair; item < num_items; item += npairs, ++it) {
      const Item t = decode(item);
      const int nbase = t.ncol0 + cgrp * EPI2_COLS;
      const int nlim = min(ep.N, t.ncol0 + t.ncols);      // a column slice ends before the tile does (multiple of 64)
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mrow0 = t.mt * 256 + rank * 128 + quarter * 32;
      float4 breg = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr ((EpiTma<EF>::value || kIn) && (EF & EF_BIAS) != 0) {
        if (nbase + 4 * lane < ep.N) breg = __ldg((const float4*)(ep.bias + nbase + 4 * lane));
      }
      auto issue_in = [&](int c) {                                // input tile of chunk c -> box c % NBOX (whole warp calls)
        mbar_expect_tx_e(&aux_bar[warp * 4 + c % NBOX], BOXB);
        tma_load_2d_e(&tmX, &aux_bar[warp * 4 + c % NBOX], epi_stage + warp * EPI2_STAGE_BYTES + (c % NBOX) * BOXB, nbase + c * 32, mrow0);
      };
      if constexpr (kIn) {
        // the input tiles of this warp's first chunks: issued now, they land under the tile's main loop
        bulk_wait_read<0>();           // (elected lane) the previous tile's stores have read the boxes
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NBOX; ++c)
          if (nbase + c * 32 < nlim && mrow0 < ep.M) issue_in(c);          // warp-uniform
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256 + cgrp * EPI2_COLS;
#pragma unroll 1
      for (int c = 0; c < EPI2_COLS / 32; ++c) {
        float v[32];
        const int n0 = nbase + c * 32;
        if (n0 >= nlim) break;                                    // past the item's column slice / the matrix (warp-uniform)
        if constexpr (kIn) {
          tmem_ld32(taddr + c * 32, v);
          if (n0 < nlim && mrow0 < ep.M) {                        // warp-uniform
            const int bx = c % NBOX;
            mbar_wait(&aux_bar[warp * 4 + bx], (aux_phase >> bx) & 1u);
            aux_phase ^= 1u << bx;
            if constexpr (kAux) epi_finish_aux_tma<EF>(ep, v, stage + bx * BOXB, lane, mrow0, n0, c, breg, &tmC, dotacc);
            else epi_finish_res_tma<EF>(ep, v, stage + bx * BOXB, lane, mrow0, n0, c, breg, &tmC);
            if (NBOX < 4 && c + NBOX < 4 && n0 + NBOX * 32 < nlim) {      // box reused within the tile (fp32 boxes)
              bulk_wait_read<0>();
              __syncwarp();
              issue_in(c + NBOX);
            }
          }
        } else if constexpr (EpiTma<EF>::value) {
          tmem_ld32(taddr + c * 32, v);
          epi_finish_tma<EF>(ep, v, stage, lane, mrow0, n0, g++, c, breg, &tmC, &tmC2);
        } else {
          float4 b4, pre[8];
          epi_prefetch<EF>(ep, lane, mrow0, n0, b4, pre);     // global reads first: latency overlaps the TMEM load
          tmem_ld32(taddr + c * 32, v);
          epi_finish<EF>(ep, v, stage, lane, mrow0, n0, b4, pre);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
    }
  }

  if constexpr (EpiTma<EF>::value || kIn) {
    if (warp < EPI2_WARPS) bulk_wait_all();                  // staging smem must outlive the last TMA stores
  }
#ifdef SC_GEMM_TRACE
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < 256) g_gemm_trace[1][blockIdx.x] = gtimer();
#endif
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == EPI2_WARPS + 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <bool A_MN, bool B_MN, int EF>
int launch2(const sc_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb, int splits, cudaStream_t st) {
  CUtensorMap tc_ = ta, tc2_ = ta;          // placeholders for the kinds that do not store through the TMA
  CUtensorMap tx_ = ta;
  if constexpr (EpiAuxTma<EF>::value) {
    int rc;
    if ((rc = sc_get_tensor_map_sw(d->C, d->N, d->M, d->ldc, 32, 32, 64, &tc_))) return rc;
    if constexpr ((EF & EF_RESID_BF) != 0) rc = sc_get_tensor_map_sw(d->residual, d->N, d->M, d->ldr, 32, 32, 64, &tx_);
    else if constexpr ((EF & EF_ROWDOT) != 0) rc = sc_get_tensor_map_sw(d->dot_aux, d->N, d->M, d->ldc, 32, 32, 64, &tx_);
    else rc = sc_get_tensor_map_sw(d->mul_aux, d->N, d->M, d->ldc, 32, 32, 64, &tx_);
    if (rc) return rc;
  }
  if constexpr (EpiResTma<EF>::value) {
    int rc;
    if ((rc = sc_get_tensor_map_any(d->C, d->N, d->M, d->ldc, 32, 32, 128, 4, &tc_))) return rc;
    if ((rc = sc_get_tensor_map_any(d->residual, d->N, d->M, d->ldr, 32, 32, 128, 4, &tx_))) return rc;
  }
  if constexpr (EpiTma<EF>::value) {
    int rc;
    if ((rc = sc_get_tensor_map_sw(d->C, d->N, d->M, d->ldc, 32, 32, 64, &tc_))) return rc;
    if constexpr (EpiTma<EF>::two) {
      if ((rc = sc_get_tensor_map_sw(d->C2, d->N, d->M, d->ldc, 32, 32, 64, &tc2_))) return rc;
    }
  }
  auto kern = gemm_tc2_kernel<A_MN, B_MN, EF>;
  static sc_device_once once;  // per template instantiation
  if (once.first()) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
    once.done();
  }
  const int tiles_m = ceil_div(d->M, 256), tiles_n = ceil_div(d->N, 256);
  const int kb_total = ceil_div(d->K, BK);
  int kb_per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kb_per);
  sc_gemm_desc dd = *d;
  dd.split_k = splits;
  EpiParams ep = make_epi(&dd);
  const int items = tiles_m * tiles_n * splits;
  const int max_pairs = sc_num_sms() / 2;
  const int pairs = items < max_pairs ? items : max_pairs;
  // Incomplete last wave (no split-K): cut its tiles into 2 or 4 column slices when that shortens the schedule.  An item
  // costs its k-blocks x the slice's share of the MMA time + a fixed part (pipeline fill, epilogue tail) of ~3 k-blocks.
  int full_items = items, sub_shift = 0;
  const int rem = items % pairs;
  static const int tail_split = getenv("SC_GEMM_TAIL_SPLIT") ? atoi(getenv("SC_GEMM_TAIL_SPLIT")) : 1;
  if (tail_split && splits == 1 && rem != 0 && items > pairs && (EF & (EF_ROWDOT | EF_ATOMIC)) == 0 && EF != EF_GENERIC) {
    double best = kb_total + 3.0;
    for (int sh = 1; sh <= 2; ++sh) {
      if (B_MN && sh > SC_TC2_MAX_SHIFT_BMN) break;
      const int waves = ceil_div((long)rem << sh, pairs);
      const double cost = waves * ((double)kb_total / (1 << sh) + 3.0);
      if (cost < best * 0.9) {
        best = cost;
        sub_shift = sh;
      }
    }
    if (sub_shift) {
      full_items = items - rem;
      sc_count_kernel(SC_K_GEMM_TC2_TAIL, 0);
    }
  }
  SC_CUDA(sc_launch_pdl(kern, dim3(2 * pairs), dim3(THREADS2), SMEM2_BYTES, st, ta, tb, tc_, tc2_, tx_, tiles_m, tiles_n, splits,
                        kb_total, kb_per, full_items, sub_shift, ep));
  return SC_OK;
}

}  // namespace

extern void sc_count_kernel(int kind, int n);
bool& sc_gemm_dot_fused();

// Same contract as sc_gemm_tc; the caller has already validated alignment.  Returns SC_ERR_UNSUPPORTED when the shape is
// a poor fit for 256 x 256 pair tiles (the 1-CTA kernel then runs).
int sc_gemm_tc2(const sc_gemm_desc* d, cudaStream_t st) {
  using namespace tc;
  const bool a_mn = d->trans_a != 0, b_mn = d->trans_b != 0;
  if (d->M < 512 || d->N < 256 || (a_mn && !b_mn)) return SC_ERR_UNSUPPORTED;
  const int waste = ceil_div(d->N, 256) * 256 - d->N;
  if (waste * 8 > d->N) return SC_ERR_UNSUPPORTED;       // > 12.5 % padded columns: 128-wide tiles fit better
  const int kb_total = ceil_div(d->K, BK);
  int splits = d->split_k;
  if (splits < 0) splits = sc_pick_splits(ceil_div(d->M, 256) * ceil_div(d->N, 256), kb_total, sc_num_sms() / 2);
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  if (splits > 1 && !(d->accumulate && d->c_dtype == SC_F32 && !d->C2)) {
    sc_set_error("sc_gemm: split_k > 1 requires accumulate=1 into an fp32 C and no C2");
    return SC_ERR_INVALID;
  }
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = sc_get_tensor_map(d->A, d->K, d->M, d->lda, BK, 128, &ta);
  else rc = sc_get_tensor_map(d->A, d->M, d->K, d->lda, 64, BK, &ta);
  if (rc) return rc;
  if (!b_mn) rc = sc_get_tensor_map(d->B, d->K, d->N, d->ldb, BK, 128, &tb);
  else rc = sc_get_tensor_map(d->B, d->N, d->K, d->ldb, 64, BK, &tb);
  if (rc) return rc;
  const int ef = sc_select_epilogue(d, splits);
  if (d->colsum_out && (ef == EF_GENERIC || (ef & (EF_OUT_F32 | EF_ATOMIC | EF_RESID)) != 0 || ((uintptr_t)d->colsum_out & 15) != 0)) {
    sc_set_error("sc_gemm: colsum_out needs a bf16-output specialised epilogue and a 16-byte aligned pointer");
    return SC_ERR_UNSUPPORTED;
  }
  sc_count_kernel(SC_K_GEMM_TC2, 1);
#define SC_L2(A_, B_, EF_) return launch2<A_, B_, EF_>(d, ta, tb, splits, st);
  if (!a_mn && !b_mn) {
    if (ef == EF_BIAS) SC_L2(false, false, EF_BIAS)
    if (ef == 0) SC_L2(false, false, 0)
    if (ef == (EF_BIAS | EF_QGELU | EF_C2)) SC_L2(false, false, EF_BIAS | EF_QGELU | EF_C2)
    if (ef == (EF_BIAS | EF_QGELU | EF_C2 | EF_C2_DERIV)) SC_L2(false, false, EF_BIAS | EF_QGELU | EF_C2 | EF_C2_DERIV)
    if (ef == (EF_BIAS | EF_GELU | EF_C2 | EF_C2_DERIV)) SC_L2(false, false, EF_BIAS | EF_GELU | EF_C2 | EF_C2_DERIV)
    if (ef == (EF_BIAS | EF_RESID | EF_OUT_F32)) SC_L2(false, false, EF_BIAS | EF_RESID | EF_OUT_F32)
    if (ef == (EF_BIAS | EF_RESID_BF)) SC_L2(false, false, EF_BIAS | EF_RESID_BF)
    if (ef == (EF_BIAS | EF_GELU | EF_C2)) SC_L2(false, false, EF_BIAS | EF_GELU | EF_C2)
    if (ef == EF_OUT_F32) SC_L2(false, false, EF_OUT_F32)
    SC_L2(false, false, EF_GENERIC)
  }
  if (!a_mn && b_mn) {
    if (ef == 0 && d->dot_out) {
      sc_gemm_dot_fused() = true;
      SC_L2(false, true, EF_ROWDOT)
    }
    if (ef == 0) SC_L2(false, true, 0)
    if (ef == EF_MULAUX_QGELU) SC_L2(false, true, EF_MULAUX_QGELU)
    if (ef == EF_MULAUX_GELU) SC_L2(false, true, EF_MULAUX_GELU)
    if (ef == EF_MULAUX_DERIV) SC_L2(false, true, EF_MULAUX_DERIV)
    if (ef == EF_OUT_F32) SC_L2(false, true, EF_OUT_F32)
    if (ef == (EF_OUT_F32 | EF_ACCUM)) SC_L2(false, true, EF_OUT_F32 | EF_ACCUM)
    SC_L2(false, true, EF_GENERIC)
  }
  if (ef == (EF_ATOMIC | EF_OUT_F32)) SC_L2(true, true, EF_ATOMIC | EF_OUT_F32)
  SC_L2(true, true, EF_GENERIC)
#undef SC_L2
}

#ifdef SC_GEMM_TRACE
extern "C" int sc_debug_gemm_trace(unsigned long long* out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_gemm_trace, sizeof(unsigned long long) * 2 * 256);
}
#endif
