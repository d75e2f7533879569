// Tensor-core attention core (bf16, fp32 accumulate) for the self-attention shapes of the path:
// vision L=196/48 (hd 64), text L=77 causal (hd 64), MAE decoder L=197 (hd 48).
// Replaces the bmm/softmax/bmm of nn.MultiheadAttention (modules/module_seg_vit.py:189,
// module_clip_ttransformer.py:46) and timm Attention (module_mae.py:122-135) and their backward.
//
// Flash-style: no [L, L] matrix ever touches HBM.  One CTA = up to 8 warps, one warp = 16 rows.
//   fwd : K, V of the (batch, head) staged once in swizzled smem; per warp online softmax over 64-key blocks
//   bwd : (1) dQ pass  - warp owns 16 queries, K/V in smem: S, dP -> dS -> dQ += dS K; also writes delta = rowsum(dO o O)
//         (2) dK/dV pass - warp owns 16 keys, Q/dO in smem: S^T, dP^T -> dV += P^T dO, dK += dS^T Q
// MMA: mma.sync.m16n8k16 bf16 with ldmatrix operand fetch (legacy tensor path; the GEMMs of the path use
// tcgen05 -- attention is ~4% of the FLOPs, a tcgen05/TMEM version is the planned next step).
#include "common.cuh"

namespace {

constexpr int AW = 8;              // max warps per CTA (each warp loops over 16-row blocks)
constexpr int AT = AW * 32;
constexpr float LOG2E = 1.4426950408889634f;

SC_DEVINL uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SC_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
SC_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
SC_DEVINL void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
SC_DEVINL uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}
SC_DEVINL float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
SC_DEVINL float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// smem tile: rows of 128 B (64 bf16; HD < 64 leaves the tail unused), 16-byte chunks XOR-swizzled by row&7
SC_DEVINL uint32_t sw(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// Stage `rows` rows (zero beyond) of a [*, HD] slice with row stride rs into a swizzled tile of rows_pad rows,
// asynchronously (LDGSTS: no register round trip); complete with stage_wait().
template <int HD>
SC_DEVINL void stage_tile(uint8_t* dst, const bf16* src, long rs, int rows, int rows_pad) {
  constexpr int CH = HD / 8;
  const uint32_t d0 = smem_addr(dst);
  for (int idx = threadIdx.x; idx < rows_pad * CH; idx += blockDim.x) {
    const int r = idx / CH, c = idx - r * CH;
    const bf16* g = src + (long)(r < rows ? r : 0) * rs + c * 8;
    const int nbytes = r < rows ? 16 : 0;      // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + sw(r, c)), "l"(g), "r"(nbytes) : "memory");
  }
}
SC_DEVINL void stage_wait() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}
SC_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A fragments (16 rows x HD) of a row-major global matrix; rows >= nrows read as zero.
template <int HD>
SC_DEVINL void load_a_frags(uint32_t (&f)[HD / 16][4], const bf16* base, long rs, int row0, int nrows, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    const int c = ks * 16 + 2 * t;
    f[ks][0] = r0 < nrows ? *(const uint32_t*)(base + (long)r0 * rs + c) : 0u;
    f[ks][1] = r1 < nrows ? *(const uint32_t*)(base + (long)r1 * rs + c) : 0u;
    f[ks][2] = r0 < nrows ? *(const uint32_t*)(base + (long)r0 * rs + c + 8) : 0u;
    f[ks][3] = r1 < nrows ? *(const uint32_t*)(base + (long)r1 * rs + c + 8) : 0u;
  }
}

// acc[16 x 16*NP] += A(16 x HD) * T[rows r0.., HD]^T   with T an smem tile whose rows are the n index
template <int HD, int NP>
SC_DEVINL void mma_a_tileT(float (&acc)[2 * NP][4], const uint32_t (&a)[HD / 16][4], uint32_t tile, int r0, int np_valid,
                           int lane) {
  const int m = lane >> 3, l8 = lane & 7;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (p < np_valid) {
        uint32_t b0, b1, b2, b3;
        const int row = r0 + p * 16 + (m >> 1) * 8 + l8;
        ldsm_x4(tile + sw(row, ks * 2 + (m & 1)), b0, b1, b2, b3);
        mma_bf16(acc[2 * p], a[ks], b0, b1);
        mma_bf16(acc[2 * p + 1], a[ks], b2, b3);
      }
    }
  }
}

// acc[16 x HD] += A(16 x 16 single k-step) * T[rows r0..r0+15, HD]   (T rows are the k index: transposed fetch)
template <int HD>
SC_DEVINL void mma_a_tile(float (&acc)[HD / 8][4], const uint32_t (&a)[4], uint32_t tile, int r0, int lane) {
  const int m = lane >> 3, l8 = lane & 7;
#pragma unroll
  for (int dp = 0; dp < HD / 16; ++dp) {
    uint32_t b0, b1, b2, b3;
    const int row = r0 + (m & 1) * 8 + l8;
    ldsm_x4_t(tile + sw(row, dp * 2 + (m >> 1)), b0, b1, b2, b3);
    mma_bf16(acc[2 * dp], a, b0, b1);
    mma_bf16(acc[2 * dp + 1], a, b2, b3);
  }
}

template <int HD>
SC_DEVINL void store_rows(bf16* base, long rs, int row0, int nrows, const float (&acc)[HD / 8][4], float s0, float s1, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    const int c = j * 8 + 2 * t;
    if (r0 < nrows) *(uint32_t*)(base + (long)r0 * rs + c) = pack2(acc[j][0] * s0, acc[j][1] * s0);
    if (r1 < nrows) *(uint32_t*)(base + (long)r1 * rs + c) = pack2(acc[j][2] * s1, acc[j][3] * s1);
  }
}

// =================================================================================== forward
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(AT, 2) attn_fwd_mma_kernel(sc_attn_desc a, int lk_pad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + lk_pad * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Lq = a.Lq, Lk = a.Lk;
  stage_tile<HD>(sK, (const bf16*)a.k + (long)b * a.k_bs + h * HD, a.k_rs, Lk, lk_pad);
  stage_tile<HD>(sV, (const bf16*)a.v + (long)b * a.v_bs + h * HD, a.v_rs, Lk, lk_pad);
  stage_wait();
  const uint32_t tK = smem_addr(sK), tV = smem_addr(sV);
  const float c = a.scale * LOG2E;
  const bf16* qb = (const bf16*)a.q + (long)b * a.q_bs + h * HD;
  const int nwarps = blockDim.x >> 5;
  for (int row0 = warp * 16; row0 < Lq; row0 += nwarps * 16) {
  uint32_t qf[HD / 16][4];
  load_a_frags<HD>(qf, qb, a.q_rs, row0, Lq, lane);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  const int kend = CAUSAL ? min(Lk, row0 + 16) : Lk;
  for (int kb = 0; kb < kend; kb += 64) {
    const int np = min(4, (kend - kb + 15) >> 4);
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    mma_a_tileT<HD, 4>(s, qf, tK, kb, np, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
    const bool need_mask = (kb + 64 > kend) || (CAUSAL && kb + 64 > row0);     // warp-uniform
    if (need_mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = kb + j * 8 + 2 * t + (e & 1);
          const int row = row0 + g + ((e >> 1) << 3);
          const bool ok = col < kend && (!CAUSAL || col <= row);
          s[j][e] = ok ? s[j][e] : -INFINITY;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
    const float al0 = ex2((m0 - n0) * c), al1 = ex2((m1 - n1) * c);
    m0 = n0;
    m1 = n1;
    l0 *= al0;
    l1 *= al1;
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
      o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1;
    }
    const float mc0 = m0 * c, mc1 = m1 * c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(fmaf(s[j][0], c, -mc0));
      s[j][1] = ex2(fmaf(s[j][1], c, -mc0));
      s[j][2] = ex2(fmaf(s[j][2], c, -mc1));
      s[j][3] = ex2(fmaf(s[j][3], c, -mc1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < np) {
        uint32_t pa[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                          pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
        mma_a_tile<HD>(o, pa, tV, kb + kk * 16, lane);
      }
    }
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  bf16* ob = (bf16*)a.o + (long)b * a.o_bs + h * HD;
  store_rows<HD>(ob, a.o_rs, row0, Lq, o, 1.f / l0, 1.f / l1, lane);
  if (t == 0 && a.lse) {
    float* lse = a.lse + ((long)b * a.H + h) * Lq;
    if (row0 + g < Lq) lse[row0 + g] = m0 * a.scale + logf(l0);
    if (row0 + g + 8 < Lq) lse[row0 + g + 8] = m1 * a.scale + logf(l1);
  }
  }   // row-block loop
}

// =================================================================================== backward: dQ (+ delta)
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(AT, 2) attn_bwd_dq_mma_kernel(sc_attn_bwd_desc gd, float* __restrict__ delta, int lk_pad) {
  const sc_attn_desc& a = gd.fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + lk_pad * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Lq = a.Lq, Lk = a.Lk;
  stage_tile<HD>(sK, (const bf16*)a.k + (long)b * a.k_bs + h * HD, a.k_rs, Lk, lk_pad);
  stage_tile<HD>(sV, (const bf16*)a.v + (long)b * a.v_bs + h * HD, a.v_rs, Lk, lk_pad);
  stage_wait();
  const long qoff = (long)b * a.q_bs + h * HD, ooff = (long)b * a.o_bs + h * HD;
  const int nwarps = blockDim.x >> 5;
  for (int row0 = warp * 16; row0 < Lq; row0 += nwarps * 16) {
  uint32_t qf[HD / 16][4], dof[HD / 16][4];
  load_a_frags<HD>(qf, (const bf16*)a.q + qoff, a.q_rs, row0, Lq, lane);
  load_a_frags<HD>(dof, (const bf16*)gd.d_o + ooff, a.o_rs, row0, Lq, lane);
  // delta = rowsum(dO o O)
  float d0 = 0.f, d1 = 0.f;
  {
    uint32_t of[HD / 16][4];
    load_a_frags<HD>(of, (const bf16*)a.o + ooff, a.o_rs, row0, Lq, lane);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = __bfloat1622float2(*(const __nv_bfloat162*)&dof[ks][e]);
        const float2 y = __bfloat1622float2(*(const __nv_bfloat162*)&of[ks][e]);
        const float p = x.x * y.x + x.y * y.y;
        if (e & 1) d1 += p; else d0 += p;
      }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
  }
  const float* lse = a.lse + ((long)b * a.H + h) * Lq;
  float* dl = delta + ((long)b * a.H + h) * Lq;
  const int r0 = row0 + g, r1 = row0 + g + 8;
  const float lse0 = r0 < Lq ? lse[r0] * LOG2E : 0.f, lse1 = r1 < Lq ? lse[r1] * LOG2E : 0.f;
  if (t == 0) {
    if (r0 < Lq) dl[r0] = d0;
    if (r1 < Lq) dl[r1] = d1;
  }
  const uint32_t tK = smem_addr(sK), tV = smem_addr(sV);
  const float c = a.scale * LOG2E;
  float dq[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  const int kend = CAUSAL ? min(Lk, row0 + 16) : Lk;
  for (int kb = 0; kb < kend; kb += 32) {
    const int np = min(2, (kend - kb + 15) >> 4);
    float s[4][4], dp[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
    mma_a_tileT<HD, 2>(s, qf, tK, kb, np, lane);
    mma_a_tileT<HD, 2>(dp, dof, tV, kb, np, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = kb + j * 8 + 2 * t + (e & 1);
        const int row = row0 + g + ((e >> 1) << 3);
        const bool ok = col < kend && (!CAUSAL || col <= row);
        const float p = ok ? ex2(fmaf(s[j][e], c, -((e >> 1) ? lse1 : lse0))) : 0.f;
        s[j][e] = p * (dp[j][e] - ((e >> 1) ? d1 : d0));     // dS
      }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      if (kk < np) {
        uint32_t pa[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                          pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
        mma_a_tile<HD>(dq, pa, tK, kb + kk * 16, lane);
      }
    }
  }
  store_rows<HD>((bf16*)gd.d_q + qoff, a.q_rs, row0, Lq, dq, a.scale, a.scale, lane);
  }   // row-block loop
}

// =================================================================================== backward: dK, dV
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(AT, 2) attn_bwd_dkv_mma_kernel(sc_attn_bwd_desc gd, const float* __restrict__ delta,
                                                                 int lq_pad) {
  const sc_attn_desc& a = gd.fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + lq_pad * 128;
  float* sLse = (float*)(smem + 2 * lq_pad * 128);
  float* sDelta = sLse + lq_pad;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Lq = a.Lq, Lk = a.Lk;
  stage_tile<HD>(sQ, (const bf16*)a.q + (long)b * a.q_bs + h * HD, a.q_rs, Lq, lq_pad);
  stage_tile<HD>(sdO, (const bf16*)gd.d_o + (long)b * a.o_bs + h * HD, a.o_rs, Lq, lq_pad);
  for (int i = threadIdx.x; i < lq_pad; i += blockDim.x) {
    const long o = ((long)b * a.H + h) * Lq + i;
    sLse[i] = i < Lq ? a.lse[o] * LOG2E : 0.f;
    sDelta[i] = i < Lq ? delta[o] : 0.f;
  }
  stage_wait();
  const int nwarps = blockDim.x >> 5;
  for (int key0 = warp * 16; key0 < Lk; key0 += nwarps * 16) {
  const long koff = (long)b * a.k_bs + h * HD, voff = (long)b * a.v_bs + h * HD;
  uint32_t kf[HD / 16][4], vf[HD / 16][4];
  load_a_frags<HD>(kf, (const bf16*)a.k + koff, a.k_rs, key0, Lk, lane);
  load_a_frags<HD>(vf, (const bf16*)a.v + voff, a.v_rs, key0, Lk, lane);
  const uint32_t tQ = smem_addr(sQ), tdO = smem_addr(sdO);
  const float c = a.scale * LOG2E;
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  const int qstart = CAUSAL ? (key0 & ~15) : 0;     // queries < key0 never see these keys
  for (int q0 = qstart; q0 < Lq; q0 += 16) {
    float st[2][4], dpt[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
    }
    mma_a_tileT<HD, 1>(st, kf, tQ, q0, 1, lane);       // S^T  [16 keys x 16 queries]
    mma_a_tileT<HD, 1>(dpt, vf, tdO, q0, 1, lane);     // dP^T
    float pt[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = q0 + j * 8 + 2 * t + (e & 1);
        const int key = key0 + g + ((e >> 1) << 3);
        const bool ok = qi < Lq && (!CAUSAL || key <= qi);
        const float p = ok ? ex2(fmaf(st[j][e], c, -sLse[qi])) : 0.f;
        pt[j][e] = p;
        st[j][e] = p * (dpt[j][e] - sDelta[qi]);       // dS^T
      }
    uint32_t pa[4] = {pack2(pt[0][0], pt[0][1]), pack2(pt[0][2], pt[0][3]), pack2(pt[1][0], pt[1][1]), pack2(pt[1][2], pt[1][3])};
    uint32_t da[4] = {pack2(st[0][0], st[0][1]), pack2(st[0][2], st[0][3]), pack2(st[1][0], st[1][1]), pack2(st[1][2], st[1][3])};
    mma_a_tile<HD>(dv, pa, tdO, q0, lane);
    mma_a_tile<HD>(dk, da, tQ, q0, lane);
  }
  store_rows<HD>((bf16*)gd.d_k + koff, a.k_rs, key0, Lk, dk, a.scale, a.scale, lane);
  store_rows<HD>((bf16*)gd.d_v + voff, a.v_rs, key0, Lk, dv, 1.f, 1.f, lane);
  }   // key-block loop
}

template <typename K>
int set_smem_attr(K kern, size_t bytes) {
  SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SC_OK;
}

bool aligned_for_mma(const sc_attn_desc* a) {
  auto ok = [](const void* p, long bs, long rs) { return ((uintptr_t)p & 15) == 0 && bs % 8 == 0 && rs % 8 == 0; };
  return ok(a->q, a->q_bs, a->q_rs) && ok(a->k, a->k_bs, a->k_rs) && ok(a->v, a->v_bs, a->v_rs) && ok(a->o, a->o_bs, a->o_rs);
}

}  // namespace

extern void sc_count_kernel(int kind, int n);

// Whether the tensor-core kernels cover this problem (otherwise the generic fp32 kernels run).
bool sc_attn_mma_supported(const sc_attn_desc* a) {
  return a->dtype == SC_BF16 && (a->hd == 64 || a->hd == 48 || a->hd == 32) && a->Lq >= 1 && a->Lk >= 1 &&
         a->Lq <= 1024 && a->Lk <= 1024 && a->B <= 65535 && a->H <= 65535 && aligned_for_mma(a);
}

#define SC_ATT_DISPATCH(CALL)                          \
  if (a->hd == 64) { if (a->causal) { CALL(64, true) } else { CALL(64, false) } }   \
  else if (a->hd == 48) { if (a->causal) { CALL(48, true) } else { CALL(48, false) } } \
  else { if (a->causal) { CALL(32, true) } else { CALL(32, false) } }

// threads per CTA: every warp loops over 16-row blocks; pick the warp count that balances the passes
static int att_threads(int rows) {
  const int nblocks = (rows + 15) / 16;
  const int passes = (nblocks + AW - 1) / AW;
  const int t = ((nblocks + passes - 1) / passes) * 32;
  // never fewer than four warps: with 8 centre queries a CTA had ONE warp to stage its 52 KB of K / V (cp.async), and an
  // SM held four such warps -- the call was load-latency bound; the extra warps only stage, then leave the row loop
  return t < 128 ? 128 : t;
}

int sc_attention_fwd_mma(const sc_attn_desc* a, cudaStream_t st) {
  const int lk_pad = (a->Lk + 15) & ~15;
  const size_t smem = (size_t)2 * lk_pad * 128;
  dim3 grid(1, a->H, a->B);
  const int AT_ = att_threads(a->Lq);
  sc_count_kernel(SC_K_ATTN_MMA, 1);
#define CALL(HD_, C_)                                                                  \
  {                                                                                    \
    int rc = set_smem_attr(attn_fwd_mma_kernel<HD_, C_>, smem);                        \
    if (rc) return rc;                                                                 \
    attn_fwd_mma_kernel<HD_, C_><<<grid, AT_, smem, st>>>(*a, lk_pad);                  \
  }
  SC_ATT_DISPATCH(CALL)
#undef CALL
  SC_LAUNCH_CHECK();
  return SC_OK;
}

// delta: fp32 scratch [B, H, Lq] provided by the caller through the lse-sized workspace below
int sc_attention_bwd_mma(const sc_attn_bwd_desc* g, float* delta, cudaStream_t st) {
  const sc_attn_desc* a = &g->fwd;
  const int lk_pad = (a->Lk + 15) & ~15, lq_pad = (a->Lq + 15) & ~15;
  const size_t smem_q = (size_t)2 * lk_pad * 128;
  const size_t smem_kv = (size_t)2 * lq_pad * 128 + 2 * lq_pad * sizeof(float);
  dim3 grid_q(1, a->H, a->B), grid_kv(1, a->H, a->B);
  const int tq = att_threads(a->Lq), tkv = att_threads(a->Lk);
  sc_count_kernel(SC_K_ATTN_MMA, 2);
#define CALL(HD_, C_)                                                                          \
  {                                                                                            \
    int rc = set_smem_attr(attn_bwd_dq_mma_kernel<HD_, C_>, smem_q);                           \
    if (rc) return rc;                                                                         \
    rc = set_smem_attr(attn_bwd_dkv_mma_kernel<HD_, C_>, smem_kv);                             \
    if (rc) return rc;                                                                         \
    attn_bwd_dq_mma_kernel<HD_, C_><<<grid_q, tq, smem_q, st>>>(*g, delta, lk_pad);            \
    attn_bwd_dkv_mma_kernel<HD_, C_><<<grid_kv, tkv, smem_kv, st>>>(*g, delta, lq_pad);         \
  }
  SC_ATT_DISPATCH(CALL)
#undef CALL
  SC_LAUNCH_CHECK();
  return SC_OK;
}
