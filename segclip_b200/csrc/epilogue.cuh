// Fused GEMM epilogue shared by the tcgen05 kernel and the fp32 FMA kernel.
// Semantics: see sc_gemm in include/segclip_b200.h.
#pragma once
#include "common.cuh"

struct EpiParams {
  int M, N;
  float alpha;
  const float* bias;
  const float* rowbias;
  long ld_rowbias;
  const int* rowbias_idx;
  int rowbias_mod;
  int act;
  const float* residual;
  long ldr;
  void* C;
  long ldc;
  int c_dtype;
  void* C2;
  int c2_dtype;
  int accumulate;
  int atomic;  // accumulate with atomics (split-K)
};

static inline EpiParams make_epi(const sc_gemm_desc* d) {
  EpiParams p;
  p.M = d->M;
  p.N = d->N;
  p.alpha = d->alpha;
  p.bias = d->bias;
  p.rowbias = d->rowbias;
  p.ld_rowbias = d->ld_rowbias;
  p.rowbias_idx = d->rowbias_idx;
  p.rowbias_mod = d->rowbias_mod > 0 ? d->rowbias_mod : 1;
  p.act = d->act;
  p.residual = d->residual;
  p.ldr = d->ldr;
  p.C = d->C;
  p.ldc = d->ldc;
  p.c_dtype = d->c_dtype;
  p.C2 = d->C2;
  p.c2_dtype = d->c2_dtype;
  p.accumulate = d->accumulate;
  p.atomic = d->split_k > 1;
  return p;
}

// scalar element epilogue (FMA kernel, column tails)
SC_DEVINL void epi_store_scalar(const EpiParams& p, int m, int n, float acc) {
  float v = p.alpha * acc;
  if (p.bias) v += p.bias[n];
  if (p.rowbias) {
    int r = p.rowbias_idx ? p.rowbias_idx[m] : (m % p.rowbias_mod);
    v += p.rowbias[(long)r * p.ld_rowbias + n];
  }
  long off = (long)m * p.ldc + n;
  if (p.C2) st_any(p.C2, off, p.c2_dtype, v);
  v = act_fwd(v, p.act);
  if (p.residual) v += p.residual[(long)m * p.ldr + n];
  if (p.atomic) {
    atomicAdd((float*)p.C + off, v);
  } else if (p.accumulate) {
    ((float*)p.C)[off] += v;
  } else {
    st_any(p.C, off, p.c_dtype, v);
  }
}

// 8 consecutive columns of one row, 16B-aligned everywhere (tcgen05 kernel).
SC_DEVINL void st8(void* base, long off, int dtype, const float* v) {
  if (dtype == SC_F32) {
    float4* p = (float4*)((float*)base + off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *(uint32_t*)&a;
    u.y = *(uint32_t*)&b;
    u.z = *(uint32_t*)&c;
    u.w = *(uint32_t*)&d;
    *(uint4*)((bf16*)base + off) = u;
  }
}

SC_DEVINL void epi_store8(const EpiParams& p, int m, int n, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= p.alpha;
  if (p.bias) {
    float4 b0 = __ldg((const float4*)(p.bias + n));
    float4 b1 = __ldg((const float4*)(p.bias + n + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (p.rowbias) {
    int r = p.rowbias_idx ? p.rowbias_idx[m] : (m % p.rowbias_mod);
    const float* rb = p.rowbias + (long)r * p.ld_rowbias + n;
    float4 b0 = __ldg((const float4*)rb);
    float4 b1 = __ldg((const float4*)(rb + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  long off = (long)m * p.ldc + n;
  if (p.C2) st8(p.C2, off, p.c2_dtype, v);
  if (p.act != SC_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fwd(v[i], p.act);
  }
  if (p.residual) {
    const float* r = p.residual + (long)m * p.ldr + n;
    float4 r0 = *(const float4*)r;
    float4 r1 = *(const float4*)(r + 4);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (p.atomic) {
    float* c = (float*)p.C + off;
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(c + i, v[i]);
  } else if (p.accumulate) {
    float4* c = (float4*)((float*)p.C + off);
    float4 c0 = c[0], c1 = c[1];
    c[0] = make_float4(c0.x + v[0], c0.y + v[1], c0.z + v[2], c0.w + v[3]);
    c[1] = make_float4(c1.x + v[4], c1.y + v[5], c1.z + v[6], c1.w + v[7]);
  } else {
    st8(p.C, off, p.c_dtype, v);
  }
}
