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
  const void* residual;
  int residual_dtype;
  long ldr;
  void* C;
  long ldc;
  int c_dtype;
  void* C2;
  int c2_dtype;
  int c2_is_act_grad;   // C2 := act'(v) instead of v
  int accumulate;
  int atomic;  // accumulate with atomics (split-K)
  const void* mul_aux;  // optional: v *= act'(mul_aux[m,n]) (fused activation backward), leading dim ldc
  int mul_aux_dtype;
  int mul_aux_act;
  float* colsum_out;
  float* dot_out;       // per-head row dots with the aux tile (see sc_gemm_desc.dot_out)
  int dot_L;
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
  p.residual_dtype = d->residual_dtype;
  p.ldr = d->ldr;
  p.C = d->C;
  p.ldc = d->ldc;
  p.c_dtype = d->c_dtype;
  p.C2 = d->C2;
  p.c2_dtype = d->c2_dtype;
  p.c2_is_act_grad = d->c2_is_act_grad;
  p.accumulate = d->accumulate;
  p.atomic = d->split_k > 1;
  p.mul_aux = d->mul_aux;
  p.mul_aux_dtype = d->mul_aux_dtype;
  p.mul_aux_act = d->mul_aux_act;
  p.colsum_out = d->colsum_out;
  p.dot_out = d->dot_out;
  p.dot_L = d->dot_L;
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
  if (p.C2) st_any(p.C2, off, p.c2_dtype, p.c2_is_act_grad ? act_grad(v, p.act) : v);
  v = act_fwd(v, p.act);
  if (p.mul_aux) v *= act_grad(ld_any(p.mul_aux, off, p.mul_aux_dtype), p.mul_aux_act);
  if (p.residual) v += ld_any(p.residual, (long)m * p.ldr + n, p.residual_dtype);
  if (p.atomic) {
    atomicAdd((float*)p.C + off, v);
  } else if (p.accumulate) {
    ((float*)p.C)[off] += v;
  } else {
    st_any(p.C, off, p.c_dtype, v);
  }
}

// 4 consecutive columns of one row, 16-byte aligned everywhere (tcgen05 kernel, after the smem transpose:
// the 8 lanes of a quarter-warp cover 32 consecutive columns of the same row -> full-line global accesses).
SC_DEVINL void st4(void* base, long off, int dtype, const float4& v) {
  if (dtype == SC_F32) {
    *(float4*)((float*)base + off) = v;
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *(uint32_t*)&a;
    u.y = *(uint32_t*)&b;
    *(uint2*)((bf16*)base + off) = u;
  }
}
SC_DEVINL float4 ld4(const void* base, long off, int dtype) {
  if (dtype == SC_F32) return *(const float4*)((const float*)base + off);
  uint2 u = *(const uint2*)((const bf16*)base + off);
  float2 a = __bfloat1622float2(*(__nv_bfloat162*)&u.x), b = __bfloat1622float2(*(__nv_bfloat162*)&u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

SC_DEVINL void epi_store4(const EpiParams& p, int m, int n, float4 v) {
  v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha;
  if (p.bias) {
    const float4 b = __ldg((const float4*)(p.bias + n));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (p.rowbias) {
    const int r = p.rowbias_idx ? p.rowbias_idx[m] : (m % p.rowbias_mod);
    const float4 b = __ldg((const float4*)(p.rowbias + (long)r * p.ld_rowbias + n));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  const long off = (long)m * p.ldc + n;
  if (p.C2) {
    if (p.c2_is_act_grad) st4(p.C2, off, p.c2_dtype, make_float4(act_grad(v.x, p.act), act_grad(v.y, p.act), act_grad(v.z, p.act), act_grad(v.w, p.act)));
    else st4(p.C2, off, p.c2_dtype, v);
  }
  if (p.act != SC_ACT_NONE) {
    v.x = act_fwd(v.x, p.act); v.y = act_fwd(v.y, p.act); v.z = act_fwd(v.z, p.act); v.w = act_fwd(v.w, p.act);
  }
  if (p.mul_aux) {
    const float4 a = ld4(p.mul_aux, off, p.mul_aux_dtype);
    v.x *= act_grad(a.x, p.mul_aux_act); v.y *= act_grad(a.y, p.mul_aux_act);
    v.z *= act_grad(a.z, p.mul_aux_act); v.w *= act_grad(a.w, p.mul_aux_act);
  }
  if (p.residual) {
    const float4 r = ld4(p.residual, (long)m * p.ldr + n, p.residual_dtype);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  if (p.atomic) {
    float* c = (float*)p.C + off;
    atomicAdd(c, v.x); atomicAdd(c + 1, v.y); atomicAdd(c + 2, v.z); atomicAdd(c + 3, v.w);
  } else if (p.accumulate) {
    float4* c = (float4*)((float*)p.C + off);
    const float4 o = *c;
    *c = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
  } else {
    st4(p.C, off, p.c_dtype, v);
  }
}
