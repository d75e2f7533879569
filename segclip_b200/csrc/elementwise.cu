// HBM-bound glue kernels of the SegCLIP hot path: activation backward, bias gradients, parameter
// shadow casts, patch extraction, embedding, row gather/scatter, MAE masking, pooling.
// All are one-pass, coalesced along the feature dimension, fp32 math.
#include <stdlib.h>

#include "common.cuh"

extern void sc_count_launch(int n);

namespace {

// ---------------------------------------------------------------- activation backward / convert / scale
__global__ void act_bwd_kernel(const void* __restrict__ dy, int dy_dt, const void* __restrict__ pre, int pre_dt,
                               void* __restrict__ dx, int dx_dt, long n, int act) {
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  for (int j = 0; j < 4 && i + j < n; ++j)
    st_any(dx, i + j, dx_dt, ld_any(dy, i + j, dy_dt) * act_grad(ld_any(pre, i + j, pre_dt), act));
}
// dst = src * (scale ? *scale : 1)
__global__ void convert_kernel(const void* __restrict__ src, int sdt, void* __restrict__ dst, int ddt, long n,
                               const float* __restrict__ scale) {
  const float s = scale ? *scale : 1.0f;
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  for (int j = 0; j < 4 && i + j < n; ++j) st_any(dst, i + j, ddt, ld_any(src, i + j, sdt) * s);
}

// ---------------------------------------------------------------- column sums (bias gradients)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long ld, long rows, int cols,
                                                      float* __restrict__ out, int rows_per_block) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (col < cols)
    for (long r = r0 + ty; r < r1; r += 8) s += to_f32(x[r * ld + col]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][tx];
    if (col < cols) atomicAdd(out + col, s);
  }
}

// vectorised variant: cols % 8 == 0, ld % 8 == 0; a warp reads 256 consecutive columns of one row (full lines)
template <typename T> SC_DEVINL void load8(const T* p, float (&v)[8]);
template <> SC_DEVINL void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> SC_DEVINL void load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *(const uint4*)p;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*(const __nv_bfloat162*)&w[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, long ld, long rows, int cols,
                                                          float* __restrict__ out, int rows_per_block) {
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < cols) {
    long r = r0 + warp;
    for (; r + 8 < r1; r += 16) {      // two rows in flight
      float a[8], b[8];
      load8<T>(x + r * ld + col, a);
      load8<T>(x + (r + 8) * ld + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += a[i] + b[i];
    }
    for (; r < r1; r += 8) {
      float a[8];
      load8<T>(x + r * ld + col, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += a[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) atomicAdd(out + c, s);
}

// ---------------------------------------------------------------- multi-tensor cast fp32 -> T
__global__ void cast_multi_kernel(const sc_cast_item* __restrict__ items, int n_items, int dst_dtype) {
  // binary search the item that owns this block
  int lo = 0, hi = n_items - 1;
  const long blk = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const sc_cast_item it = items[lo];
  long i = ((blk - it.first_block) * blockDim.x + threadIdx.x) * 4;
  const float* s = (const float*)it.src;
  if (dst_dtype == SC_BF16 && i + 3 < it.n && (((uintptr_t)s | (uintptr_t)it.dst) & 15) == 0) {
    // one 16-byte load, one 8-byte store per thread (the weights of a step: 0.6 GB read, 0.3 GB written)
    const float4 v = *(const float4*)(s + i);
    __nv_bfloat162 lo2 = __floats2bfloat162_rn(v.x, v.y), hi2 = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *(uint32_t*)&lo2;
    u.y = *(uint32_t*)&hi2;
    *(uint2*)((bf16*)it.dst + i) = u;
    return;
  }
  for (int j = 0; j < 4 && i + j < it.n; ++j) st_any(it.dst, i + j, dst_dtype, s[i + j]);
}

// block-diagonal expansion of a grouped 1x1 conv weight [C, C/groups] -> dense [C, C]
__global__ void blockdiag_expand_kernel(const float* __restrict__ w, void* __restrict__ out, int C, int cg, int dtype) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)C * C) return;
  int o = i / C, c = i % C;
  float v = (c / cg == o / cg) ? w[(long)o * cg + (c % cg)] : 0.f;
  st_any(out, i, dtype, v);
}
// gradient: dW[o, j] += dDense[o, (o/cg)*cg + j]
__global__ void blockdiag_reduce_kernel(const float* __restrict__ dense, float* __restrict__ dw, int C, int cg) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)C * cg) return;
  int o = i / cg, j = i % cg;
  dw[i] += dense[(long)o * C + (o / cg) * cg + j];
}

// ---------------------------------------------------------------- patch extraction (im2col for stride == kernel)
template <typename T>
__global__ void im2col_kernel(const float* __restrict__ img, T* __restrict__ out, long ld, const int* __restrict__ patch_idx,
                              int rows_per_img, int grid_w, int p, int res_h, int res_w) {
  const long row = blockIdx.x;
  const int b = row / rows_per_img;
  const int pid = patch_idx ? patch_idx[row] : (int)(row % rows_per_img);
  const int py0 = (pid / grid_w) * p, px0 = (pid % grid_w) * p;
  const int K = 3 * p * p;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int c = k / (p * p), r = (k / p) % p, q = k % p;
    out[row * ld + k] = from_f32<T>(img[(((long)b * 3 + c) * res_h + py0 + r) * res_w + px0 + q]);
  }
}

// V consecutive pixels of a patch row per thread (V = 4 for patch 16: one 16-byte load, one 8- / 16-byte store; V = 2 for patch 14)
template <typename T, int V>
__global__ void __launch_bounds__(1024) im2col_vec_kernel(const float* __restrict__ img, T* __restrict__ out, long ld,
                                                          const int* __restrict__ patch_idx, int rows_per_img, int grid_w, int p,
                                                          int res_h, int res_w, long rows) {
  const int KV = 3 * p * p / V;                      // vectors per row
  const int per_cta = blockDim.x / KV;               // rows per CTA (host: blockDim.x = per_cta * KV)
  const long row = (long)blockIdx.x * per_cta + threadIdx.x / KV;
  if (row >= rows) return;
  const int kv = threadIdx.x % KV;
  const int b = row / rows_per_img;
  const int pid = patch_idx ? patch_idx[row] : (int)(row % rows_per_img);
  const int py0 = (pid / grid_w) * p, px0 = (pid % grid_w) * p;
  const int k = kv * V;
  const int c = k / (p * p), r = (k / p) % p, q = k % p;
  const float* src = img + (((long)b * 3 + c) * res_h + py0 + r) * res_w + px0 + q;
  T* dst = out + row * ld + k;
  if constexpr (V == 4) {
    const float4 x = *(const float4*)src;
    if constexpr (sizeof(T) == 4) {
      *(float4*)dst = x;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
      uint2 u;
      u.x = *(uint32_t*)&lo;
      u.y = *(uint32_t*)&hi;
      *(uint2*)dst = u;
    }
  } else {
    const float2 x = *(const float2*)src;
    if constexpr (sizeof(T) == 4) *(float2*)dst = x;
    else *(__nv_bfloat162*)dst = __floats2bfloat162_rn(x.x, x.y);
  }
}

// ---------------------------------------------------------------- bicubic resize of the positional table (inference)
// VisualTransformer.get_pos_embed (modules/module_clip_vtransformer.py:35-53): F.interpolate(mode='bicubic',
// align_corners=False) of the [g, g, D] patch table to [h, w, D].  Same arithmetic as ATen's upsample_bicubic2d: source
// coordinate scale * (dst + 0.5) - 0.5 (not clamped), cubic convolution with A = -0.75, taps clamped to the border.
SC_DEVINL float cubic1(float x) { return ((-0.75f + 2.f) * x - (-0.75f + 3.f)) * x * x + 1.f; }
SC_DEVINL float cubic2(float x) { return ((-0.75f * x - 5.f * -0.75f) * x + 8.f * -0.75f) * x - 4.f * -0.75f; }
__global__ void bicubic_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int gh, int gw, int h, int w, int D) {
  const int oy = blockIdx.x / w, ox = blockIdx.x % w;
  const float sy = (float)gh / h * (oy + 0.5f) - 0.5f, sx = (float)gw / w * (ox + 0.5f) - 0.5f;
  const int iy = (int)floorf(sy), ix = (int)floorf(sx);
  const float ty = sy - iy, tx = sx - ix;
  const float wy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
  const float wx[4] = {cubic2(tx + 1.f), cubic1(tx), cubic1(1.f - tx), cubic2(2.f - tx)};
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int y = min(max(iy - 1 + i, 0), gh - 1);
      float rowv = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = min(max(ix - 1 + j, 0), gw - 1);
        rowv += wx[j] * src[((long)y * gw + x) * D + d];
      }
      acc += wy[i] * rowv;
    }
    dst[((long)oy * w + ox) * D + d] = acc;
  }
}

// ---------------------------------------------------------------- uint8 image boundary
// out[b,c,y,x] = (img[b,c,y,x] / 255 - mean[c]) / std[c]: the normalisation the reference loaders do on the host in float64
// (dataloaders/rawimage_util.py), moved after the H2D copy so only 1 byte per pixel crosses PCIe.
__global__ void u8_normalize_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, long n, int hw, float m0, float m1,
                                    float m2, float s0, float s1, float s2) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const int c = (int)((i / hw) % 3);
  const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
  if (i + 3 < n && (hw & 3) == 0) {
    const uchar4 u = *(const uchar4*)(img + i);
    *(float4*)(out + i) = make_float4((u.x / 255.0f - m) / s, (u.y / 255.0f - m) / s, (u.z / 255.0f - m) / s, (u.w / 255.0f - m) / s);
  } else {
    for (long j = i; j < i + 4 && j < n; ++j) {
      const int cj = (int)((j / hw) % 3);
      const float mj = cj == 0 ? m0 : (cj == 1 ? m1 : m2), sj = cj == 0 ? s0 : (cj == 1 ? s1 : s2);
      out[j] = (img[j] / 255.0f - mj) / sj;
    }
  }
}

// ---------------------------------------------------------------- text embedding
__global__ void embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                             const float* __restrict__ pos, float* __restrict__ out, int T, int W) {
  const long row = blockIdx.x;
  const long id = ids[row];
  const int t = row % T;
  for (int d = threadIdx.x; d < W; d += blockDim.x) out[row * W + d] = tok[id * W + d] + pos[(long)t * W + d];
}

__global__ void eot_index_kernel(const long long* __restrict__ ids, int* __restrict__ eot_rows, int B, int T) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long best = ids[(long)b * T];
  int arg = 0;
  for (int t = 1; t < T; ++t) {
    long long v = ids[(long)b * T + t];
    if (v > best) { best = v; arg = t; }   // first maximum, like torch.argmax
  }
  eot_rows[b] = b * T + arg;
}

__global__ void gather_rows_kernel(const void* __restrict__ src, int sdt, const int* __restrict__ idx, float* __restrict__ out, int D) {
  const long r = blockIdx.x;
  const long s = idx[r];
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[r * D + d] = ld_any(src, s * D + d, sdt);
}
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ out, int D) {
  const long r = blockIdx.x;
  const long s = idx[r];
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[s * D + d] = src[r * D + d];
}

// out[idx[r] + idx_offset, :] += src[r, :] (fp32 atomics): gradient of an embedding lookup -- token embedding rows by token id
// (int64), positional rows of the masked visual pass by patch index (int32)
template <typename TI>
__global__ void scatter_add_rows_kernel(const void* __restrict__ src, int sdt, const TI* __restrict__ idx, long idx_offset,
                                        float* __restrict__ out, int D) {
  const long r = blockIdx.x;
  const long s = (long)idx[r] + idx_offset;
  for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(out + s * D + d, ld_any(src, r * D + d, sdt));
}

// ---------------------------------------------------------------- MAE random masking by rank
// ids_restore[b,i] = rank of noise[b,i] (noise[b,0] forced to -1); kept tokens have rank < keep.
__global__ void mae_mask_kernel(const float* __restrict__ u, int L1, int keep, int* __restrict__ ids_restore,
                                int* __restrict__ ids_keep, float* __restrict__ mask, int* __restrict__ patch_idx) {
  extern __shared__ float nz[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < L1; i += blockDim.x) nz[i] = (i == 0) ? -1.f : u[(long)b * L1 + i];
  __syncthreads();
  for (int i = threadIdx.x; i < L1; i += blockDim.x) {
    const float v = nz[i];
    int rank = 0;
    for (int j = 0; j < L1; ++j) rank += (nz[j] < v) || (nz[j] == v && j < i);
    ids_restore[(long)b * L1 + i] = rank;
    mask[(long)b * L1 + i] = rank < keep ? 0.f : 1.f;
    if (rank < keep) {
      ids_keep[(long)b * keep + rank] = i;
      if (rank > 0) patch_idx[(long)b * (keep - 1) + rank - 1] = i - 1;   // CLS is always rank 0
    }
  }
}

// ---------------------------------------------------------------- pooling
__global__ void pool_max_kernel(const float* __restrict__ x, float* __restrict__ out, int* __restrict__ arg, int G, int D) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float best = x[((long)b * G) * D + d];
    int a = 0;
    for (int g = 1; g < G; ++g) {
      float v = x[((long)b * G + g) * D + d];
      if (v > best) { best = v; a = g; }
    }
    out[(long)b * D + d] = best;
    arg[(long)b * D + d] = a;
  }
}
__global__ void pool_max_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, float* __restrict__ dx, int G, int D) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int a = arg[(long)b * D + d];
    const float g0 = dout[(long)b * D + d];
    for (int g = 0; g < G; ++g) dx[((long)b * G + g) * D + d] = (g == a) ? g0 : 0.f;
  }
}
// out[b,0,:] = mean_l x[b,l,:]; out[b,1+l,:] = x[b,l,:]
__global__ void mean_cat_kernel(const float* __restrict__ x, void* __restrict__ out, int n, int D, int dtype) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < n; ++l) {
      float v = x[((long)b * n + l) * D + d];
      s += v;
      st_any(out, ((long)b * (n + 1) + 1 + l) * D + d, dtype, v);
    }
    st_any(out, ((long)b * (n + 1)) * D + d, dtype, s / n);
  }
}
__global__ void mean_cat_bwd_kernel(const void* __restrict__ dout, float* __restrict__ dx, int n, int D, int dtype) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float m = ld_any(dout, ((long)b * (n + 1)) * D + d, dtype) / n;
    for (int l = 0; l < n; ++l) dx[((long)b * n + l) * D + d] = ld_any(dout, ((long)b * (n + 1) + 1 + l) * D + d, dtype) + m;
  }
}

}  // namespace

extern "C" {

int sc_act_bwd(const void* dy, int dy_dtype, const void* pre, int pre_dtype, void* dx, int dx_dtype, int64_t n, int act,
               void* stream) {
  SC_CHECK_ARG(dy && pre && dx && n > 0, "sc_act_bwd: bad args");
  sc_count_launch(1);
  act_bwd_kernel<<<ceil_div(n, 4 * 256), 256, 0, (cudaStream_t)stream>>>(dy, dy_dtype, pre, pre_dtype, dx, dx_dtype, n, act);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, const float* scale_dev, void* stream) {
  SC_CHECK_ARG(src && dst && n > 0, "sc_convert: bad args");
  sc_count_launch(1);
  convert_kernel<<<ceil_div(n, 4 * 256), 256, 0, (cudaStream_t)stream>>>(src, src_dtype, dst, dst_dtype, n, scale_dev);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(x && out && rows > 0 && cols > 0, "sc_colsum: bad args");
  static const bool force_scalar = getenv("SC_COLSUM_SCALAR") != nullptr;
  if (!force_scalar && cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    const int gx = ceil_div(cols, 256);
    int gy = (int)((4L * sc_num_sms() + gx - 1) / gx);
    if (gy > ceil_div(rows, 32)) gy = ceil_div(rows, 32);
    if (gy < 1) gy = 1;
    const int rpb = ceil_div(rows, gy);
    gy = ceil_div(rows, rpb);
    sc_count_launch(1);
    if (dtype == SC_F32) colsum_vec_kernel<float><<<dim3(gx, gy), 256, 0, st>>>((const float*)x, ld, rows, cols, out, rpb);
    else colsum_vec_kernel<bf16><<<dim3(gx, gy), 256, 0, st>>>((const bf16*)x, ld, rows, cols, out, rpb);
    SC_LAUNCH_CHECK();
    return SC_OK;
  }
  const int gx = ceil_div(cols, 32);
  int gy = (int)((4L * sc_num_sms() + gx - 1) / gx);
  if (gy > ceil_div(rows, 64)) gy = ceil_div(rows, 64);
  if (gy < 1) gy = 1;
  const int rpb = ceil_div(rows, gy);
  gy = ceil_div(rows, rpb);
  sc_count_launch(1);
  if (dtype == SC_F32) colsum_kernel<float><<<dim3(gx, gy), 256, 0, st>>>((const float*)x, ld, rows, cols, out, rpb);
  else colsum_kernel<bf16><<<dim3(gx, gy), 256, 0, st>>>((const bf16*)x, ld, rows, cols, out, rpb);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_cast_multi(const sc_cast_item* items_dev, int n_items, int64_t total_blocks, int dst_dtype, void* stream) {
  SC_CHECK_ARG(items_dev && n_items > 0 && total_blocks > 0, "sc_cast_multi: bad args");
  sc_count_launch(1);
  cast_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, dst_dtype);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_blockdiag_expand(const float* w, void* dense, int C, int groups, int dtype, void* stream) {
  SC_CHECK_ARG(w && dense && C % groups == 0, "sc_blockdiag_expand: bad args");
  sc_count_launch(1);
  blockdiag_expand_kernel<<<ceil_div((long)C * C, 256), 256, 0, (cudaStream_t)stream>>>(w, dense, C, C / groups, dtype);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_blockdiag_reduce(const float* dense_grad, float* dw, int C, int groups, void* stream) {
  SC_CHECK_ARG(dense_grad && dw && C % groups == 0, "sc_blockdiag_reduce: bad args");
  sc_count_launch(1);
  blockdiag_reduce_kernel<<<ceil_div((long)C * (C / groups), 256), 256, 0, (cudaStream_t)stream>>>(dense_grad, dw, C, C / groups);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_im2col(const float* image, void* out, int out_dtype, int64_t ld, const int32_t* patch_idx, int64_t rows,
              int rows_per_img, int grid_h, int grid_w, int patch, void* stream) {
  SC_CHECK_ARG(image && out && rows > 0 && ld >= 3 * patch * patch && grid_h > 0 && grid_w > 0, "sc_im2col: bad args");
  sc_count_launch(1);
  {
    const int res_w = grid_w * patch, es = out_dtype == SC_F32 ? 4 : 2;
    const int V = (patch % 4 == 0 && res_w % 4 == 0 && ld % 4 == 0) ? 4 : ((patch % 2 == 0 && res_w % 2 == 0 && ld % 2 == 0) ? 2 : 1);
    const int KV = 3 * patch * patch / (V > 1 ? V : 1);
    if (V > 1 && KV <= 1024 && ((uintptr_t)image & 15) == 0 && ((uintptr_t)out & 15) == 0 && (ld * es) % (V * es) == 0) {
      const int per_cta = KV <= 256 ? 256 / KV : 1, threads = per_cta * KV;
      const unsigned grid = (unsigned)((rows + per_cta - 1) / per_cta);
#define SC_I2C(T_, V_) im2col_vec_kernel<T_, V_><<<grid, threads, 0, (cudaStream_t)stream>>>(image, (T_*)out, ld, patch_idx, rows_per_img, grid_w, patch, grid_h * patch, res_w, rows)
      if (out_dtype == SC_F32) { if (V == 4) SC_I2C(float, 4); else SC_I2C(float, 2); }
      else { if (V == 4) SC_I2C(bf16, 4); else SC_I2C(bf16, 2); }
#undef SC_I2C
      SC_LAUNCH_CHECK();
      return SC_OK;
    }
  }
  if (out_dtype == SC_F32)
    im2col_kernel<float><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(image, (float*)out, ld, patch_idx, rows_per_img, grid_w, patch,
                                                                          grid_h * patch, grid_w * patch);
  else
    im2col_kernel<bf16><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(image, (bf16*)out, ld, patch_idx, rows_per_img, grid_w, patch,
                                                                         grid_h * patch, grid_w * patch);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_bicubic_resize(const float* src, float* dst, int src_h, int src_w, int dst_h, int dst_w, int D, void* stream) {
  SC_CHECK_ARG(src && dst && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && D > 0, "sc_bicubic_resize: bad args");
  sc_count_launch(1);
  bicubic_resize_kernel<<<dst_h * dst_w, 256, 0, (cudaStream_t)stream>>>(src, dst, src_h, src_w, dst_h, dst_w, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_u8_normalize(const uint8_t* img, float* out, int64_t n, int hw, const float* mean3, const float* std3, void* stream) {
  SC_CHECK_ARG(img && out && n > 0 && hw > 0 && mean3 && std3, "sc_u8_normalize: bad args");
  sc_count_launch(1);
  u8_normalize_kernel<<<ceil_div(n, 4 * 256), 256, 0, (cudaStream_t)stream>>>(img, out, n, hw, mean3[0], mean3[1], mean3[2], std3[0],
                                                                               std3[1], std3[2]);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_text_embed(const int64_t* ids, const float* tok, const float* pos, float* out, int32_t* eot_rows, int B, int T,
                  int W, void* stream) {
  SC_CHECK_ARG(ids && tok && pos && out && eot_rows, "sc_text_embed: null pointer");
  sc_count_launch(2);
  embed_kernel<<<B * T, 128, 0, (cudaStream_t)stream>>>((const long long*)ids, tok, pos, out, T, W);
  eot_index_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>((const long long*)ids, eot_rows, B, T);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_gather_rows(const void* src, int src_dtype, const int32_t* idx, float* out, int64_t rows, int D, void* stream) {
  SC_CHECK_ARG(src && idx && out && rows > 0, "sc_gather_rows: bad args");
  sc_count_launch(1);
  gather_rows_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(src, src_dtype, idx, out, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_scatter_rows(const float* src, const int32_t* idx, float* out, int64_t rows, int D, void* stream) {
  SC_CHECK_ARG(src && idx && out && rows > 0, "sc_scatter_rows: bad args");
  sc_count_launch(1);
  scatter_rows_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(src, idx, out, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_scatter_add_rows(const void* src, int src_dtype, const void* idx, int idx_is_int64, int64_t idx_offset, float* out,
                        int64_t rows, int D, void* stream) {
  SC_CHECK_ARG(src && idx && out && rows > 0 && D > 0, "sc_scatter_add_rows: bad args");
  sc_count_launch(1);
  if (idx_is_int64)
    scatter_add_rows_kernel<long long><<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(src, src_dtype, (const long long*)idx, idx_offset, out, D);
  else
    scatter_add_rows_kernel<int><<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(src, src_dtype, (const int*)idx, idx_offset, out, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_mae_mask(const float* u, int B, int L1, int keep, int32_t* ids_restore, int32_t* ids_keep, float* mask,
                int32_t* patch_idx, void* stream) {
  SC_CHECK_ARG(u && ids_restore && ids_keep && mask && patch_idx && keep >= 1 && keep <= L1, "sc_mae_mask: bad args");
  sc_count_launch(1);
  mae_mask_kernel<<<B, 256, L1 * sizeof(float), (cudaStream_t)stream>>>(u, L1, keep, ids_restore, ids_keep, mask, patch_idx);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_pool_max(const float* x, float* out, int32_t* arg, int B, int G, int D, void* stream) {
  SC_CHECK_ARG(x && out && arg, "sc_pool_max: null pointer");
  sc_count_launch(1);
  pool_max_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, out, arg, G, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_pool_max_bwd(const float* dout, const int32_t* arg, float* dx, int B, int G, int D, void* stream) {
  SC_CHECK_ARG(dout && dx && arg, "sc_pool_max_bwd: null pointer");
  sc_count_launch(1);
  pool_max_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dout, arg, dx, G, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_mean_cat(const float* x, void* out, int out_dtype, int B, int n, int D, void* stream) {
  SC_CHECK_ARG(x && out, "sc_mean_cat: null pointer");
  sc_count_launch(1);
  mean_cat_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, out, n, D, out_dtype);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_mean_cat_bwd(const void* dout, int dout_dtype, float* dx, int B, int n, int D, void* stream) {
  SC_CHECK_ARG(dout && dx, "sc_mean_cat_bwd: null pointer");
  sc_count_launch(1);
  mean_cat_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dout, dx, n, D, dout_dtype);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

}  // extern "C"
