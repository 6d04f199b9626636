// attention_simt.cu -- exact-fp32 multi-head attention core on CUDA cores (any Lq/Lk, D in {32,64}).
// Parity-mode engine, and the engine for the 3-token ME-MHACL fusion (ME-MHACL/model.py:69-73) whose
// tiles are far below a tensor-core tile.  Follows torch's need_weights branch: q is scaled by
// 1/sqrt(D) before the product, softmax over keys, P.V; the head-averaged weights the reference
// discards (MultimodalModel.py:139 `attn_output, _ = ...`) are not produced.  The probability
// matrix is never written: forward keeps the per-row log-sum-exp, backward recomputes P.
#include "common.cuh"

namespace mmsa {

constexpr int kAttnWarps = 8;   // rows per block
constexpr int kChunk = 32;      // keys (or queries) staged per shared-memory chunk

template <typename T, int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_fwd_simt_kernel(int H, int Lq, int Lk, const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                     int64_t ldk, const T* __restrict__ v, int64_t ldv, T* __restrict__ o, int64_t ldo,
                     float* __restrict__ lse, float scale) {
  constexpr int DL = D / 32;
  __shared__ float Ks[kChunk][D + 1];
  __shared__ float Vs[kChunk][D + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int i = blockIdx.x * kAttnWarps + warp;
  const bool row_ok = i < Lq;
  float qr[D];
  if (row_ok) {
    const T* qp = q + ((int64_t)b * Lq + i) * ldq + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) qr[d] = to_f(qp[d]) * scale;
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) qr[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f, acc[DL];
#pragma unroll
  for (int t = 0; t < DL; ++t) acc[t] = 0.f;
  for (int j0 = 0; j0 < Lk; j0 += kChunk) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kChunk * D; idx += blockDim.x) {
      int jj = idx / D, d = idx % D;
      int j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < Lk) {
        kv = to_f(k[((int64_t)b * Lk + j) * ldk + h * D + d]);
        vv = to_f(v[((int64_t)b * Lk + j) * ldv + h * D + d]);
      }
      Ks[jj][d] = kv; Vs[jj][d] = vv;
    }
    __syncthreads();
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < Lk) {
      s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(qr[d], Ks[lane][d], s);
    }
    float mn = fmaxf(m, warp_max(s));
    float p = (j < Lk) ? expf(s - mn) : 0.f;
    float alpha = expf(m - mn);          // m = -inf on the first chunk -> alpha = 0
    l = l * alpha + warp_sum(p);
#pragma unroll
    for (int t = 0; t < DL; ++t) acc[t] *= alpha;
    for (int jj = 0; jj < kChunk; ++jj) {
      float pj = __shfl_sync(0xffffffffu, p, jj);
#pragma unroll
      for (int t = 0; t < DL; ++t) acc[t] = fmaf(pj, Vs[jj][lane + 32 * t], acc[t]);
    }
    m = mn;
  }
  if (row_ok) {
    T* op = o + ((int64_t)b * Lq + i) * ldo + h * D;
#pragma unroll
    for (int t = 0; t < DL; ++t) op[lane + 32 * t] = from_f<T>(acc[t] / l);
    if (lane == 0) lse[((int64_t)b * H + h) * Lq + i] = m + logf(l);
  }
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d].  D/VN lanes per (row, head), one 16-byte vector each, so a
// warp reads whole contiguous row segments of O and dO; the partial dots meet in a sub-warp shuffle tree.
template <typename T, int D>
__global__ void __launch_bounds__(256)
attn_delta_kernel(int64_t rows /*B*Lq*/, int H, int Lq, const T* __restrict__ o, int64_t ldo,
                  const T* __restrict__ dout, int64_t lddo, float* __restrict__ delta, int vec_ok) {
  constexpr int VN = VecN<T>::N;
  constexpr int G = D / VN;                      // lanes per (row, head): 4, 8 or 16
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t item = t / G;
  const int sub = (int)(t % G);
  const bool ok = item < rows * H;
  const int64_t r = ok ? item / H : 0;
  const int h = ok ? (int)(item % H) : 0;
  float s = 0.f;
  if (ok) {
    const T* po = o + r * ldo + h * D + sub * VN;
    const T* pd = dout + r * lddo + h * D + sub * VN;
    if (vec_ok) {
      float a[VN], b[VN];
      load_vec<T>(po, a); load_vec<T>(pd, b);
#pragma unroll
      for (int j = 0; j < VN; ++j) s += a[j] * b[j];
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j) s += to_f(po[j]) * to_f(pd[j]);
    }
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (ok && sub == 0) {
    const int64_t b = r / Lq, i = r % Lq;
    delta[(b * H + h) * Lq + i] = s;
  }
}

// dQ: one warp per query row, loops over key chunks
template <typename T, int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_bwd_dq_simt_kernel(int H, int Lq, int Lk, const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                        int64_t ldk, const T* __restrict__ v, int64_t ldv, const T* __restrict__ dout, int64_t lddo,
                        const float* __restrict__ lse, const float* __restrict__ delta, T* __restrict__ dq,
                        int64_t lddq, float scale) {
  constexpr int DL = D / 32;
  __shared__ float Ks[kChunk][D + 1];
  __shared__ float Vs[kChunk][D + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int i = blockIdx.x * kAttnWarps + warp;
  const bool row_ok = i < Lq;
  float qr[D], dor[D];
  float lse_i = 0.f, del_i = 0.f;
  if (row_ok) {
    const T* qp = q + ((int64_t)b * Lq + i) * ldq + h * D;
    const T* dp = dout + ((int64_t)b * Lq + i) * lddo + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) { qr[d] = to_f(qp[d]) * scale; dor[d] = to_f(dp[d]); }
    lse_i = lse[((int64_t)b * H + h) * Lq + i];
    del_i = delta[((int64_t)b * H + h) * Lq + i];
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) { qr[d] = 0.f; dor[d] = 0.f; }
  }
  float acc[DL];
#pragma unroll
  for (int t = 0; t < DL; ++t) acc[t] = 0.f;
  for (int j0 = 0; j0 < Lk; j0 += kChunk) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kChunk * D; idx += blockDim.x) {
      int jj = idx / D, d = idx % D;
      int j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < Lk) {
        kv = to_f(k[((int64_t)b * Lk + j) * ldk + h * D + d]);
        vv = to_f(v[((int64_t)b * Lk + j) * ldv + h * D + d]);
      }
      Ks[jj][d] = kv; Vs[jj][d] = vv;
    }
    __syncthreads();
    const int j = j0 + lane;
    float ds = 0.f;
    if (j < Lk && row_ok) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) { s = fmaf(qr[d], Ks[lane][d], s); dp = fmaf(dor[d], Vs[lane][d], dp); }
      float p = expf(s - lse_i);
      ds = p * (dp - del_i);
    }
    for (int jj = 0; jj < kChunk; ++jj) {
      float dj = __shfl_sync(0xffffffffu, ds, jj);
#pragma unroll
      for (int t = 0; t < DL; ++t) acc[t] = fmaf(dj, Ks[jj][lane + 32 * t], acc[t]);
    }
  }
  if (row_ok) {
    T* op = dq + ((int64_t)b * Lq + i) * lddq + h * D;
#pragma unroll
    for (int t = 0; t < DL; ++t) op[lane + 32 * t] = from_f<T>(acc[t] * scale);
  }
}

// dK, dV: one warp per key row, loops over query chunks
template <typename T, int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_bwd_dkv_simt_kernel(int H, int Lq, int Lk, const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                         int64_t ldk, const T* __restrict__ v, int64_t ldv, const T* __restrict__ dout, int64_t lddo,
                         const float* __restrict__ lse, const float* __restrict__ delta, T* __restrict__ dk,
                         int64_t lddk, T* __restrict__ dv, int64_t lddv, float scale) {
  constexpr int DL = D / 32;
  __shared__ float Qs[kChunk][D + 1];
  __shared__ float Os[kChunk][D + 1];
  __shared__ float Ls[kChunk], Ds[kChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int j = blockIdx.x * kAttnWarps + warp;
  const bool row_ok = j < Lk;
  float kr[D], vr[D];
  if (row_ok) {
    const T* kp = k + ((int64_t)b * Lk + j) * ldk + h * D;
    const T* vp = v + ((int64_t)b * Lk + j) * ldv + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) { kr[d] = to_f(kp[d]); vr[d] = to_f(vp[d]); }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) { kr[d] = 0.f; vr[d] = 0.f; }
  }
  float dka[DL], dva[DL];
#pragma unroll
  for (int t = 0; t < DL; ++t) { dka[t] = 0.f; dva[t] = 0.f; }
  for (int i0 = 0; i0 < Lq; i0 += kChunk) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kChunk * D; idx += blockDim.x) {
      int ii = idx / D, d = idx % D;
      int i = i0 + ii;
      float qv = 0.f, ov = 0.f;
      if (i < Lq) {
        qv = to_f(q[((int64_t)b * Lq + i) * ldq + h * D + d]) * scale;
        ov = to_f(dout[((int64_t)b * Lq + i) * lddo + h * D + d]);
      }
      Qs[ii][d] = qv; Os[ii][d] = ov;
    }
    if (threadIdx.x < kChunk) {
      int i = i0 + threadIdx.x;
      Ls[threadIdx.x] = i < Lq ? lse[((int64_t)b * H + h) * Lq + i] : 0.f;
      Ds[threadIdx.x] = i < Lq ? delta[((int64_t)b * H + h) * Lq + i] : 0.f;
    }
    __syncthreads();
    const int i = i0 + lane;
    float p = 0.f, ds = 0.f;
    if (i < Lq && row_ok) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) { s = fmaf(Qs[lane][d], kr[d], s); dp = fmaf(Os[lane][d], vr[d], dp); }
      p = expf(s - Ls[lane]);
      ds = p * (dp - Ds[lane]);
    }
    for (int ii = 0; ii < kChunk; ++ii) {
      float pi = __shfl_sync(0xffffffffu, p, ii);
      float di = __shfl_sync(0xffffffffu, ds, ii);
#pragma unroll
      for (int t = 0; t < DL; ++t) {
        dva[t] = fmaf(pi, Os[ii][lane + 32 * t], dva[t]);
        dka[t] = fmaf(di, Qs[ii][lane + 32 * t], dka[t]);   // Qs already carries the 1/sqrt(D) scale
      }
    }
  }
  if (row_ok) {
    T* kp = dk + ((int64_t)b * Lk + j) * lddk + h * D;
    T* vp = dv + ((int64_t)b * Lk + j) * lddv + h * D;
#pragma unroll
    for (int t = 0; t < DL; ++t) { kp[lane + 32 * t] = from_f<T>(dka[t]); vp[lane + 32 * t] = from_f<T>(dva[t]); }
  }
}

template <typename T, int D>
int attn_fwd_simt(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                  const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(Lq, kAttnWarps), (unsigned)(B * H));
  ProfScope prof("attn_fwd_simt", s, (double)sizeof(T) * D * (double)B * H * (2.0 * Lq + 2.0 * Lk));
  attn_fwd_simt_kernel<T, D><<<grid, kAttnWarps * 32, 0, s>>>((int)H, (int)Lq, (int)Lk, (const T*)q, ldq, (const T*)k, ldk,
                                                             (const T*)v, ldv, (T*)o, ldo, lse, 1.f / sqrtf((float)D));
  MMSA_LAUNCH_CHECK("attn_fwd_simt_kernel");
  return MMSA_OK;
}

template <typename T, int D>
int attn_delta(int64_t B, int64_t H, int64_t Lq, const void* o, int64_t ldo, const void* dout, int64_t lddo, float* delta,
               cudaStream_t s) {
  constexpr int VN = VecN<T>::N;
  const int64_t threads = B * Lq * H * (D / VN);
  const int vec_ok = ((uintptr_t)o % 16 == 0) && ((uintptr_t)dout % 16 == 0) && (ldo * sizeof(T)) % 16 == 0 &&
                     (lddo * sizeof(T)) % 16 == 0;
  ProfScope prof("attn_delta", s, (double)sizeof(T) * D * (double)B * H * 2.0 * Lq);
  attn_delta_kernel<T, D><<<(unsigned)ceil_div(threads, 256), 256, 0, s>>>(B * Lq, (int)H, (int)Lq, (const T*)o, ldo,
                                                                          (const T*)dout, lddo, delta, vec_ok);
  MMSA_LAUNCH_CHECK("attn_delta_kernel");
  return MMSA_OK;
}

template <typename T, int D>
int attn_bwd_simt(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                  const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                  float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, cudaStream_t s) {
  int rc = attn_delta<T, D>(B, H, Lq, o, ldo, dout, lddo, delta, s);
  if (rc) return rc;
  const float scale = 1.f / sqrtf((float)D);
  dim3 g1((unsigned)ceil_div(Lq, kAttnWarps), (unsigned)(B * H));
  {
  ProfScope prof("attn_bwd_dq_simt", s, (double)sizeof(T) * D * (double)B * H * (3.0 * Lq + 2.0 * Lk));
  attn_bwd_dq_simt_kernel<T, D><<<g1, kAttnWarps * 32, 0, s>>>((int)H, (int)Lq, (int)Lk, (const T*)q, ldq, (const T*)k, ldk,
                                                              (const T*)v, ldv, (const T*)dout, lddo, lse, delta, (T*)dq,
                                                              lddq, scale);
  }
  MMSA_LAUNCH_CHECK("attn_bwd_dq_simt_kernel");
  dim3 g2((unsigned)ceil_div(Lk, kAttnWarps), (unsigned)(B * H));
  ProfScope prof("attn_bwd_dkv_simt", s, (double)sizeof(T) * D * (double)B * H * (2.0 * Lq + 4.0 * Lk));
  attn_bwd_dkv_simt_kernel<T, D><<<g2, kAttnWarps * 32, 0, s>>>((int)H, (int)Lq, (int)Lk, (const T*)q, ldq, (const T*)k,
                                                               ldk, (const T*)v, ldv, (const T*)dout, lddo, lse, delta,
                                                               (T*)dk, lddk, (T*)dv, lddv, scale);
  MMSA_LAUNCH_CHECK("attn_bwd_dkv_simt_kernel");
  return MMSA_OK;
}

// explicit instantiations used by attention.cu
template int attn_fwd_simt<float, 32>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, void*, int64_t, float*, cudaStream_t);
template int attn_fwd_simt<float, 64>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, void*, int64_t, float*, cudaStream_t);
template int attn_fwd_simt<bf16, 32>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, void*, int64_t, float*, cudaStream_t);
template int attn_fwd_simt<bf16, 64>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, void*, int64_t, float*, cudaStream_t);
template int attn_bwd_simt<float, 32>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*, float*, void*, int64_t, void*, int64_t, void*, int64_t, cudaStream_t);
template int attn_bwd_simt<float, 64>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*, float*, void*, int64_t, void*, int64_t, void*, int64_t, cudaStream_t);
template int attn_bwd_simt<bf16, 32>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*, float*, void*, int64_t, void*, int64_t, void*, int64_t, cudaStream_t);
template int attn_bwd_simt<bf16, 64>(int64_t, int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*, float*, void*, int64_t, void*, int64_t, void*, int64_t, cudaStream_t);
template int attn_delta<bf16, 64>(int64_t, int64_t, int64_t, const void*, int64_t, const void*, int64_t, float*, cudaStream_t);

}  // namespace mmsa
