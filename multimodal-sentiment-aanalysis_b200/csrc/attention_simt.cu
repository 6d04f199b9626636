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

// Dropout on the attention PROBABILITIES (nn.MultiheadAttention(dropout=0.3) in training mode, as nn.TransformerEncoderLayer
// builds it: MultimodalModel.py:89-95): O = (P o M / (1-p)) V with P the soft-max over keys.  The keep decision of element
// (b, h, i, j) is either read from an explicit uint8 mask [B,H,Lq,Lk] (parity tests) or drawn from Philox4x32-10 at counter
// offset + ((b*H + h)*Lq + i)*Lk + j, so the backward kernels re-draw the same mask instead of storing B*H*Lq*Lk bytes.
// rng_state (device {seed, position}) makes the stream position graph-replay safe (see mmsa_rng_advance).
struct AttnDrop {
  float p;                  // 0: no dropout
  const uint8_t* mask;      // explicit keep mask or null
  uint64_t seed, offset;
  const uint64_t* rng_state;
};

__device__ __forceinline__ uint32_t attn_philox(uint64_t seed, uint64_t ctr) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}
// keep / (1-p) of probability element `idx` (flat [B,H,Lq,Lk] index); 1 when dropout is off
__device__ __forceinline__ float attn_keep_scale(const AttnDrop& d, uint64_t seed, uint64_t base, int64_t idx) {
  if (d.p <= 0.f) return 1.f;
  bool keep;
  if (d.mask != nullptr) keep = d.mask[idx] != 0;
  else keep = ((float)(attn_philox(seed, base + (uint64_t)idx) >> 8) * (1.f / 16777216.f)) >= d.p;
  return keep ? 1.f / (1.f - d.p) : 0.f;
}
#define MMSA_ATTN_DROP_PROLOGUE(d)                                                   \
  uint64_t dseed = (d).seed, dbase = (d).offset;                                     \
  if ((d).rng_state != nullptr) { dseed = (d).rng_state[0]; dbase += (d).rng_state[1]; }

template <typename T, int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_fwd_simt_kernel(int H, int Lq, int Lk, const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                     int64_t ldk, const T* __restrict__ v, int64_t ldv, T* __restrict__ o, int64_t ldo,
                     float* __restrict__ lse, float scale, AttnDrop drop) {
  constexpr int DL = D / 32;
  __shared__ float Ks[kChunk][D + 1];
  __shared__ float Vs[kChunk][D + 1];
  MMSA_ATTN_DROP_PROLOGUE(drop)
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
    l = l * alpha + warp_sum(p);            // the soft-max normaliser sees every key; dropout applies to P afterwards
    if (drop.p > 0.f && j < Lk && row_ok) p *= attn_keep_scale(drop, dseed, dbase, ((int64_t)bh * Lq + i) * Lk + j);
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
                        int64_t lddq, float scale, AttnDrop drop) {
  constexpr int DL = D / 32;
  __shared__ float Ks[kChunk][D + 1];
  __shared__ float Vs[kChunk][D + 1];
  MMSA_ATTN_DROP_PROLOGUE(drop)
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
      if (drop.p > 0.f) dp *= attn_keep_scale(drop, dseed, dbase, ((int64_t)bh * Lq + i) * Lk + j);   // dP = dP' o M/(1-p)
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
                         int64_t lddk, T* __restrict__ dv, int64_t lddv, float scale, AttnDrop drop) {
  constexpr int DL = D / 32;
  MMSA_ATTN_DROP_PROLOGUE(drop)
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
      if (drop.p > 0.f) {
        const float ks = attn_keep_scale(drop, dseed, dbase, ((int64_t)bh * Lq + i) * Lk + j);
        dp *= ks;                       // dP = dP' o M/(1-p)
        ds = p * (dp - Ds[lane]);
        p *= ks;                        // dV accumulates the DROPPED probabilities
      } else {
        ds = p * (dp - Ds[lane]);
      }
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
int attn_fwd_simt_drop(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                       const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, AttnDrop drop, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(Lq, kAttnWarps), (unsigned)(B * H));
  ProfScope prof("attn_fwd_simt", s, (double)sizeof(T) * D * (double)B * H * (2.0 * Lq + 2.0 * Lk));
  attn_fwd_simt_kernel<T, D><<<grid, kAttnWarps * 32, 0, s>>>((int)H, (int)Lq, (int)Lk, (const T*)q, ldq, (const T*)k, ldk,
                                                             (const T*)v, ldv, (T*)o, ldo, lse, 1.f / sqrtf((float)D), drop);
  MMSA_LAUNCH_CHECK("attn_fwd_simt_kernel");
  return MMSA_OK;
}
template <typename T, int D>
int attn_fwd_simt(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                  const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s) {
  return attn_fwd_simt_drop<T, D>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, AttnDrop{0.f, nullptr, 0, 0, nullptr}, s);
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
int attn_bwd_simt_drop(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                       const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                       float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, AttnDrop drop,
                       cudaStream_t s) {
  int rc = attn_delta<T, D>(B, H, Lq, o, ldo, dout, lddo, delta, s);
  if (rc) return rc;
  const float scale = 1.f / sqrtf((float)D);
  dim3 g1((unsigned)ceil_div(Lq, kAttnWarps), (unsigned)(B * H));
  {
  ProfScope prof("attn_bwd_dq_simt", s, (double)sizeof(T) * D * (double)B * H * (3.0 * Lq + 2.0 * Lk));
  attn_bwd_dq_simt_kernel<T, D><<<g1, kAttnWarps * 32, 0, s>>>((int)H, (int)Lq, (int)Lk, (const T*)q, ldq, (const T*)k, ldk,
                                                              (const T*)v, ldv, (const T*)dout, lddo, lse, delta, (T*)dq,
                                                              lddq, scale, drop);
  }
  MMSA_LAUNCH_CHECK("attn_bwd_dq_simt_kernel");
  dim3 g2((unsigned)ceil_div(Lk, kAttnWarps), (unsigned)(B * H));
  ProfScope prof("attn_bwd_dkv_simt", s, (double)sizeof(T) * D * (double)B * H * (2.0 * Lq + 4.0 * Lk));
  attn_bwd_dkv_simt_kernel<T, D><<<g2, kAttnWarps * 32, 0, s>>>((int)H, (int)Lq, (int)Lk, (const T*)q, ldq, (const T*)k,
                                                               ldk, (const T*)v, ldv, (const T*)dout, lddo, lse, delta,
                                                               (T*)dk, lddk, (T*)dv, lddv, scale, drop);
  MMSA_LAUNCH_CHECK("attn_bwd_dkv_simt_kernel");
  return MMSA_OK;
}
template <typename T, int D>
int attn_bwd_simt(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                  const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                  float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, cudaStream_t s) {
  return attn_bwd_simt_drop<T, D>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv,
                                  lddv, AttnDrop{0.f, nullptr, 0, 0, nullptr}, s);
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

using namespace mmsa;

extern "C" {

int mmsa_attn_dropout_fwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const void* q, int64_t ldq,
                          const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, float* lse,
                          float dropout_p, const uint8_t* keep_mask, uint64_t seed, uint64_t offset,
                          const uint64_t* rng_state, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(D == 32 || D == 64, "mmsa_attn_dropout_fwd: head dim %lld not in {32,64}", (long long)D);
  MMSA_REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0, "mmsa_attn_dropout_fwd: bad shape");
  MMSA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "mmsa_attn_dropout_fwd: dropout_p out of [0,1)");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  AttnDrop dr{dropout_p, keep_mask, seed, offset, rng_state};
  if (dtype == MMSA_F32)
    return D == 64 ? attn_fwd_simt_drop<float, 64>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, dr, s)
                   : attn_fwd_simt_drop<float, 32>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, dr, s);
  if (dtype == MMSA_BF16)
    return D == 64 ? attn_fwd_simt_drop<bf16, 64>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, dr, s)
                   : attn_fwd_simt_drop<bf16, 32>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, dr, s);
  set_error("mmsa_attn_dropout_fwd: bad dtype %d", dtype);
  return MMSA_ERR_ARG;
}

int mmsa_attn_dropout_bwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const void* q, int64_t ldq,
                          const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                          const void* dout, int64_t lddo, const float* lse, float* delta, void* dq, int64_t lddq, void* dk,
                          int64_t lddk, void* dv, int64_t lddv, float dropout_p, const uint8_t* keep_mask, uint64_t seed,
                          uint64_t offset, const uint64_t* rng_state, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(D == 32 || D == 64, "mmsa_attn_dropout_bwd: head dim %lld not in {32,64}", (long long)D);
  MMSA_REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0, "mmsa_attn_dropout_bwd: bad shape");
  MMSA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "mmsa_attn_dropout_bwd: dropout_p out of [0,1)");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  AttnDrop dr{dropout_p, keep_mask, seed, offset, rng_state};
#define MMSA_ADB(T_, D_) attn_bwd_simt_drop<T_, D_>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, \
                                                    lddq, dk, lddk, dv, lddv, dr, s)
  if (dtype == MMSA_F32) return D == 64 ? MMSA_ADB(float, 64) : MMSA_ADB(float, 32);
  if (dtype == MMSA_BF16) return D == 64 ? MMSA_ADB(bf16, 64) : MMSA_ADB(bf16, 32);
#undef MMSA_ADB
  set_error("mmsa_attn_dropout_bwd: bad dtype %d", dtype);
  return MMSA_ERR_ARG;
}

}  // extern "C"
