// rowwise.cu -- memory-bound row kernels of the fusion block:
//   sigmoid gate + blend + LayerNorm (fwd/bwd), token pooling, modality softmax + weighted concat,
//   activations, L2 row normalisation.
// All are HBM-bound: one warp owns one row, 16-byte vector loads/stores, warp-shuffle reductions,
// fp32 math regardless of the storage type.  Reference arithmetic: MultimodalModel.py:147-149
// (gate/blend/LN), :76 and :401 (pooling), :171-176,:299-306 (weights+concat), :234-235 (normalize).
#include "common.cuh"

namespace mmsa {

constexpr int kRowWarps = 8;        // rows per 256-thread block
constexpr int kMaxRowFloats = 32;   // floats held per lane  => E <= 1024

// ------------------------------------------------------------------ gate + blend + LayerNorm fwd
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32)
gate_ln_fwd_kernel(int64_t M, int E, const T* __restrict__ gate_pre, const T* __restrict__ q,
                   const T* __restrict__ attn, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, T* __restrict__ g_out,
                   T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int VN = VecN<T>::N;
  constexpr int NV = kMaxRowFloats / VN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + warp;
  if (row >= M) return;
  const int nvec = E / VN;
  const int64_t base = row * (int64_t)E;
  float u[kMaxRowFloats];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int v = lane + 32 * i;
    if (v < nvec) {
      float gp[VN], qv[VN], av[VN], gv[VN];
      load_vec<T>(gate_pre + base + v * VN, gp);
      load_vec<T>(q + base + v * VN, qv);
      load_vec<T>(attn + base + v * VN, av);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float g = round_to<T>(sigmoidf_(gp[j]));
        gv[j] = g;
        float uu = g * qv[j] + (1.f - g) * av[j];
        u[i * VN + j] = uu;
        sum += uu;
      }
      store_vec<T>(g_out + base + v * VN, gv);
    }
  }
  const float mean = warp_sum(sum) / (float)E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < VN; ++j) { float d = u[i * VN + j] - mean; sq += d * d; }
    }
  }
  const float var = warp_sum(sq) / (float)E;
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  if (y != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int v = lane + 32 * i;
      if (v < nvec) {
        float out[VN];
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          int c = v * VN + j;
          out[j] = (u[i * VN + j] - mean) * rstd * gamma[c] + beta[c];
        }
        store_vec<T>(y + base + v * VN, out);
      }
    }
  }
}

// ------------------------------------------------------------------ gate + blend + LayerNorm bwd
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32)
gate_ln_bwd_kernel(int64_t M, int E, const T* __restrict__ dy, int64_t dy_rows_per_sample,
                   const T* __restrict__ g, const T* __restrict__ q, const T* __restrict__ attn,
                   const float* __restrict__ gamma, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const T* __restrict__ dq_bcast, int64_t bcast_rows,
                   const T* __restrict__ dq_add, T* __restrict__ dq_part, T* __restrict__ dattn_part, int64_t ldp,
                   T* __restrict__ dgate_pre, float* __restrict__ partials /* [gridDim.x, 2, E] */) {
  constexpr int VN = VecN<T>::N;
  constexpr int NV = kMaxRowFloats / VN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = E / VN;
  float dgam[kMaxRowFloats], dbet[kMaxRowFloats];
#pragma unroll
  for (int i = 0; i < kMaxRowFloats; ++i) { dgam[i] = 0.f; dbet[i] = 0.f; }
  const float dy_scale = dy_rows_per_sample > 0 ? 1.f / (float)dy_rows_per_sample : 1.f;
  const float bc_scale = bcast_rows > 0 ? 1.f / (float)bcast_rows : 0.f;

  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < M; row += (int64_t)gridDim.x * kRowWarps) {
    const int64_t base = row * (int64_t)E;
    const int64_t dybase = (dy_rows_per_sample > 0 ? row / dy_rows_per_sample : row) * (int64_t)E;
    const float mu = mean[row], rs = rstd[row];
    float xh[kMaxRowFloats], dyg[kMaxRowFloats];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int v = lane + 32 * i;
      if (v < nvec) {
        float gv[VN], qv[VN], av[VN], dyv[VN];
        load_vec<T>(g + base + v * VN, gv);
        load_vec<T>(q + base + v * VN, qv);
        load_vec<T>(attn + base + v * VN, av);
        load_vec<T>(dy + dybase + v * VN, dyv);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          float uu = gv[j] * qv[j] + (1.f - gv[j]) * av[j];
          float x = (uu - mu) * rs;
          float d = dyv[j] * dy_scale;
          float dg_ = d * gamma[v * VN + j];
          xh[i * VN + j] = x;
          dyg[i * VN + j] = dg_;
          c1 += dg_;
          c2 += dg_ * x;
          dgam[i * VN + j] += d * x;
          dbet[i * VN + j] += d;
        }
      }
    }
    c1 = warp_sum(c1) / (float)E;
    c2 = warp_sum(c2) / (float)E;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int v = lane + 32 * i;
      if (v < nvec) {
        float gv[VN], qv[VN], av[VN], o1[VN], o2[VN], o3[VN];
        load_vec<T>(g + base + v * VN, gv);
        load_vec<T>(q + base + v * VN, qv);
        load_vec<T>(attn + base + v * VN, av);
        float bc[VN];
        if (dq_bcast != nullptr) {
          load_vec<T>(dq_bcast + (row / bcast_rows) * (int64_t)E + v * VN, bc);
        } else {
#pragma unroll
          for (int j = 0; j < VN; ++j) bc[j] = 0.f;
        }
        if (dq_add != nullptr) {
          float ad[VN];
          load_vec<T>(dq_add + base + v * VN, ad);
#pragma unroll
          for (int j = 0; j < VN; ++j) bc[j] = bc[j] * bc_scale + ad[j];
        } else {
#pragma unroll
          for (int j = 0; j < VN; ++j) bc[j] *= bc_scale;
        }
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          float du = rs * (dyg[i * VN + j] - c1 - xh[i * VN + j] * c2);
          o1[j] = du * gv[j] + bc[j];
          o2[j] = du * (1.f - gv[j]);
          o3[j] = du * (qv[j] - av[j]) * gv[j] * (1.f - gv[j]);
        }
        store_vec<T>(dq_part + row * ldp + v * VN, o1);          // ldp: row stride of the two partial-gradient outputs (they
        store_vec<T>(dattn_part + row * ldp + v * VN, o2);       // may be the two column halves of one [M, 2E] buffer)
        store_vec<T>(dgate_pre + base + v * VN, o3);
      }
    }
  }
  // cross-warp reduction of dgamma/dbeta partials through shared memory
  __shared__ float red[kRowWarps][2][32];
#pragma unroll
  for (int i = 0; i < kMaxRowFloats; ++i) {
    // element i of lane `lane` is column ((lane + 32*(i/VN))*VN + i%VN)
    __syncthreads();
    red[warp][0][lane] = dgam[i];
    red[warp][1][lane] = dbet[i];
    __syncthreads();
    if (warp == 0) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) { a += red[w][0][lane]; b += red[w][1][lane]; }
      int v = lane + 32 * (i / VN);
      int c = v * VN + (i % VN);
      if (v < nvec) {
        partials[((int64_t)blockIdx.x * 2 + 0) * E + c] = a;
        partials[((int64_t)blockIdx.x * 2 + 1) * E + c] = b;
      }
    }
  }
}

// ------------------------------------------------------------------ residual add + LayerNorm (post-norm encoder layer)
// y = LayerNorm(x + r) (r may be null: plain LayerNorm), nn.TransformerEncoderLayer's norm1(x + sa(x)) / norm2(x + ff(x))
// and Subnetwork.norm (MultimodalModel.py:83-105).  One warp per row, the row in registers, fp32 math.
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32)
add_ln_fwd_kernel(int64_t M, int E, const T* __restrict__ x, const T* __restrict__ r, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, T* __restrict__ y, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out) {
  constexpr int VN = VecN<T>::N;
  constexpr int NV = kMaxRowFloats / VN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + warp;
  if (row >= M) return;
  const int nvec = E / VN;
  const int64_t base = row * (int64_t)E;
  float u[kMaxRowFloats];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float xv[VN], rv[VN];
      load_vec<T>(x + base + v * VN, xv);
      if (r != nullptr) load_vec<T>(r + base + v * VN, rv);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const float uu = r != nullptr ? round_to<T>(xv[j] + rv[j]) : xv[j];     // torch materialises x + r in the storage type
        u[i * VN + j] = uu;
        sum += uu;
      }
    }
  }
  const float mean = warp_sum(sum) / (float)E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < VN; ++j) { const float d = u[i * VN + j] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)E + eps);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float out[VN];
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const int c = v * VN + j;
        out[j] = (u[i * VN + j] - mean) * rstd * gamma[c] + beta[c];
      }
      store_vec<T>(y + base + v * VN, out);
    }
  }
}

// du = dL/d(x + r) (the same tensor is the gradient of both summands); dgamma / dbeta partials per block
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32)
add_ln_bwd_kernel(int64_t M, int E, const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ r,
                  const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                  T* __restrict__ du_out, float* __restrict__ partials /* [gridDim.x, 2, E] */) {
  constexpr int VN = VecN<T>::N;
  constexpr int NV = kMaxRowFloats / VN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = E / VN;
  float dgam[kMaxRowFloats], dbet[kMaxRowFloats];
#pragma unroll
  for (int i = 0; i < kMaxRowFloats; ++i) { dgam[i] = 0.f; dbet[i] = 0.f; }
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < M; row += (int64_t)gridDim.x * kRowWarps) {
    const int64_t base = row * (int64_t)E;
    const float mu = mean[row], rs = rstd[row];
    float xh[kMaxRowFloats], dyg[kMaxRowFloats];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float xv[VN], rv[VN], dyv[VN];
        load_vec<T>(x + base + v * VN, xv);
        if (r != nullptr) load_vec<T>(r + base + v * VN, rv);
        load_vec<T>(dy + base + v * VN, dyv);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float uu = r != nullptr ? round_to<T>(xv[j] + rv[j]) : xv[j];
          const float xn = (uu - mu) * rs;
          const float dg_ = dyv[j] * gamma[v * VN + j];
          xh[i * VN + j] = xn;
          dyg[i * VN + j] = dg_;
          c1 += dg_;
          c2 += dg_ * xn;
          dgam[i * VN + j] += dyv[j] * xn;
          dbet[i * VN + j] += dyv[j];
        }
      }
    }
    c1 = warp_sum(c1) / (float)E;
    c2 = warp_sum(c2) / (float)E;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[VN];
#pragma unroll
        for (int j = 0; j < VN; ++j) o[j] = rs * (dyg[i * VN + j] - c1 - xh[i * VN + j] * c2);
        store_vec<T>(du_out + base + v * VN, o);
      }
    }
  }
  __shared__ float red[kRowWarps][2][32];
#pragma unroll
  for (int i = 0; i < kMaxRowFloats; ++i) {
    __syncthreads();
    red[warp][0][lane] = dgam[i];
    red[warp][1][lane] = dbet[i];
    __syncthreads();
    if (warp == 0) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) { a += red[w][0][lane]; b += red[w][1][lane]; }
      const int v = lane + 32 * (i / VN);
      const int c = v * VN + (i % VN);
      if (v < nvec) {
        partials[((int64_t)blockIdx.x * 2 + 0) * E + c] = a;
        partials[((int64_t)blockIdx.x * 2 + 1) * E + c] = b;
      }
    }
  }
}

// y[m, :] = x[m, :] + pe[m % L, :]: the sinusoidal table of PositionalEncoding.forward (MultimodalModel.py:19-20),
// broadcast over the batch.  The backward is the identity.
template <typename T>
__global__ void add_rows_kernel(int64_t M, int E, int64_t L, const T* __restrict__ x, const float* __restrict__ pe, T* __restrict__ y) {
  const int64_t total = M * (int64_t)E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / E; const int c = (int)(i % E);
    y[i] = from_f<T>(to_f(x[i]) + pe[(m % L) * E + c]);
  }
}

// ------------------------------------------------------------------ gate + blend + LayerNorm + token mean-pool
// A block whose output only feeds a token mean-pool (the re-skinned path: t' and v' are pooled right
// away) never needs y[M,E] in HBM: LN(u) stays in fp32 registers and only the pooled sums are written --
// so the pooled features carry no bf16 rounding -- and the query stream q is pooled on the way (the
// raw-feature slot, MultimodalModel.py:299).
// Both directions are HBM-bound streams of 1.5 KB rows.  Each warp owns a ring of shared-memory row
// slots filled by 1-D bulk async copies (cp.async.bulk, the TMA engine; completion on a per-slot
// mbarrier): the rows two steps ahead are already in flight while a row is being reduced, so the
// memory system always sees several KB outstanding per warp instead of one synchronous row.
// F = fp32 values per lane (E <= 32*F).
template <typename T> __device__ __forceinline__ void unpack_vec(const uint4& r, float* out);
template <> __device__ __forceinline__ void unpack_vec<float>(const uint4& r, float* out) {
  out[0] = __uint_as_float(r.x); out[1] = __uint_as_float(r.y); out[2] = __uint_as_float(r.z); out[3] = __uint_as_float(r.w);
}
template <> __device__ __forceinline__ void unpack_vec<bf16>(const uint4& r, float* out) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}

// sigmoid for the gate: exact expf/division in fp32 storage (1e-5 parity mode); ex2.approx + rcp.approx when the
// result is rounded to bf16 anyway (2^-22 relative error against a 2^-9 rounding step)
template <typename T> __device__ __forceinline__ float gate_sigmoid(float x) { return sigmoidf_(x); }
template <> __device__ __forceinline__ float gate_sigmoid<bf16>(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// row slots per warp (prefetch depth): two for bf16 rows; fp32 rows (parity mode) are twice as large and get one
template <typename T> struct PipeStages { static constexpr int N = sizeof(T) == 2 ? 2 : 1; };

__device__ __forceinline__ uint32_t rp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rp_bar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
}
__device__ __forceinline__ void rp_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rp_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void rp_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint4 rp_lds(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// shared-memory layout of the two pooled kernels:
//   [0, 2E floats)                 gamma, beta (fwd) / gamma (bwd)
//   [vec_bytes, +64)               kRowWarps * kPipeStages mbarriers
//   [.., + warps*stages*NS*row)    row slots; after the main loop the same bytes hold the cross-warp reduction
template <typename T, int F>
__global__ void __launch_bounds__(kRowWarps * 32, (sizeof(T) == 2 ? 2 : 1))
gate_ln_pool_fwd_kernel(int L, int E, const T* __restrict__ gate_pre, const T* __restrict__ q,
                        const T* __restrict__ attn, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, T* __restrict__ g_out,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out,
                        float* __restrict__ pooled_y, float* __restrict__ pooled_q, T* __restrict__ pooled_q_lp) {
  constexpr int VN = VecN<T>::N;
  constexpr int NV = F / VN;
  constexpr int NS = 3;                               // streams: gate_pre, q, attn
  constexpr int kPipeStages = PipeStages<T>::N;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  float* gm_s = reinterpret_cast<float*>(sm_raw);
  float* bt_s = gm_s + E;
  const uint32_t row_bytes = (uint32_t)E * sizeof(T);
  const uint32_t bar0 = rp_smem_u32(sm_raw) + 2u * E * 4u;
  uint8_t* slots = sm_raw + 2 * E * 4 + 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int nvec = E / VN;
  const uint32_t wslot = rp_smem_u32(slots) + (uint32_t)warp * kPipeStages * NS * row_bytes;
  const uint32_t wbar = bar0 + (uint32_t)warp * kPipeStages * 8u;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) rp_bar_init(wbar + 8u * s);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < E; c += blockDim.x) { gm_s[c] = gamma[c]; bt_s[c] = beta[c]; }
  __syncthreads();
  const int nrows = warp < L ? (L - warp + kRowWarps - 1) / kRowWarps : 0;
  auto issue = [&](int k) {
    const int s = k % kPipeStages;
    const int64_t base = (b * L + warp + (int64_t)k * kRowWarps) * (int64_t)E;
    const uint32_t dst = wslot + (uint32_t)s * NS * row_bytes, bar = wbar + 8u * s;
    rp_expect(bar, NS * row_bytes);
    rp_copy(dst, gate_pre + base, row_bytes, bar);
    rp_copy(dst + row_bytes, q + base, row_bytes, bar);
    rp_copy(dst + 2 * row_bytes, attn + base, row_bytes, bar);
  };
  if (lane == 0)
    for (int k = 0; k < kPipeStages && k < nrows; ++k) issue(k);
  float ysum[F], qsum[F];
#pragma unroll
  for (int i = 0; i < F; ++i) { ysum[i] = 0.f; qsum[i] = 0.f; }
  for (int k = 0; k < nrows; ++k) {
    const int s = k % kPipeStages;
    const int64_t row = b * L + warp + (int64_t)k * kRowWarps;
    const int64_t base = row * (int64_t)E;
    const uint32_t src = wslot + (uint32_t)s * NS * row_bytes;
    rp_wait(wbar + 8u * s, (uint32_t)(k / kPipeStages) & 1u);
    float u[F];
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};      // four independent partial sums: the row reductions are latency chains
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float gp[VN], qv[VN], av[VN], gv[VN];
        unpack_vec<T>(rp_lds(src + v * 16), gp);
        unpack_vec<T>(rp_lds(src + row_bytes + v * 16), qv);
        unpack_vec<T>(rp_lds(src + 2 * row_bytes + v * 16), av);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float g = round_to<T>(gate_sigmoid<T>(gp[j]));
          gv[j] = g;
          const float uu = g * qv[j] + (1.f - g) * av[j];
          u[i * VN + j] = uu;
          sum4[j & 3] += uu;
          qsum[i * VN + j] += qv[j];
        }
        store_vec<T>(g_out + base + v * VN, gv);
      } else {
#pragma unroll
        for (int j = 0; j < VN; ++j) u[i * VN + j] = 0.f;
      }
    }
    __syncwarp();                                        // every lane has read the slot: refill it
    if (lane == 0 && k + kPipeStages < nrows) issue(k + kPipeStages);
    const float mean = warp_sum((sum4[0] + sum4[1]) + (sum4[2] + sum4[3])) / (float)E;
    float sq4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < VN; ++j) { const float d = u[i * VN + j] - mean; sq4[j & 3] += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum((sq4[0] + sq4[1]) + (sq4[2] + sq4[3])) / (float)E + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
#pragma unroll
        for (int h4 = 0; h4 < VN / 4; ++h4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gm_s + v * VN + h4 * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(bt_s + v * VN + h4 * 4);
          const int o = i * VN + h4 * 4;
          ysum[o + 0] += (u[o + 0] - mean) * rstd * g4.x + b4.x;
          ysum[o + 1] += (u[o + 1] - mean) * rstd * g4.y + b4.y;
          ysum[o + 2] += (u[o + 2] - mean) * rstd * g4.z + b4.z;
          ysum[o + 3] += (u[o + 3] - mean) * rstd * g4.w + b4.w;
        }
      }
    }
  }
  // cross-warp sums in a fixed order (deterministic), then the two pooled rows of this sample;
  // the reduction buffer reuses the row slots (all copies have landed and been consumed by now)
  float* red = reinterpret_cast<float*>(slots);
  const float invL = 1.f / (float)L;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && pooled_q == nullptr && pooled_q_lp == nullptr) break;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < VN; ++j) red[warp * E + v * VN + j] = pass == 0 ? ysum[i * VN + j] : qsum[i * VN + j];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) a += red[w * E + c];
      a *= invL;
      if (pass == 0) pooled_y[b * E + c] = a;
      else {
        if (pooled_q) pooled_q[b * E + c] = a;
        if (pooled_q_lp) pooled_q_lp[b * E + c] = from_f<T>(a);
      }
    }
  }
}

// backward of gate_ln_pool_fwd.  dy of every row of sample b is dpooled_y[b,:] / L; the gradient w.r.t.
// the pooled query stream (dpooled_q[b,:] / L) and an optional per-row extra gradient dq_add are folded
// into dq_part, so no [M,E] broadcast is ever materialised.  Each warp owns a CONTIGUOUS range of rows,
// so the per-sample vectors (dy * gamma) live in registers and change at most every L rows; g, q, attn
// (and dq_add) arrive through the warp's bulk-copy ring.  dgamma[c] = sum_b dy[b,c] * sum_{rows of b} xhat[row,c]
// is accumulated per lane and reduced over blocks by reduce_pool_partials_kernel, which also forms
// dbeta[c] = sum_b dpooled_y[b,c] (every row of a sample carries dpooled_y / L).
template <typename T, int F>
__global__ void __launch_bounds__(kRowWarps * 32, (sizeof(T) == 2 ? 2 : 1))
gate_ln_pool_bwd_kernel(int64_t M, int L, int E, int64_t rows_per_warp, const float* __restrict__ dpooled_y,
                        const float* __restrict__ dpooled_q, const T* __restrict__ dq_add,
                        const T* __restrict__ g, const T* __restrict__ q, const T* __restrict__ attn,
                        const float* __restrict__ gamma, const float* __restrict__ mean,
                        const float* __restrict__ rstd, T* __restrict__ dq_part, T* __restrict__ dattn_part, int64_t ldp,
                        T* __restrict__ dgate_pre, float* __restrict__ partials /* [gridDim.x, 2, E] */) {
  constexpr int VN = VecN<T>::N;
  constexpr int NV = F / VN;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  float* gm_s = reinterpret_cast<float*>(sm_raw);
  const uint32_t row_bytes = (uint32_t)E * sizeof(T);
  const uint32_t bar0 = rp_smem_u32(sm_raw) + 2u * E * 4u;
  uint8_t* slots = sm_raw + 2 * E * 4 + 128;
  const int NS = dq_add != nullptr ? 4 : 3;           // streams: g, q, attn (, dq_add)
  constexpr int kPipeStages = PipeStages<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = E / VN;
  const uint32_t wslot = rp_smem_u32(slots) + (uint32_t)warp * kPipeStages * NS * row_bytes;
  const uint32_t wbar = bar0 + (uint32_t)warp * kPipeStages * 8u;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) rp_bar_init(wbar + 8u * s);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < E; c += blockDim.x) gm_s[c] = gamma[c];
  __syncthreads();
  const int64_t gw = (int64_t)blockIdx.x * kRowWarps + warp;
  const int64_t r0 = gw * rows_per_warp;
  const int64_t r1 = r0 + rows_per_warp < M ? r0 + rows_per_warp : M;
  const int nrows = r1 > r0 ? (int)(r1 - r0) : 0;
  auto issue = [&](int k) {
    const int s = k % kPipeStages;
    const int64_t base = (r0 + k) * (int64_t)E;
    const uint32_t dst = wslot + (uint32_t)s * NS * row_bytes, bar = wbar + 8u * s;
    rp_expect(bar, NS * row_bytes);
    rp_copy(dst, g + base, row_bytes, bar);
    rp_copy(dst + row_bytes, q + base, row_bytes, bar);
    rp_copy(dst + 2 * row_bytes, attn + base, row_bytes, bar);
    if (NS == 4) rp_copy(dst + 3 * row_bytes, dq_add + base, row_bytes, bar);
  };
  if (lane == 0)
    for (int k = 0; k < kPipeStages && k < nrows; ++k) issue(k);
  const float invL = 1.f / (float)L;
  float dgv[F], xs[F], dgam[F];      // dy*gamma of the current sample; sum of xhat over its rows; dgamma partial
#pragma unroll
  for (int i = 0; i < F; ++i) { dgv[i] = 0.f; xs[i] = 0.f; dgam[i] = 0.f; }
  float c1 = 0.f;
  int64_t cur_b = -1;
  auto flush = [&]() {               // fold the finished sample segment into dgamma
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const float4* src = reinterpret_cast<const float4*>(dpooled_y + cur_b * (int64_t)E + v * VN);
#pragma unroll
        for (int h = 0; h < VN / 4; ++h) {
          const float4 d4 = __ldg(src + h);
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { dgam[i * VN + h * 4 + j] += dv[j] * invL * xs[i * VN + h * 4 + j]; xs[i * VN + h * 4 + j] = 0.f; }
        }
      }
    }
  };
  for (int k = 0; k < nrows; ++k) {
    const int s = k % kPipeStages;
    const int64_t row = r0 + k;
    const int64_t base = row * (int64_t)E;
    const int64_t b = row / L;
    if (b != cur_b) {
      if (cur_b >= 0) flush();
      cur_b = b;
      float s0 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
          const float4* src = reinterpret_cast<const float4*>(dpooled_y + b * (int64_t)E + v * VN);
#pragma unroll
          for (int h = 0; h < VN / 4; ++h) {
            const float4 d4 = __ldg(src + h);
            const float4 g4 = *reinterpret_cast<const float4*>(gm_s + v * VN + h * 4);
            dgv[i * VN + h * 4 + 0] = d4.x * invL * g4.x; dgv[i * VN + h * 4 + 1] = d4.y * invL * g4.y;
            dgv[i * VN + h * 4 + 2] = d4.z * invL * g4.z; dgv[i * VN + h * 4 + 3] = d4.w * invL * g4.w;
            s0 += dgv[i * VN + h * 4 + 0] + dgv[i * VN + h * 4 + 1] + dgv[i * VN + h * 4 + 2] + dgv[i * VN + h * 4 + 3];
          }
        }
      }
      c1 = warp_sum(s0) / (float)E;
    }
    const float mu = mean[row], rs = rstd[row];
    const uint32_t src = wslot + (uint32_t)s * NS * row_bytes;
    rp_wait(wbar + 8u * s, (uint32_t)(k / kPipeStages) & 1u);
    float c24[4] = {0.f, 0.f, 0.f, 0.f};       // independent partial sums (latency chain otherwise)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float g8[VN], q8[VN], a8[VN];
        unpack_vec<T>(rp_lds(src + v * 16), g8);
        unpack_vec<T>(rp_lds(src + row_bytes + v * 16), q8);
        unpack_vec<T>(rp_lds(src + 2 * row_bytes + v * 16), a8);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float uu = g8[j] * q8[j] + (1.f - g8[j]) * a8[j];
          c24[j & 3] += dgv[i * VN + j] * ((uu - mu) * rs);
        }
      }
    }
    const float c2 = warp_sum((c24[0] + c24[1]) + (c24[2] + c24[3])) / (float)E;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float g8[VN], q8[VN], a8[VN], o1[VN], o2[VN], o3[VN], ex[VN];
        unpack_vec<T>(rp_lds(src + v * 16), g8);
        unpack_vec<T>(rp_lds(src + row_bytes + v * 16), q8);
        unpack_vec<T>(rp_lds(src + 2 * row_bytes + v * 16), a8);
        if (NS == 4) unpack_vec<T>(rp_lds(src + 3 * row_bytes + v * 16), ex);
        else {
#pragma unroll
          for (int j = 0; j < VN; ++j) ex[j] = 0.f;
        }
        if (dpooled_q != nullptr) {
          const float4* dq4 = reinterpret_cast<const float4*>(dpooled_q + b * (int64_t)E + v * VN);
#pragma unroll
          for (int h = 0; h < VN / 4; ++h) {
            const float4 t4 = __ldg(dq4 + h);
            ex[h * 4 + 0] += t4.x * invL; ex[h * 4 + 1] += t4.y * invL; ex[h * 4 + 2] += t4.z * invL; ex[h * 4 + 3] += t4.w * invL;
          }
        }
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float gg = g8[j];
          const float x = (gg * q8[j] + (1.f - gg) * a8[j] - mu) * rs;
          const float du = rs * (dgv[i * VN + j] - c1 - x * c2);
          xs[i * VN + j] += x;
          o1[j] = du * gg + ex[j];
          o2[j] = du * (1.f - gg);
          o3[j] = du * (q8[j] - a8[j]) * gg * (1.f - gg);
        }
        store_vec<T>(dq_part + row * ldp + v * VN, o1);          // ldp: row stride of the two partial-gradient outputs (they
        store_vec<T>(dattn_part + row * ldp + v * VN, o2);       // may be the two column halves of one [M, 2E] buffer)
        store_vec<T>(dgate_pre + base + v * VN, o3);
      }
    }
    __syncwarp();                                        // every lane has read the slot: refill it
    if (lane == 0 && k + kPipeStages < nrows) issue(k + kPipeStages);
  }
  if (cur_b >= 0) flush();
  // block partial of dgamma: cross-warp sum in warp order through the (now idle) row slots
  float* red = reinterpret_cast<float*>(slots);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < VN; ++j) red[warp * E + v * VN + j] = dgam[i * VN + j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) a += red[w * E + c];
    partials[((int64_t)blockIdx.x * 2 + 0) * E + c] = a;
  }
}

// dgamma[c] = sum_k partials[k,0,c], dbeta[c] = sum_k partials[k,1,c]; 32 columns x 8 k-slices per block,
// k-slices combined in slice order (deterministic).  dbeta_src (pooled form): dbeta[c] = sum_b dbeta_src[b,c].
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int64_t nblk, int E, const float* __restrict__ dbeta_src,
                       int64_t nb, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float a = 0.f, b = 0.f;
  if (c < E) {
    for (int64_t k = ty; k < nblk; k += 8) {
      a += partials[(k * 2 + 0) * E + c];
      if (dbeta_src == nullptr) b += partials[(k * 2 + 1) * E + c];
    }
    if (dbeta_src != nullptr)
      for (int64_t k = ty; k < nb; k += 8) b += dbeta_src[k * E + c];
  }
  red[0][ty][tx] = a; red[1][ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < E) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sa += red[0][k][tx]; sb += red[1][k][tx]; }
    dgamma[c] = sa;
    dbeta[c] = sb;
  }
}

// ------------------------------------------------------------------ token pooling
// block = (E/VN) x RY threads, one sample per block; thread (vx, ry) walks rows ry, ry+RY, ...
template <typename T, bool IS_MAX>
__global__ void pool_fwd_kernel(int64_t L, int E, const T* __restrict__ x, T* __restrict__ y,
                                int32_t* __restrict__ argmax) {
  constexpr int VN = VecN<T>::N;
  extern __shared__ float sm[];                       // [RY][E] (+ [RY][E] ints for max)
  const int nvec = E / VN, RY = blockDim.y;
  const int vx = threadIdx.x, ry = threadIdx.y;
  const int64_t b = blockIdx.x;
  const T* xb = x + b * L * (int64_t)E;
  float acc[VN];
  int arg[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) { acc[j] = IS_MAX ? -INFINITY : 0.f; arg[j] = 0; }
  if (vx < nvec) {
    for (int64_t r = ry; r < L; r += RY) {
      float v[VN];
      load_vec<T>(xb + r * E + vx * VN, v);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        if (IS_MAX) { if (v[j] > acc[j]) { acc[j] = v[j]; arg[j] = (int)r; } }
        else acc[j] += v[j];
      }
    }
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      sm[ry * E + vx * VN + j] = acc[j];
      if (IS_MAX) reinterpret_cast<int*>(sm + RY * E)[ry * E + vx * VN + j] = arg[j];
    }
  }
  __syncthreads();
  if (ry == 0 && vx < nvec) {
    float out[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      int c = vx * VN + j;
      float a = sm[c];
      int ai = IS_MAX ? reinterpret_cast<int*>(sm + RY * E)[c] : 0;
      for (int k = 1; k < RY; ++k) {
        float t = sm[k * E + c];
        if (IS_MAX) {
          int ti = reinterpret_cast<int*>(sm + RY * E)[k * E + c];
          if (t > a || (t == a && ti < ai)) { a = t; ai = ti; }
        } else a += t;
      }
      out[j] = IS_MAX ? a : a / (float)L;
      if (IS_MAX) argmax[b * E + c] = ai;
    }
    store_vec<T>(y + b * E + vx * VN, out);
  }
}

template <typename T, bool IS_MAX>
__global__ void pool_bwd_kernel(int64_t B, int64_t L, int E, const T* __restrict__ dy,
                                const int32_t* __restrict__ argmax, T* __restrict__ dx) {
  constexpr int VN = VecN<T>::N;
  const int nvec = E / VN;
  int64_t total = B * L * nvec;
  const float inv = 1.f / (float)L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int v = (int)(i % nvec);
    int64_t row = i / nvec;
    int64_t b = row / L;
    int l = (int)(row % L);
    float d[VN], o[VN];
    load_vec<T>(dy + b * E + v * VN, d);
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      if (IS_MAX) o[j] = (argmax[b * E + v * VN + j] == l) ? d[j] : 0.f;
      else o[j] = d[j] * inv;
    }
    store_vec<T>(dx + row * E + v * VN, o);
  }
}

// ------------------------------------------------------------------ modality softmax + weighted concat
struct SlotPtrs { const void* p[4]; void* d[4]; };

template <typename T>
__global__ void __launch_bounds__(256)
modal_concat_fwd_kernel(int64_t B, int E, int S, const float* __restrict__ logits, SlotPtrs sp,
                        float* __restrict__ w, T* __restrict__ fused) {
  // one block per sample: every thread forms the S <= 4 softmax weights itself, then the S*E/4 vectors of the
  // weighted concat are spread over the 256 threads (the [B,*] tail is latency-bound: short per-thread loops)
  const int64_t b = blockIdx.x;
  float lg[4], mx = -INFINITY, den = 0.f;
  for (int s = 0; s < S; ++s) { lg[s] = logits[b * S + s]; mx = fmaxf(mx, lg[s]); }
  for (int s = 0; s < S; ++s) { lg[s] = expf(lg[s] - mx); den += lg[s]; }
  for (int s = 0; s < S; ++s) { lg[s] /= den; if (threadIdx.x == 0) w[b * S + s] = lg[s]; }
  const int ev = E / 4;
  for (int v = threadIdx.x; v < S * ev; v += blockDim.x) {
    const int sl = v / ev, c = (v - sl * ev) * 4;
    float x[4];
    load_vec<float>(reinterpret_cast<const float*>(sp.p[sl]) + b * (int64_t)E + c, x);
    float ws = lg[0];
#pragma unroll
    for (int t = 1; t < 4; ++t) if (sl == t) ws = lg[t];
    T* dst = fused + b * (int64_t)S * E + (int64_t)sl * E + c;
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[t] = from_f<T>(x[t] * ws);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
modal_concat_bwd_kernel(int64_t B, int E, int S, const float* __restrict__ dfused,
                        const float* __restrict__ w, SlotPtrs sp, T* __restrict__ dlogits) {
  // one block per sample; dw[s] = <slot_s, dfused_s> reduced warp by warp, then over the 8 warps in warp order
  __shared__ float red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  float ws[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < S; ++s) ws[s] = w[b * S + s];
  const int ev = E / 4;
  for (int v = threadIdx.x; v < S * ev; v += blockDim.x) {
    const int sl = v / ev, c = (v - sl * ev) * 4;
    float x[4], d[4];
    load_vec<float>(reinterpret_cast<const float*>(sp.p[sl]) + b * (int64_t)E + c, x);
    load_vec<float>(dfused + b * (int64_t)S * E + (int64_t)sl * E + c, d);
    float wsl = ws[0];
#pragma unroll
    for (int t = 1; t < 4; ++t) if (sl == t) wsl = ws[t];
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) { dot += x[t] * d[t]; d[t] *= wsl; }
#pragma unroll
    for (int t = 0; t < 4; ++t) if (sl == t) acc[t] += dot;
    float* dslot = reinterpret_cast<float*>(sp.d[sl]);
    if (dslot != nullptr) store_vec<float>(dslot + b * (int64_t)E + c, d);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) { const float r = warp_sum(acc[t]); if (lane == 0) red[t][warp] = r; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dw[4], dot = 0.f;
    for (int s = 0; s < S; ++s) {
      float a = 0.f;
      for (int k = 0; k < 8; ++k) a += red[s][k];
      dw[s] = a;
      dot += ws[s] * a;
    }
    for (int s = 0; s < S; ++s) dlogits[b * S + s] = from_f<T>(ws[s] * (dw[s] - dot));
  }
}


// ------------------------------------------------------------------ modality head: GELU -> Linear(Hd, S) -> softmax -> weighted concat
// One kernel for the rest of `attention_weights` behind its first Linear and the weighted concat that consumes it
// (MultimodalModel.py:173-175, 299-306): h = GELU(h_pre) [Hd <= 256], logits = W2 h + b2 [S <= 4], w = softmax(logits),
// fused = [slot_0 * w_0 | ... | slot_{S-1} * w_{S-1}].  Replaces act_fwd + skinny_fwd + modal_concat_fwd + a cast on the
// (latency-bound) critical path of the [B,*] tail.  One block per sample.  Storage type T: h and W2 are the GEMM-operand
// copies (h is rounded to T before the dot product, as the unfused path does); fused is written in fp32 (autograd
// boundary) and, when fused_lp != null, also in T (the next Linear's operand).
template <typename T>
__global__ void __launch_bounds__(256)
modal_head_fwd_kernel(int64_t B, int E, int S, int Hd, const float* __restrict__ h_pre, const T* __restrict__ w2,
                      const float* __restrict__ b2, SlotPtrs sp, T* __restrict__ hg_out, float* __restrict__ w_out,
                      float* __restrict__ fused, T* __restrict__ fused_lp) {
  __shared__ float hs[256];
  __shared__ float red[4][8];
  __shared__ float wsm[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  float hv = 0.f;
  if ((int)threadIdx.x < Hd) {
    hv = round_to<T>(gelu_erf(h_pre[b * Hd + threadIdx.x]));
    hg_out[b * Hd + threadIdx.x] = from_f<T>(hv);
  }
  hs[threadIdx.x] = hv;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  if ((int)threadIdx.x < Hd)
    for (int s = 0; s < S; ++s) part[s] = hv * to_f(w2[(int64_t)s * Hd + threadIdx.x]);
#pragma unroll
  for (int s = 0; s < 4; ++s) { const float r = warp_sum(part[s]); if (lane == 0) red[s][warp] = r; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lg[4], mx = -INFINITY, den = 0.f;
    for (int s = 0; s < S; ++s) {
      float a = 0.f;
      for (int k = 0; k < 8; ++k) a += red[s][k];
      a += b2 ? b2[s] : 0.f;                 // bias last, as the GEMM epilogue of the unfused path adds it
      lg[s] = a; mx = fmaxf(mx, a);
    }
    for (int s = 0; s < S; ++s) { lg[s] = expf(lg[s] - mx); den += lg[s]; }
    for (int s = 0; s < S; ++s) { lg[s] /= den; wsm[s] = lg[s]; w_out[b * S + s] = lg[s]; }
  }
  __syncthreads();
  const int ev = E / 4;
  for (int v = threadIdx.x; v < S * ev; v += blockDim.x) {
    const int sl = v / ev, c = (v - sl * ev) * 4;
    float x[4];
    load_vec<float>(reinterpret_cast<const float*>(sp.p[sl]) + b * (int64_t)E + c, x);
    const float ws = wsm[sl];
#pragma unroll
    for (int t = 0; t < 4; ++t) x[t] *= ws;
    const int64_t o = b * (int64_t)S * E + (int64_t)sl * E + c;
    store_vec<float>(fused + o, x);
    if (fused_lp != nullptr) {
#pragma unroll
      for (int t = 0; t < 4; ++t) fused_lp[o + t] = from_f<T>(x[t]);
    }
  }
}

// backward of the above: dslot_s = dfused_s * w_s, dw_s = <slot_s, dfused_s>, dlogits = w * (dw - <w, dw>) (rounded to T: the
// operand of the W2 weight gradient), dh_pre = (W2^T dlogits) * GELU'(h_pre) (T: the operand of the first Linear's gradients)
template <typename T>
__global__ void __launch_bounds__(256)
modal_head_bwd_kernel(int64_t B, int E, int S, int Hd, const float* __restrict__ dfused, const float* __restrict__ w,
                      SlotPtrs sp, const float* __restrict__ h_pre, const T* __restrict__ w2, T* __restrict__ dlogits,
                      T* __restrict__ dh_pre) {
  __shared__ float red[4][8];
  __shared__ float dl[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  float ws[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < S; ++s) ws[s] = w[b * S + s];
  const int ev = E / 4;
  for (int v = threadIdx.x; v < S * ev; v += blockDim.x) {
    const int sl = v / ev, c = (v - sl * ev) * 4;
    float x[4], d[4];
    load_vec<float>(reinterpret_cast<const float*>(sp.p[sl]) + b * (int64_t)E + c, x);
    load_vec<float>(dfused + b * (int64_t)S * E + (int64_t)sl * E + c, d);
    float wsl = ws[0];
#pragma unroll
    for (int t = 1; t < 4; ++t) if (sl == t) wsl = ws[t];
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) { dot += x[t] * d[t]; d[t] *= wsl; }
#pragma unroll
    for (int t = 0; t < 4; ++t) if (sl == t) acc[t] += dot;
    float* dslot = reinterpret_cast<float*>(sp.d[sl]);
    if (dslot != nullptr) store_vec<float>(dslot + b * (int64_t)E + c, d);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) { const float r = warp_sum(acc[t]); if (lane == 0) red[t][warp] = r; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dw[4], dot = 0.f;
    for (int s = 0; s < S; ++s) {
      float a = 0.f;
      for (int k = 0; k < 8; ++k) a += red[s][k];
      dw[s] = a;
      dot += ws[s] * a;
    }
    for (int s = 0; s < S; ++s) {
      const float g = round_to<T>(ws[s] * (dw[s] - dot));
      dl[s] = g;
      dlogits[b * S + s] = from_f<T>(g);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < Hd) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += dl[s] * to_f(w2[(int64_t)s * Hd + threadIdx.x]);
    dh_pre[b * Hd + threadIdx.x] = from_f<T>(a * gelu_erf_grad(h_pre[b * Hd + threadIdx.x]));
  }
}

// ------------------------------------------------------------------ activations
template <typename T>
__global__ void act_fwd_kernel(int64_t n, const float* __restrict__ x, int act, T* __restrict__ y) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    y[k] = from_f<T>(apply_act(x[k], act));
}

__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case MMSA_ACT_SIGMOID: { float s = sigmoidf_(x); return s * (1.f - s); }
    case MMSA_ACT_GELU: return gelu_erf_grad(x);
    case MMSA_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}

template <typename T>
__global__ void act_bwd_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ dy, int act, T* __restrict__ dx) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    dx[k] = from_f<T>(dy[k] * act_grad(x[k], act));
}

// ------------------------------------------------------------------ L2 row normalisation
template <typename T>
__global__ void l2norm_fwd_kernel(int64_t B, int E, const T* __restrict__ x, T* __restrict__ y, float* __restrict__ norm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kRowWarps + warp;
  if (b >= B) return;
  float ss = 0.f;
  for (int c = lane; c < E; c += 32) { float v = to_f(x[b * E + c]); ss += v * v; }
  float nrm = sqrtf(warp_sum(ss));
  float den = fmaxf(nrm, 1e-12f);
  if (lane == 0) norm[b] = nrm;
  for (int c = lane; c < E; c += 32) y[b * E + c] = from_f<T>(to_f(x[b * E + c]) / den);
}

// dx = (dy - y * <y, dy>) / max(norm, eps); dy = dy1 (+ dy2), both fp32
template <typename T>
__global__ void l2norm_bwd_kernel(int64_t B, int E, const T* __restrict__ y, const float* __restrict__ norm,
                                  const float* __restrict__ dy1, const float* __restrict__ dy2, T* __restrict__ dx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kRowWarps + warp;
  if (b >= B) return;
  float dot = 0.f;
  for (int c = lane; c < E; c += 32) {
    float d = dy1[b * E + c] + (dy2 ? dy2[b * E + c] : 0.f);
    dot += d * to_f(y[b * E + c]);
  }
  dot = warp_sum(dot);
  float nrm = norm[b];
  // F.normalize clamps the denominator at eps: below it the map is linear x/eps
  bool clamped = nrm < 1e-12f;
  float den = fmaxf(nrm, 1e-12f);
  for (int c = lane; c < E; c += 32) {
    float d = dy1[b * E + c] + (dy2 ? dy2[b * E + c] : 0.f);
    float g = clamped ? d / den : (d - to_f(y[b * E + c]) * dot) / den;
    dx[b * E + c] = from_f<T>(g);
  }
}

}  // namespace mmsa

using namespace mmsa;

static inline unsigned grid_for(int64_t n, int threads, int64_t cap = 148 * 16) {
  int64_t b = ceil_div(n, threads);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

extern "C" {

int mmsa_gate_ln_fwd(int dtype, int64_t M, int64_t E, const void* gate_pre, const void* q, const void* attn,
                     const float* gamma, const float* beta, float eps, void* g_out, void* y,
                     float* mean, float* rstd, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_gate_ln_fwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  if (M == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("gate_ln_fwd", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * (y ? 5.0 : 4.0));
  unsigned grid = (unsigned)ceil_div(M, kRowWarps);
  MMSA_DISPATCH_DTYPE(dtype, T, (gate_ln_fwd_kernel<T><<<grid, kRowWarps * 32, 0, s>>>(
      M, (int)E, (const T*)gate_pre, (const T*)q, (const T*)attn, gamma, beta, eps, (T*)g_out, (T*)y, mean, rstd)));
  MMSA_LAUNCH_CHECK("gate_ln_fwd_kernel");
  return MMSA_OK;
}

int64_t mmsa_gate_ln_bwd_blocks(int64_t M) {
  int64_t b = ceil_div(M, kRowWarps);
  if (b > 148 * 2) b = 148 * 2;
  if (b < 1) b = 1;
  return b;
}

int mmsa_gate_ln_bwd(int dtype, int64_t M, int64_t E, const void* dy, int64_t dy_rows_per_sample,
                     const void* g, const void* q, const void* attn, const float* gamma,
                     const float* mean, const float* rstd, const void* dq_bcast, int64_t bcast_rows,
                     const void* dq_add, void* dq_part, void* dattn_part, int64_t ld_parts, void* dgate_pre,
                     float* dgamma, float* dbeta, float* partials, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_gate_ln_bwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  MMSA_REQUIRE(dq_bcast == nullptr || bcast_rows > 0, "mmsa_gate_ln_bwd: dq_bcast needs bcast_rows > 0");
  if (ld_parts <= 0) ld_parts = E;
  MMSA_REQUIRE(ld_parts >= E && ld_parts % 8 == 0, "mmsa_gate_ln_bwd: ld_parts must be >= E and a multiple of 8");
  if (M == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("gate_ln_bwd", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * (dy_rows_per_sample > 0 ? 6.0 : 7.0));
  int64_t nblk = mmsa_gate_ln_bwd_blocks(M);
  MMSA_DISPATCH_DTYPE(dtype, T, (gate_ln_bwd_kernel<T><<<(unsigned)nblk, kRowWarps * 32, 0, s>>>(
      M, (int)E, (const T*)dy, dy_rows_per_sample, (const T*)g, (const T*)q, (const T*)attn, gamma, mean, rstd,
      (const T*)dq_bcast, bcast_rows, (const T*)dq_add, (T*)dq_part, (T*)dattn_part, ld_parts, (T*)dgate_pre, partials)));
  MMSA_LAUNCH_CHECK("gate_ln_bwd_kernel");
  reduce_partials_kernel<<<(unsigned)ceil_div(E, 32), 256, 0, s>>>(partials, nblk, (int)E, nullptr, 0, dgamma, dbeta);
  MMSA_LAUNCH_CHECK("reduce_partials_kernel");
  return MMSA_OK;
}


int mmsa_add_ln_fwd(int dtype, int64_t M, int64_t E, const void* x, const void* r, const float* gamma, const float* beta,
                    float eps, void* y, float* mean, float* rstd, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_add_ln_fwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  if (M == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("add_ln_fwd", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * (r ? 3.0 : 2.0));
  unsigned grid = (unsigned)ceil_div(M, kRowWarps);
  MMSA_DISPATCH_DTYPE(dtype, T, (add_ln_fwd_kernel<T><<<grid, kRowWarps * 32, 0, s>>>(
      M, (int)E, (const T*)x, (const T*)r, gamma, beta, eps, (T*)y, mean, rstd)));
  MMSA_LAUNCH_CHECK("add_ln_fwd_kernel");
  return MMSA_OK;
}

int mmsa_add_ln_bwd(int dtype, int64_t M, int64_t E, const void* dy, const void* x, const void* r, const float* gamma,
                    const float* mean, const float* rstd, void* du, float* dgamma, float* dbeta, float* partials,
                    void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_add_ln_bwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  if (M == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t nblk = mmsa_gate_ln_bwd_blocks(M);
  {
    ProfScope prof("add_ln_bwd", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * (r ? 4.0 : 3.0));
    MMSA_DISPATCH_DTYPE(dtype, T, (add_ln_bwd_kernel<T><<<(unsigned)nblk, kRowWarps * 32, 0, s>>>(
        M, (int)E, (const T*)dy, (const T*)x, (const T*)r, gamma, mean, rstd, (T*)du, partials)));
  }
  MMSA_LAUNCH_CHECK("add_ln_bwd_kernel");
  reduce_partials_kernel<<<(unsigned)ceil_div(E, 32), 256, 0, s>>>(partials, nblk, (int)E, nullptr, 0, dgamma, dbeta);
  MMSA_LAUNCH_CHECK("reduce_partials_kernel");
  return MMSA_OK;
}

int mmsa_add_rows(int dtype, int64_t M, int64_t E, int64_t L, const void* x, const float* pe, void* y, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(M >= 0 && E > 0 && L > 0 && pe != nullptr, "mmsa_add_rows: bad arguments");
  if (M == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("add_rows", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * 2.0);
  MMSA_DISPATCH_DTYPE(dtype, T, (add_rows_kernel<T><<<grid_for(M * E, 256), 256, 0, s>>>(M, (int)E, L, (const T*)x, pe, (T*)y)));
  MMSA_LAUNCH_CHECK("add_rows_kernel");
  return MMSA_OK;
}

int mmsa_gate_ln_pool_fwd(int dtype, int64_t B, int64_t L, int64_t E, const void* gate_pre, const void* q,
                          const void* attn, const float* gamma, const float* beta, float eps, void* g_out,
                          float* mean, float* rstd, float* pooled_y, float* pooled_q, void* pooled_q_lp,
                          void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_gate_ln_pool_fwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  MMSA_REQUIRE(L > 0 && L < (1 << 30) && pooled_y != nullptr, "mmsa_gate_ln_pool_fwd: bad arguments");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("gate_ln_pool_fwd", s, (double)B * L * E * (dtype == MMSA_F32 ? 4 : 2) * 4.0);
  MMSA_REQUIRE(((uintptr_t)gate_pre | (uintptr_t)q | (uintptr_t)attn | (uintptr_t)g_out) % 16 == 0,
               "mmsa_gate_ln_pool_fwd: operands must be 16-byte aligned");
#define MMSA_GLP_FWD(F_)                                                                                        \
  MMSA_DISPATCH_DTYPE(dtype, T, {                                                                               \
    auto kfn = gate_ln_pool_fwd_kernel<T, F_>;                                                                  \
    size_t slots = (size_t)kRowWarps * PipeStages<T>::N * 3 * E * sizeof(T);                                    \
    size_t red = (size_t)kRowWarps * E * sizeof(float);                                                         \
    size_t smem = (size_t)2 * E * sizeof(float) + 128 + (slots > red ? slots : red);                            \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                          \
    kfn<<<(unsigned)B, kRowWarps * 32, smem, s>>>((int)L, (int)E, (const T*)gate_pre, (const T*)q, (const T*)attn, \
                                                  gamma, beta, eps, (T*)g_out, mean, rstd, pooled_y, pooled_q,   \
                                                  (T*)pooled_q_lp);                                             \
  })
  if (E <= 256) MMSA_GLP_FWD(8);
  else if (E <= 512) MMSA_GLP_FWD(16);
  else if (E <= 768) MMSA_GLP_FWD(24);
  else MMSA_GLP_FWD(32);
#undef MMSA_GLP_FWD
  MMSA_LAUNCH_CHECK("gate_ln_pool_fwd_kernel");
  return MMSA_OK;
}

int mmsa_gate_ln_pool_bwd(int dtype, int64_t B, int64_t L, int64_t E, const float* dpooled_y,
                          const float* dpooled_q, const void* dq_add, const void* g, const void* q,
                          const void* attn, const float* gamma, const float* mean, const float* rstd,
                          void* dq_part, void* dattn_part, int64_t ld_parts, void* dgate_pre, float* dgamma, float* dbeta,
                          float* partials, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_gate_ln_pool_bwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  MMSA_REQUIRE(L > 0 && L < (1 << 30) && dpooled_y != nullptr, "mmsa_gate_ln_pool_bwd: bad arguments");
  if (ld_parts <= 0) ld_parts = E;
  MMSA_REQUIRE(ld_parts >= E && ld_parts % 8 == 0, "mmsa_gate_ln_pool_bwd: ld_parts must be >= E and a multiple of 8");
  MMSA_REQUIRE(((uintptr_t)g | (uintptr_t)q | (uintptr_t)attn | (uintptr_t)dq_add | (uintptr_t)dq_part | (uintptr_t)dattn_part |
                (uintptr_t)dgate_pre | (uintptr_t)dpooled_y | (uintptr_t)dpooled_q) % 16 == 0,
               "mmsa_gate_ln_pool_bwd: operands must be 16-byte aligned");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = B * L;
  int64_t nblk = mmsa_gate_ln_bwd_blocks(M);
  {
    ProfScope prof("gate_ln_pool_bwd", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * (dq_add ? 7.0 : 6.0));
    const int64_t rpw = ceil_div(M, nblk * kRowWarps);
    const int ns = dq_add ? 4 : 3;
#define MMSA_GLP_BWD(F_)                                                                                        \
  MMSA_DISPATCH_DTYPE(dtype, T, {                                                                               \
    auto kfn = gate_ln_pool_bwd_kernel<T, F_>;                                                                  \
    size_t slots = (size_t)kRowWarps * PipeStages<T>::N * ns * E * sizeof(T);                                   \
    size_t red = (size_t)kRowWarps * E * sizeof(float);                                                         \
    size_t smem = (size_t)2 * E * sizeof(float) + 128 + (slots > red ? slots : red);                            \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                          \
    kfn<<<(unsigned)nblk, kRowWarps * 32, smem, s>>>(                                                           \
      M, (int)L, (int)E, rpw, dpooled_y, dpooled_q, (const T*)dq_add, (const T*)g, (const T*)q, (const T*)attn, gamma, \
      mean, rstd, (T*)dq_part, (T*)dattn_part, ld_parts, (T*)dgate_pre, partials);                              \
  })
    if (E <= 256) MMSA_GLP_BWD(8);
    else if (E <= 512) MMSA_GLP_BWD(16);
    else if (E <= 768) MMSA_GLP_BWD(24);
    else MMSA_GLP_BWD(32);
#undef MMSA_GLP_BWD
  }
  MMSA_LAUNCH_CHECK("gate_ln_pool_bwd_kernel");
  reduce_partials_kernel<<<(unsigned)ceil_div(E, 32), 256, 0, s>>>(partials, nblk, (int)E, dpooled_y, B, dgamma, dbeta);
  MMSA_LAUNCH_CHECK("reduce_partials_kernel");
  return MMSA_OK;
}

int mmsa_pool_fwd(int dtype, int64_t B, int64_t L, int64_t E, const void* x, int is_max, void* y,
                  int32_t* argmax, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E > 0 && E <= 2048, "mmsa_pool_fwd: E=%lld must be a multiple of 8 and <= 2048", (long long)E);
  MMSA_REQUIRE(L > 0, "mmsa_pool_fwd: L must be > 0");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("pool_fwd", s, (double)B * (L + 1) * E * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, {
    constexpr int VN = VecN<T>::N;
    int nvec = (int)E / VN;
    int tx = ((nvec + 31) / 32) * 32;
    int ry = 1024 / tx; if (ry > 4) ry = 4; if (ry > L) ry = (int)L; if (ry < 1) ry = 1;
    size_t smem = (size_t)ry * E * sizeof(float) * (is_max ? 2 : 1);
    dim3 block(tx, ry);
    if (is_max) {
      cudaFuncSetAttribute(pool_fwd_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      pool_fwd_kernel<T, true><<<(unsigned)B, block, smem, s>>>(L, (int)E, (const T*)x, (T*)y, argmax);
    } else {
      pool_fwd_kernel<T, false><<<(unsigned)B, block, smem, s>>>(L, (int)E, (const T*)x, (T*)y, argmax);
    }
  });
  MMSA_LAUNCH_CHECK("pool_fwd_kernel");
  return MMSA_OK;
}

int mmsa_pool_bwd(int dtype, int64_t B, int64_t L, int64_t E, const void* dy, int is_max,
                  const int32_t* argmax, void* dx, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E > 0, "mmsa_pool_bwd: E must be a multiple of 8");
  if (B == 0 || L == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("pool_bwd", s, (double)B * (L + 1) * E * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, {
    int64_t total = B * L * (E / VecN<T>::N);
    if (is_max) pool_bwd_kernel<T, true><<<grid_for(total, 256), 256, 0, s>>>(B, L, (int)E, (const T*)dy, argmax, (T*)dx);
    else pool_bwd_kernel<T, false><<<grid_for(total, 256), 256, 0, s>>>(B, L, (int)E, (const T*)dy, argmax, (T*)dx);
  });
  MMSA_LAUNCH_CHECK("pool_bwd_kernel");
  return MMSA_OK;
}

int mmsa_modal_concat_fwd(int dtype, int64_t B, int64_t E, int S, const void* logits,
                          const void* const* slots_host, float* w, void* fused, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(S >= 1 && S <= 4, "mmsa_modal_concat_fwd: S=%d out of [1,4]", S);
  MMSA_REQUIRE(E % 4 == 0, "mmsa_modal_concat_fwd: E must be a multiple of 4");
  if (B == 0) return MMSA_OK;
  SlotPtrs sp{};
  for (int i = 0; i < S; ++i) sp.p[i] = slots_host[i];
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("modal_concat_fwd", s, (double)B * E * S * 2.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (modal_concat_fwd_kernel<T><<<(unsigned)B, 256, 0, s>>>(
      B, (int)E, S, (const float*)logits, sp, w, (T*)fused)));
  MMSA_LAUNCH_CHECK("modal_concat_fwd_kernel");
  return MMSA_OK;
}

int mmsa_modal_concat_bwd(int dtype, int64_t B, int64_t E, int S, const void* dfused, const float* w,
                          const void* const* slots_host, void* const* dslots_host, void* dlogits, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(S >= 1 && S <= 4, "mmsa_modal_concat_bwd: S=%d out of [1,4]", S);
  MMSA_REQUIRE(E % 4 == 0, "mmsa_modal_concat_bwd: E must be a multiple of 4");
  if (B == 0) return MMSA_OK;
  SlotPtrs sp{};
  for (int i = 0; i < S; ++i) { sp.p[i] = slots_host[i]; sp.d[i] = dslots_host[i]; }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("modal_concat_bwd", s, (double)B * E * S * 3.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (modal_concat_bwd_kernel<T><<<(unsigned)B, 256, 0, s>>>(
      B, (int)E, S, (const float*)dfused, w, sp, (T*)dlogits)));
  MMSA_LAUNCH_CHECK("modal_concat_bwd_kernel");
  return MMSA_OK;
}

int mmsa_modal_head_fwd(int dtype, int64_t B, int64_t E, int S, int64_t Hd, const float* h_pre, const void* w2, const float* b2,
                        const void* const* slots_host, void* hg, float* w, float* fused, void* fused_lp, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(S >= 1 && S <= 4, "mmsa_modal_head_fwd: S=%d out of [1,4]", S);
  MMSA_REQUIRE(E % 4 == 0 && Hd >= 1 && Hd <= 256, "mmsa_modal_head_fwd: E must be a multiple of 4 and Hd in [1,256]");
  if (B == 0) return MMSA_OK;
  SlotPtrs sp{};
  for (int i = 0; i < S; ++i) sp.p[i] = slots_host[i];
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("modal_head_fwd", s, (double)B * E * S * (8.0 + (fused_lp ? 2.0 : 0.0)));
  MMSA_DISPATCH_DTYPE(dtype, T, (modal_head_fwd_kernel<T><<<(unsigned)B, 256, 0, s>>>(
      B, (int)E, S, (int)Hd, h_pre, (const T*)w2, b2, sp, (T*)hg, w, fused, (T*)(dtype == MMSA_F32 ? nullptr : fused_lp))));
  MMSA_LAUNCH_CHECK("modal_head_fwd_kernel");
  return MMSA_OK;
}

int mmsa_modal_head_bwd(int dtype, int64_t B, int64_t E, int S, int64_t Hd, const float* dfused, const float* w,
                        const void* const* slots_host, void* const* dslots_host, const float* h_pre, const void* w2,
                        void* dlogits, void* dh_pre, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(S >= 1 && S <= 4, "mmsa_modal_head_bwd: S=%d out of [1,4]", S);
  MMSA_REQUIRE(E % 4 == 0 && Hd >= 1 && Hd <= 256, "mmsa_modal_head_bwd: E must be a multiple of 4 and Hd in [1,256]");
  if (B == 0) return MMSA_OK;
  SlotPtrs sp{};
  for (int i = 0; i < S; ++i) { sp.p[i] = slots_host[i]; sp.d[i] = dslots_host[i]; }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("modal_head_bwd", s, (double)B * E * S * 12.0);
  MMSA_DISPATCH_DTYPE(dtype, T, (modal_head_bwd_kernel<T><<<(unsigned)B, 256, 0, s>>>(
      B, (int)E, S, (int)Hd, dfused, w, sp, h_pre, (const T*)w2, (T*)dlogits, (T*)dh_pre)));
  MMSA_LAUNCH_CHECK("modal_head_bwd_kernel");
  return MMSA_OK;
}

int mmsa_act_fwd(int dtype, int64_t n, const void* x, int act, void* y, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (n == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("act_fwd", s, (double)n * 2.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (act_fwd_kernel<T><<<grid_for(n, 256), 256, 0, s>>>(n, (const float*)x, act, (T*)y)));
  MMSA_LAUNCH_CHECK("act_fwd_kernel");
  return MMSA_OK;
}

int mmsa_act_bwd(int dtype, int64_t n, const void* x, const void* dy, int act, void* dx, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (n == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("act_bwd", s, (double)n * 3.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (act_bwd_kernel<T><<<grid_for(n, 256), 256, 0, s>>>(n, (const float*)x, (const float*)dy, act, (T*)dx)));
  MMSA_LAUNCH_CHECK("act_bwd_kernel");
  return MMSA_OK;
}

int mmsa_l2norm_fwd(int dtype, int64_t B, int64_t E, const void* x, void* y, float* norm, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("l2norm_fwd", s, (double)B * E * 2.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (l2norm_fwd_kernel<T><<<(unsigned)ceil_div(B, kRowWarps), kRowWarps * 32, 0, s>>>(
      B, (int)E, (const T*)x, (T*)y, norm)));
  MMSA_LAUNCH_CHECK("l2norm_fwd_kernel");
  return MMSA_OK;
}

int mmsa_l2norm_bwd(int dtype, int64_t B, int64_t E, const void* y, const float* norm, const float* dy1,
                    const float* dy2, void* dx, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("l2norm_bwd", s, (double)B * E * ((dtype == MMSA_F32 ? 4 : 2) * 2.0 + (dy2 ? 8.0 : 4.0)));
  MMSA_DISPATCH_DTYPE(dtype, T, (l2norm_bwd_kernel<T><<<(unsigned)ceil_div(B, kRowWarps), kRowWarps * 32, 0, s>>>(
      B, (int)E, (const T*)y, norm, dy1, dy2, (T*)dx)));
  MMSA_LAUNCH_CHECK("l2norm_bwd_kernel");
  return MMSA_OK;
}

}  // extern "C"
