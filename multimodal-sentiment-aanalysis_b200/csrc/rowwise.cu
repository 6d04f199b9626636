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
                   const T* __restrict__ dq_add, T* __restrict__ dq_part, T* __restrict__ dattn_part, T* __restrict__ dgate_pre,
                   float* __restrict__ partials /* [gridDim.x, 2, E] */) {
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
        store_vec<T>(dq_part + base + v * VN, o1);
        store_vec<T>(dattn_part + base + v * VN, o2);
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

__global__ void reduce_partials_kernel(const float* __restrict__ partials, int64_t nblk, int E,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= E) return;
  float a = 0.f, b = 0.f;
  for (int64_t k = 0; k < nblk; ++k) {
    a += partials[(k * 2 + 0) * E + c];
    b += partials[(k * 2 + 1) * E + c];
  }
  dgamma[c] = a;
  dbeta[c] = b;
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
__global__ void modal_concat_fwd_kernel(int64_t B, int E, int S, const T* __restrict__ logits, SlotPtrs sp,
                                        float* __restrict__ w, T* __restrict__ fused) {
  constexpr int VN = VecN<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kRowWarps + warp;
  if (b >= B) return;
  float lg[4], mx = -INFINITY, den = 0.f;
  for (int s = 0; s < S; ++s) { lg[s] = to_f(logits[b * S + s]); mx = fmaxf(mx, lg[s]); }
  for (int s = 0; s < S; ++s) { lg[s] = expf(lg[s] - mx); den += lg[s]; }
  for (int s = 0; s < S; ++s) { lg[s] /= den; if (lane == 0) w[b * S + s] = lg[s]; }
  const int nvec = E / VN;
  for (int s = 0; s < S; ++s) {
    const T* src = reinterpret_cast<const T*>(sp.p[s]) + b * (int64_t)E;
    T* dst = fused + b * (int64_t)S * E + (int64_t)s * E;
    for (int v = lane; v < nvec; v += 32) {
      float x[VN];
      load_vec<T>(src + v * VN, x);
#pragma unroll
      for (int j = 0; j < VN; ++j) x[j] *= lg[s];
      store_vec<T>(dst + v * VN, x);
    }
  }
}

template <typename T>
__global__ void modal_concat_bwd_kernel(int64_t B, int E, int S, const T* __restrict__ dfused,
                                        const float* __restrict__ w, SlotPtrs sp, T* __restrict__ dlogits) {
  constexpr int VN = VecN<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kRowWarps + warp;
  if (b >= B) return;
  const int nvec = E / VN;
  float ws[4], dw[4];
  for (int s = 0; s < S; ++s) {
    ws[s] = w[b * S + s];
    const T* src = reinterpret_cast<const T*>(sp.p[s]) + b * (int64_t)E;
    const T* df = dfused + b * (int64_t)S * E + (int64_t)s * E;
    T* dslot = reinterpret_cast<T*>(sp.d[s]);
    float acc = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float x[VN], d[VN];
      load_vec<T>(src + v * VN, x);
      load_vec<T>(df + v * VN, d);
#pragma unroll
      for (int j = 0; j < VN; ++j) { acc += x[j] * d[j]; d[j] *= ws[s]; }
      if (dslot != nullptr) store_vec<T>(dslot + b * (int64_t)E + v * VN, d);
    }
    dw[s] = warp_sum(acc);
  }
  float dot = 0.f;
  for (int s = 0; s < S; ++s) dot += ws[s] * dw[s];
  if (lane == 0)
    for (int s = 0; s < S; ++s) dlogits[b * S + s] = from_f<T>(ws[s] * (dw[s] - dot));
}

// ------------------------------------------------------------------ activations
template <typename T>
__global__ void act_fwd_kernel(int64_t n, const T* __restrict__ x, int act, T* __restrict__ y) {
  constexpr int VN = VecN<T>::N;
  int64_t nvec = n / VN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[VN];
    load_vec<T>(x + i * VN, v);
#pragma unroll
    for (int j = 0; j < VN; ++j) v[j] = apply_act(v[j], act);
    store_vec<T>(y + i * VN, v);
  }
  for (int64_t k = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    y[k] = from_f<T>(apply_act(to_f(x[k]), act));
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
__global__ void act_bwd_kernel(int64_t n, const T* __restrict__ x, const T* __restrict__ dy, int act, T* __restrict__ dx) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    dx[k] = from_f<T>(to_f(dy[k]) * act_grad(to_f(x[k]), act));
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
                     const void* dq_add, void* dq_part, void* dattn_part, void* dgate_pre,
                     float* dgamma, float* dbeta, float* partials, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(E % 8 == 0 && E <= 1024 && E > 0, "mmsa_gate_ln_bwd: E=%lld must be a multiple of 8 and <= 1024", (long long)E);
  MMSA_REQUIRE(dq_bcast == nullptr || bcast_rows > 0, "mmsa_gate_ln_bwd: dq_bcast needs bcast_rows > 0");
  if (M == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("gate_ln_bwd", s, (double)M * E * (dtype == MMSA_F32 ? 4 : 2) * (dy_rows_per_sample > 0 ? 6.0 : 7.0));
  int64_t nblk = mmsa_gate_ln_bwd_blocks(M);
  MMSA_DISPATCH_DTYPE(dtype, T, (gate_ln_bwd_kernel<T><<<(unsigned)nblk, kRowWarps * 32, 0, s>>>(
      M, (int)E, (const T*)dy, dy_rows_per_sample, (const T*)g, (const T*)q, (const T*)attn, gamma, mean, rstd,
      (const T*)dq_bcast, bcast_rows, (const T*)dq_add, (T*)dq_part, (T*)dattn_part, (T*)dgate_pre, partials)));
  MMSA_LAUNCH_CHECK("gate_ln_bwd_kernel");
  reduce_partials_kernel<<<(unsigned)ceil_div(E, 128), 128, 0, s>>>(partials, nblk, (int)E, dgamma, dbeta);
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
  MMSA_REQUIRE(E % 8 == 0, "mmsa_modal_concat_fwd: E must be a multiple of 8");
  if (B == 0) return MMSA_OK;
  SlotPtrs sp{};
  for (int i = 0; i < S; ++i) sp.p[i] = slots_host[i];
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("modal_concat_fwd", s, (double)B * E * S * 2.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (modal_concat_fwd_kernel<T><<<(unsigned)ceil_div(B, kRowWarps), kRowWarps * 32, 0, s>>>(
      B, (int)E, S, (const T*)logits, sp, w, (T*)fused)));
  MMSA_LAUNCH_CHECK("modal_concat_fwd_kernel");
  return MMSA_OK;
}

int mmsa_modal_concat_bwd(int dtype, int64_t B, int64_t E, int S, const void* dfused, const float* w,
                          const void* const* slots_host, void* const* dslots_host, void* dlogits, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(S >= 1 && S <= 4, "mmsa_modal_concat_bwd: S=%d out of [1,4]", S);
  MMSA_REQUIRE(E % 8 == 0, "mmsa_modal_concat_bwd: E must be a multiple of 8");
  if (B == 0) return MMSA_OK;
  SlotPtrs sp{};
  for (int i = 0; i < S; ++i) { sp.p[i] = slots_host[i]; sp.d[i] = dslots_host[i]; }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("modal_concat_bwd", s, (double)B * E * S * 3.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (modal_concat_bwd_kernel<T><<<(unsigned)ceil_div(B, kRowWarps), kRowWarps * 32, 0, s>>>(
      B, (int)E, S, (const T*)dfused, w, sp, (T*)dlogits)));
  MMSA_LAUNCH_CHECK("modal_concat_bwd_kernel");
  return MMSA_OK;
}

int mmsa_act_fwd(int dtype, int64_t n, const void* x, int act, void* y, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (n == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("act_fwd", s, (double)n * 2.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (act_fwd_kernel<T><<<grid_for(n / 4 + 1, 256), 256, 0, s>>>(n, (const T*)x, act, (T*)y)));
  MMSA_LAUNCH_CHECK("act_fwd_kernel");
  return MMSA_OK;
}

int mmsa_act_bwd(int dtype, int64_t n, const void* x, const void* dy, int act, void* dx, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (n == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("act_bwd", s, (double)n * 3.0 * (dtype == MMSA_F32 ? 4 : 2));
  MMSA_DISPATCH_DTYPE(dtype, T, (act_bwd_kernel<T><<<grid_for(n, 256), 256, 0, s>>>(n, (const T*)x, (const T*)dy, act, (T*)dx)));
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
