// common.cuh -- shared device/host helpers for the sm_100a hot-path kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mmsa.h"

namespace mmsa {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool device_ok();

// Optional per-launch timing (mmsa_prof_enable): brackets one kernel launch with CUDA events on the
// launching stream and books its duration and algorithmic work (FLOPs or bytes) under `name`.
// Costs one predictable branch when disabled.  Not usable during CUDA-graph capture.
extern int g_prof_on;
void prof_begin(const char* name, cudaStream_t s, double work);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s_;
  bool on_;
  ProfScope(const char* name, cudaStream_t s, double work = 0.0) : s_(s), on_(g_prof_on != 0) {
    if (on_) prof_begin(name, s, work);
  }
  ~ProfScope() {
    if (on_) prof_end(s_);
  }
};

#define MMSA_REQUIRE(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      mmsa::set_error(__VA_ARGS__);            \
      return MMSA_ERR_ARG;                     \
    }                                          \
  } while (0)

#define MMSA_REQUIRE_DEVICE()                                                        \
  do {                                                                               \
    if (!mmsa::device_ok()) {                                                        \
      mmsa::set_error("mmsa: current CUDA device is not sm_100 (B200); no fallback"); \
      return MMSA_ERR_DEVICE;                                                        \
    }                                                                                \
  } while (0)

#define MMSA_LAUNCH_CHECK(name)                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      mmsa::set_error("mmsa: launch of %s failed: %s", name, cudaGetErrorString(e__));   \
      return MMSA_ERR_CUDA;                                                              \
    }                                                                                    \
    mmsa::count_launch();                                                                \
  } while (0)

// dispatch on activation storage dtype
#define MMSA_DISPATCH_DTYPE(dtype, T, ...)                    \
  do {                                                        \
    if ((dtype) == MMSA_F32) {                                \
      typedef float T;                                        \
      __VA_ARGS__;                                            \
    } else if ((dtype) == MMSA_BF16) {                        \
      typedef mmsa::bf16 T;                                   \
      __VA_ARGS__;                                            \
    } else {                                                  \
      mmsa::set_error("mmsa: unknown dtype %d", (int)(dtype)); \
      return MMSA_ERR_ARG;                                    \
    }                                                         \
  } while (0)

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
// round-trip through the storage type (so fwd and bwd recomputation see identical values)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f(from_f<T>(v)); }

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct VecN;
template <> struct VecN<float> { static constexpr int N = 4; };
template <> struct VecN<bf16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ void load_vec(const T* p, float* out);
template <> __device__ __forceinline__ void load_vec<float>(const float* p, float* out) {
  float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <> __device__ __forceinline__ void load_vec<bf16>(const bf16* p, float* out) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void store_vec(T* p, const float* in);
template <> __device__ __forceinline__ void store_vec<float>(float* p, const float* in) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <> __device__ __forceinline__ void store_vec<bf16>(bf16* p, const float* in) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum over blockDim.x threads (blockDim.x multiple of 32, <= 1024); result broadcast.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  float r = (lane < nw) ? smem32[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float kInvSqrt2Pi = 0.39894228040143267794f;
  return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * kInvSqrt2Pi * expf(-0.5f * x * x);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case MMSA_ACT_SIGMOID: return sigmoidf_(x);
    case MMSA_ACT_GELU: return gelu_erf(x);
    case MMSA_ACT_RELU: return fmaxf(x, 0.f);
    default: return x;
  }
}

// ------------------------------------------------------------------ Philox4x32-10 (dropout masks)
__device__ __forceinline__ uint32_t philox_first(uint64_t seed, uint64_t ctr) {
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


inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Per-DEVICE one-time state (function attributes such as the >48 KB shared-memory opt-in, SM counts, occupancy
// queries) is kept in arrays indexed by the current device: a process that drives several GPUs (nn.DataParallel,
// a test touching cuda:1) must see each of them initialised, not only the first one used.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}
struct PerDeviceOnce {
  unsigned char done[kMaxDevices] = {};
  bool pending() const { return !done[current_device()]; }
  void mark() { done[current_device()] = 1; }
};

// internal GEMM entry points (gemm_f32.cu / gemm_sm100.cu)
struct GemmDesc {
  // C[M,N] = act( A[M,K(+K2)] * B^T + bias + residual ), reduction over K
  int64_t M, N, K, K2;
  const void* A; int64_t lda; bool a_mn_major;   // a_mn_major: A stored [K,M] (element (m,k) at A[k*lda+m])
  const void* A2; int64_t lda2;                   // optional second K-segment of A (K-major only)
  const void* B; int64_t ldb; bool b_mn_major;   // K-major: B[n*ldb+k]; MN-major: B[k*ldb+n]
  const void* B2; int64_t ldb2; int64_t N1;       // optional second N-segment of an MN-major B (output columns >= N1 read B2):
                                                  // the weight gradient of a Linear over cat[x, x2] without the concat
  const float* bias;                              // [N] or null
  const void* residual; int64_t ldr;              // [M,N] out dtype family (in dtype), or null
  int act;
  void* C; int64_t ldc; int out_dtype;
  float alpha;                                    // scales the accumulator before bias
  float* colsum;                                  // tensor-core wgrad only: colsum[m] = sum_k A[m,k] (bias gradient), or null
};
int gemm_f32(const GemmDesc& d, cudaStream_t s);                  // fp32 operands, SIMT
int gemm_bf16_sm100(const GemmDesc& d, cudaStream_t s);           // bf16 operands, tcgen05 + TMA
bool gemm_bf16_sm100_supported(const GemmDesc& d);

}  // namespace mmsa
