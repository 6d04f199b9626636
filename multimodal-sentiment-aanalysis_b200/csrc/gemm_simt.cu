// gemm_simt.cu -- CUDA-core GEMM with generic operand majors and fp32 FMA accumulation.
// Two jobs: (1) the fp32 "parity mode" of every Linear (tensor cores have no fp32-exact mode: tf32
// keeps 10 mantissa bits and cannot meet the 1e-5 tolerance of BASELINE.json), (2) shapes the
// tensor-core tile cannot fill in bf16 mode (N < 16: the 3-way modality-weight and class heads,
// nn.Linear(64,3) / nn.Linear(128,3), MultimodalModel.py:174,198).
#include "common.cuh"

namespace mmsa {

struct SimtParams {
  int64_t M, N, K, K2;
  const void* A; int64_t lda; int a_mn;
  const void* A2; int64_t lda2;
  const void* B; int64_t ldb; int b_mn;
  const float* bias; const void* residual; int64_t ldr;
  int act; float alpha;
  void* C; int64_t ldc; int out_is_f32; int64_t split_stride;
  int64_t k_per_split;
};

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
  const int64_t Kt = p.K + p.K2;
  const int64_t kbeg = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t kend = kbeg + p.k_per_split < Kt ? kbeg + p.k_per_split : Kt;
  const T* A = reinterpret_cast<const T*>(p.A);
  const T* A2 = reinterpret_cast<const T*>(p.A2);
  const T* B = reinterpret_cast<const T*>(p.B);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += SBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk;
      if (p.a_mn) { kk = idx >> 6; mm = idx & 63; } else { mm = idx >> 4; kk = idx & 15; }
      int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < p.M && k < kend) {
        if (p.a_mn) v = to_f(A[k * p.lda + m]);
        else if (k < p.K) v = to_f(A[m * p.lda + k]);
        else v = to_f(A2[m * p.lda2 + (k - p.K)]);
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int nn, kk;
      if (p.b_mn) { kk = idx >> 6; nn = idx & 63; } else { nn = idx >> 4; kk = idx & 15; }
      int64_t n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < p.N && k < kend) v = p.b_mn ? to_f(B[k * p.ldb + n]) : to_f(B[n * p.ldb + k]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
    // two-level summation: the 16 products of one k-tile are summed first, then added to the running
    // total, so rounding error grows like sqrt(K/16) instead of sqrt(K) (fp32 parity mode, 1e-5)
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias) v += p.bias[n];
      if (p.residual) v += to_f(reinterpret_cast<const T*>(p.residual)[m * p.ldr + n]);
      v = apply_act(v, p.act);
      if (p.out_is_f32) reinterpret_cast<float*>(p.C)[(int64_t)blockIdx.z * p.split_stride + m * p.ldc + n] = v;
      else reinterpret_cast<bf16*>(p.C)[m * p.ldc + n] = __float2bfloat16_rn(v);
    }
  }
}

template <typename T>
int gemm_simt(const GemmDesc& d, int splits, cudaStream_t s) {
  SimtParams p{};
  p.M = d.M; p.N = d.N; p.K = d.K; p.K2 = d.A2 ? d.K2 : 0;
  p.A = d.A; p.lda = d.lda; p.a_mn = d.a_mn_major;
  p.A2 = d.A2; p.lda2 = d.lda2;
  p.B = d.B; p.ldb = d.ldb; p.b_mn = d.b_mn_major;
  p.bias = d.bias; p.residual = d.residual; p.ldr = d.ldr; p.act = d.act; p.alpha = d.alpha;
  p.C = d.C; p.ldc = d.ldc; p.out_is_f32 = (d.out_dtype == MMSA_F32);
  p.split_stride = d.M * d.ldc;
  int64_t Kt = p.K + p.K2;
  if (splits < 1) splits = 1;
  int64_t kps = ceil_div(ceil_div(Kt, splits), SBK) * SBK;
  p.k_per_split = kps;
  int real_splits = (int)ceil_div(Kt, kps);
  dim3 grid((unsigned)ceil_div(d.N, SBN), (unsigned)ceil_div(d.M, SBM), (unsigned)real_splits);
  ProfScope prof(sizeof(T) == 4 ? "gemm_simt_f32" : "gemm_simt_bf16", s, 2.0 * (double)d.M * (double)d.N * (double)Kt);
  gemm_simt_kernel<T><<<grid, 256, 0, s>>>(p);
  MMSA_LAUNCH_CHECK("gemm_simt_kernel");
  return MMSA_OK;
}

int gemm_simt_f32(const GemmDesc& d, int splits, cudaStream_t s) { return gemm_simt<float>(d, splits, s); }
int gemm_simt_bf16(const GemmDesc& d, int splits, cudaStream_t s) { return gemm_simt<bf16>(d, splits, s); }
int gemm_simt_real_splits(int64_t Kt, int splits) {
  if (splits < 1) splits = 1;
  int64_t kps = ceil_div(ceil_div(Kt, splits), SBK) * SBK;
  return (int)ceil_div(Kt, kps);
}

}  // namespace mmsa
