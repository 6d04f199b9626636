// linear.cu -- C ABI of the three Linear products (fwd, dgrad, wgrad) over the two GEMM engines:
//   bf16 storage  -> tcgen05/TMA kernel (gemm_sm100.cu); products below 0.13 GFLOP -> mma.sync cluster kernel
//                    (gemm_small.cu); unaligned -> CUDA-core kernel
//   fp32 storage  -> CUDA-core fp32 kernel (gemm_simt.cu), the 1e-5 parity mode
// Replaces the cuBLAS addmm calls behind nn.Linear / MHA in/out-proj / gate
// (MultimodalModel.py:86,112-121,139-147,172-198).
#include <stdlib.h>
#include "common.cuh"

namespace mmsa {

int gemm_simt_f32(const GemmDesc& d, int splits, cudaStream_t s);
int gemm_simt_bf16(const GemmDesc& d, int splits, cudaStream_t s);
int gemm_simt_real_splits(int64_t Kt, int splits);
int gemm_bf16_sm100_splits(const GemmDesc& d, int splits, int bn, cudaStream_t s);
int gemm_bf16_sm100_wgrad_pair(const GemmDesc& d, int ksplit, int* real_ksplit, cudaStream_t s);
int gemm_tc_max_clusters(int size);
int gemm_tc_bn(int64_t N, bool need_colsum);
int gemm_num_sms();
bool gemm_bf16_small_ok(const GemmDesc& d);                      // gemm_small.cu: the [B,*] tail, latency-bound
int gemm_bf16_small(const GemmDesc& d, cudaStream_t s);

static int g_gemm_engine = 0;     // test hook: 1 = keep small products on the tcgen05 / CUDA-core engines
static bool small_ok(int dtype, const GemmDesc& d) { return dtype == MMSA_BF16 && g_gemm_engine != 1 && gemm_bf16_small_ok(d); }

constexpr int kMaxSplitsWs = 16;       // K-slices of the workspace split-K (mmsa_linear_wgrad_workspace reserves 16 slabs)

static int tc_real_splits(int64_t Kt, int splits) {
  int64_t kb = ceil_div(Kt, 64);
  if (splits < 1) splits = 1;
  if (splits > kb) splits = (int)kb;
  int64_t per = ceil_div(kb, splits);
  return (int)ceil_div(kb, per);
}

static bool use_tc(int dtype, const GemmDesc& d) {
  return dtype == MMSA_BF16 && d.N >= 16 && gemm_bf16_sm100_supported(d);
}

// Small-M products of the [B,*] tail (fusion.0: 256 x 256 x 2304) have one or two output tiles and a long reduction:
// narrow tiles and a split-K cluster spread them over tens of SMs instead of two (latency-, not throughput-bound).
static void tc_plan_small(const GemmDesc& d, int* bn, int* splits) {
  *bn = 0; *splits = 1;
  const int64_t Kt = d.K + (d.A2 ? d.K2 : 0);
  const int64_t kb = ceil_div(Kt, 64);
  if (d.M > 512 || kb < 16 || d.out_dtype != MMSA_F32 || d.residual || d.act != MMSA_ACT_NONE || d.a_mn_major) return;
  const int b = d.N >= 64 ? 64 : 0;
  const int64_t tiles = ceil_div(d.M, 128) * ceil_div(d.N, 64);
  if (tiles > 32 || b == 0) return;
  int sp = (int)(kb / 4);
  if (sp > 8) sp = 8;
  while (sp > 1 && tiles * sp > 128) --sp;
  if (sp < 2) return;
  *bn = b; *splits = sp;
}

static int run_gemm(int dtype, const GemmDesc& d, int splits, cudaStream_t s) {
  if (dtype == MMSA_F32) return gemm_simt_f32(d, splits, s);
  if (splits == 1 && small_ok(dtype, d)) return gemm_bf16_small(d, s);
  if (use_tc(dtype, d)) {
    int bn = 0, sp = 1;
    tc_plan_small(d, &bn, &sp);
    return gemm_bf16_sm100_splits(d, sp, bn, s);
  }
  return gemm_simt_bf16(d, splits, s);
}

// Tile width and K-split (= cluster size, <= 8) of a tensor-core wgrad.  The output dW[Nw,Kw] has few
// tiles while the reduction (B*L rows) is long, so the K range is split over a cluster per tile.
// Cost model per candidate (BN, S): rounds x (k-blocks per split x cycles per k-block + epilogue), with
// rounds = ceil(tiles / co-resident clusters of size S) and cycles per k-block = max(MMA issue 2*BN,
// shared-memory operand reads (16 KB + BN*128 B) / 128 B per cycle).
static void tc_plan_wgrad(int64_t Nw, int64_t Kw, int64_t kb_total, int* bn_out, int* splits_out) {
  // 192 is left out: with an MN-major A the 128x192 MMA runs ~20% below the 128x256 one (scripts/gemm_majors.py)
  static const int cands[3] = {256, 128, 64};
  double best = 1e30;
  *bn_out = 128; *splits_out = 1;
  for (int ci = 0; ci < 3; ++ci) {
    const int bn = cands[ci];
    if (bn > 64 && Kw <= bn / 2) continue;                 // do not pad a narrow output to a wide tile
    const int64_t tiles = ceil_div(Nw, 128) * ceil_div(Kw, bn);
    const double waste = (double)(ceil_div(Kw, bn) * bn) / (double)Kw;
    // measured operand ingest per SM ~56 B/clk: (16 KB + bn*128 B) / 56 cycles per k-block; MMA issue 2*bn
    const double ingest = (16384.0 + bn * 128.0) / 56.0;
    const double cyc_kb = ingest > 2.0 * bn ? ingest : 2.0 * bn;
    for (int sp = 1; sp <= 8; ++sp) {
      if (sp > 1 && kb_total / sp < 8) break;
      const int64_t per = ceil_div(kb_total, sp);
      if (ceil_div(kb_total, per) != sp) continue;
      const int64_t rounds = ceil_div(tiles, gemm_tc_max_clusters(sp));
      const double cost = (double)rounds * ((double)per * cyc_kb + 5000.0) * (waste > 1.3 ? waste : 1.0);
      if (cost < best * 0.97) { best = cost; *bn_out = bn; *splits_out = sp; }
    }
  }
}

// column sums of dy[M,N] in two deterministic stages
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(int64_t M, int64_t N, const T* __restrict__ dy, int64_t ld, int64_t rows_per_blk,
                      float* __restrict__ partials) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_blk;
  const int64_t r1 = r0 + rows_per_blk < M ? r0 + rows_per_blk : M;
  float s = 0.f;
  if (col < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += to_f(dy[r * ld + col]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    partials[(int64_t)blockIdx.y * N + col] = t;
  }
}

// out[i] = sum_s partials[s*stride + i] for a [rows, cols] matrix with output row stride ldo
__global__ void reduce_splits_kernel(const float* __restrict__ partials, int splits, int64_t stride, int64_t rows,
                                     int64_t cols, int64_t ldp, float* __restrict__ out, int64_t ldo) {
  int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols, c = i % cols;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partials[(int64_t)k * stride + r * ldp + c];
    out[r * ldo + c] = s;
  }
}

// Finish of the 2-CTA weight gradient (gemm_bf16_sm100_wgrad_pair): dw[r, c] = sum_ks ws[ks][r][c] in slice order
// (deterministic), 16-byte vectors; the last blocks sum the bias-gradient partials the same way.  The workspace was
// written microseconds earlier and is L2-resident (<= 16 x 2.4 MB against 126 MB).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int ksplit, int64_t rows, int64_t cols, float* __restrict__ dw,
                    int64_t lddw, const float* __restrict__ ws_colsum, float* __restrict__ db, int mat_blocks) {
  if ((int)blockIdx.x >= mat_blocks) {
    if (db == nullptr) return;
    const int64_t r = (int64_t)(blockIdx.x - mat_blocks) * blockDim.x + threadIdx.x;
    if (r < rows) {
      float s = 0.f;
      for (int k = 0; k < ksplit; ++k) s += ws_colsum[(int64_t)k * rows + r];
      db[r] = s;
    }
    return;
  }
  const int64_t c4n = cols >> 2, total = rows * c4n, slab = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)mat_blocks * blockDim.x) {
    const int64_t r = i / c4n, c = (i % c4n) << 2;
    const float* src = ws + r * cols + c;
    float4 a = *reinterpret_cast<const float4*>(src);
    for (int k = 1; k < ksplit; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(src + (int64_t)k * slab);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(dw + r * lddw + c) = a;
  }
}

// 2-CTA 256 x 256 tiles + workspace split-K for the big weight gradients: output rows a multiple of 256, a reduction long
// enough to give every K-slice >= 8 k-blocks, a 16-byte addressable dW.  Returns the K-slice count (0: not eligible).
static int wgrad_pair_ksplit(int64_t Nw, int64_t Kw, int64_t tokens, const float* dw, int64_t lddw) {
  if (Nw % 256 != 0 || Kw % 4 != 0 || Kw < 512 || tokens < 4096 || lddw % 4 != 0 || ((uintptr_t)dw % 16) != 0) return 0;
  const int64_t tiles = (Nw / 256) * ceil_div(Kw, 256);
  int pairs = gemm_tc_max_clusters(2);
  // MMSA_WGRAD_PAIRS (tuning probe): CTA pairs a weight gradient may occupy.  The weight gradients run on a helper stream
  // beside the main backward chain; leaving some SMs free lets main-chain kernels start without waiting for a 45 us unit.
  static int pairs_env = -1;
  if (pairs_env < 0) { const char* e = getenv("MMSA_WGRAD_PAIRS"); pairs_env = e ? atoi(e) : 0; }
  if (pairs_env > 0 && pairs_env < pairs) pairs = pairs_env;
  int ks = (int)(pairs / tiles);
  if (ks > kMaxSplitsWs) ks = kMaxSplitsWs;
  const int64_t kb = ceil_div(tokens, 64);
  while (ks > 1 && kb / ks < 8) --ks;
  return ks < 1 ? 1 : ks;
}

// ------------------------------------------------------------------ skinny Linear (N <= 8 output features)
// The 3-way modality-weight and class heads (nn.Linear(64,3) / nn.Linear(128,3), MultimodalModel.py:174,198) are far
// below any GEMM tile: one warp per row forward, one thread per element for dgrad, and ONE kernel for dW and db
// (column blocks of 32 x 8 row lanes, shared-memory tree in lane order) instead of GEMM + split reduce + column sums.
constexpr int kSkinnyMaxN = 8;

__global__ void __launch_bounds__(256)
skinny_fwd_kernel(int64_t M, int N, int K, const bf16* __restrict__ x, int64_t ldx, const bf16* __restrict__ w, int64_t ldw,
                  const float* __restrict__ bias, void* __restrict__ y, int64_t ldy, int out_is_f32, int act) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 8 + warp;
  if (m >= M) return;
  float acc[kSkinnyMaxN];
#pragma unroll
  for (int c = 0; c < kSkinnyMaxN; ++c) acc[c] = 0.f;
#pragma unroll 4
  for (int k = lane; k < K; k += 32) {                 // unrolled: the loads of four k-steps are in flight together
    const float xv = to_f(x[m * ldx + k]);
#pragma unroll
    for (int c = 0; c < kSkinnyMaxN; ++c) if (c < N) acc[c] += xv * to_f(w[(int64_t)c * ldw + k]);
  }
#pragma unroll
  for (int c = 0; c < kSkinnyMaxN; ++c) if (c < N) acc[c] = warp_sum(acc[c]);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kSkinnyMaxN; ++c) if (c < N) {
      float v = apply_act(acc[c] + (bias ? bias[c] : 0.f), act);
      if (out_is_f32) reinterpret_cast<float*>(y)[m * ldy + c] = v;
      else reinterpret_cast<bf16*>(y)[m * ldy + c] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void __launch_bounds__(256)
skinny_dgrad_kernel(int64_t M, int N, int K, const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ w, int64_t ldw,
                    void* __restrict__ dx, int64_t lddx, int out_is_f32) {
  const int64_t total = M * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / K; const int k = (int)(i % K);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < kSkinnyMaxN; ++c) if (c < N) a += to_f(dy[m * lddy + c]) * to_f(w[(int64_t)c * ldw + k]);
    if (out_is_f32) reinterpret_cast<float*>(dx)[m * lddx + k] = a;
    else reinterpret_cast<bf16*>(dx)[m * lddx + k] = __float2bfloat16_rn(a);
  }
}

// block = 32 columns (k) x 32 row lanes; block x covers columns [32 x, 32 x + 32); block 0 also forms db.  The kernel is a
// pure latency chain (a [256, 3]^T x [256, 64..128] product): 32 row lanes and a 4-way unrolled row walk leave each thread
// two rounds of loads at B = 256 instead of the 32 dependent L2 round trips of an 8-lane, one-row-at-a-time walk.
constexpr int kSkinnyWgRows = 32;
__global__ void __launch_bounds__(32 * kSkinnyWgRows)
skinny_wgrad_kernel(int64_t M, int N, int K, const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                    float* __restrict__ dw, int64_t lddw, float* __restrict__ db) {
  __shared__ float red[kSkinnyWgRows][kSkinnyMaxN][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + tx;
  const bool want_b = db != nullptr && blockIdx.x == 0 && tx < N;
  float acc[kSkinnyMaxN], bsum = 0.f;
#pragma unroll
  for (int c = 0; c < kSkinnyMaxN; ++c) acc[c] = 0.f;
#pragma unroll 4
  for (int64_t m = ty; m < M; m += kSkinnyWgRows) {
    const float xv = k < K ? to_f(x[m * ldx + k]) : 0.f;
#pragma unroll
    for (int c = 0; c < kSkinnyMaxN; ++c) if (c < N) acc[c] += to_f(dy[m * lddy + c]) * xv;
    if (want_b) bsum += to_f(dy[m * lddy + tx]);
  }
#pragma unroll
  for (int c = 0; c < kSkinnyMaxN; ++c) red[ty][c][tx] = acc[c];
  __syncthreads();
  if (ty < kSkinnyMaxN && ty < N && k < K && dw != nullptr) {      // warp c sums output row c in row-lane order
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < kSkinnyWgRows; ++r) t += red[r][ty][tx];
    dw[(int64_t)ty * lddw + k] = t;
  }
  if (db != nullptr && blockIdx.x == 0) {
    __syncthreads();
    red[ty][0][tx] = bsum;
    __syncthreads();
    if (ty == 0 && tx < N) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < kSkinnyWgRows; ++r) t += red[r][0][tx];
      db[tx] = t;
    }
  }
}

static bool skinny_ok(int dtype, int64_t N) { return dtype == MMSA_BF16 && N <= kSkinnyMaxN; }

constexpr int kMaxSplits = 16;
constexpr int kColsumRowSplits = 64;

}  // namespace mmsa

using namespace mmsa;

extern "C" {

/* test hook: 0 = products below 0.13 GFLOP use the mma.sync cluster kernel (default), 1 = they stay on tcgen05 */
void mmsa_debug_gemm_engine(int engine) { g_gemm_engine = engine; }

/* tuning probe (not part of the product path): raw bf16 tcgen05 GEMM with explicit operand majors.
 * C[M,N] (fp32, ldc) = A * B^T over K; a_mn: A stored [K,M]; b_mn: B stored [K,N]. */
int mmsa_debug_gemm(int a_mn, int b_mn, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                    const void* B, int64_t ldb, float* C, int64_t ldc, int splits, int bn, void* stream) {
  MMSA_REQUIRE_DEVICE();
  GemmDesc d{};
  d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.a_mn_major = a_mn != 0; d.B = B; d.ldb = ldb; d.b_mn_major = b_mn != 0;
  d.C = C; d.ldc = ldc; d.out_dtype = MMSA_F32; d.alpha = 1.f; d.act = MMSA_ACT_NONE;
  return gemm_bf16_sm100_splits(d, splits, bn, (cudaStream_t)stream);
}

int mmsa_linear_fwd(int dtype, int64_t M, int64_t N, int64_t K, int64_t K2, const void* x, int64_t ldx,
                    const void* x2, int64_t ldx2, const void* w, int64_t ldw, const float* bias,
                    const void* residual, int64_t ldr, int act, void* y, int64_t ldy, int out_dtype,
                    void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(dtype == MMSA_F32 || dtype == MMSA_BF16, "mmsa_linear_fwd: bad dtype");
  MMSA_REQUIRE(M >= 0 && N > 0 && K > 0, "mmsa_linear_fwd: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  MMSA_REQUIRE(!(dtype == MMSA_F32 && out_dtype == MMSA_BF16), "mmsa_linear_fwd: fp32 mode writes fp32");
  if (M == 0) return MMSA_OK;
  if (skinny_ok(dtype, N) && x2 == nullptr && residual == nullptr) {
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof("skinny_fwd", s, 2.0 * (double)M * N * K);
    skinny_fwd_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, s>>>(M, (int)N, (int)K, (const bf16*)x, ldx, (const bf16*)w, ldw, bias, y, ldy,
                                                                out_dtype == MMSA_F32, act);
    MMSA_LAUNCH_CHECK("skinny_fwd_kernel");
    return MMSA_OK;
  }
  GemmDesc d{};
  d.M = M; d.N = N; d.K = K; d.K2 = x2 ? K2 : 0;
  d.A = x; d.lda = ldx; d.a_mn_major = false; d.A2 = x2; d.lda2 = ldx2;
  d.B = w; d.ldb = ldw; d.b_mn_major = false;
  d.bias = bias; d.residual = residual; d.ldr = ldr; d.act = act;
  d.C = y; d.ldc = ldy; d.out_dtype = out_dtype; d.alpha = 1.f;
  return run_gemm(dtype, d, 1, (cudaStream_t)stream);
}

int mmsa_linear_dgrad(int dtype, int64_t M, int64_t N, int64_t K, const void* dy, int64_t lddy, const void* w,
                      int64_t ldw, const void* residual, int64_t ldr, void* dx, int64_t lddx, int out_dtype,
                      void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(dtype == MMSA_F32 || dtype == MMSA_BF16, "mmsa_linear_dgrad: bad dtype");
  MMSA_REQUIRE(M >= 0 && N > 0 && K > 0, "mmsa_linear_dgrad: bad shape");
  MMSA_REQUIRE(!(dtype == MMSA_F32 && out_dtype == MMSA_BF16), "mmsa_linear_dgrad: fp32 mode writes fp32");
  if (M == 0) return MMSA_OK;
  if (skinny_ok(dtype, N) && residual == nullptr) {
    cudaStream_t s = (cudaStream_t)stream;
    int64_t blocks = ceil_div(M * K, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ProfScope prof("skinny_dgrad", s, 2.0 * (double)M * N * K);
    skinny_dgrad_kernel<<<(unsigned)blocks, 256, 0, s>>>(M, (int)N, (int)K, (const bf16*)dy, lddy, (const bf16*)w, ldw, dx, lddx,
                                                        out_dtype == MMSA_F32);
    MMSA_LAUNCH_CHECK("skinny_dgrad_kernel");
    return MMSA_OK;
  }
  GemmDesc d{};
  d.M = M; d.N = K; d.K = N; d.K2 = 0;
  d.A = dy; d.lda = lddy; d.a_mn_major = false; d.A2 = nullptr;
  d.B = w; d.ldb = ldw; d.b_mn_major = true;      // W[N,K]: reduction index n is the row -> MN-major B
  d.bias = nullptr; d.residual = residual; d.ldr = ldr; d.act = MMSA_ACT_NONE;
  d.C = dx; d.ldc = lddx; d.out_dtype = out_dtype; d.alpha = 1.f;
  return run_gemm(dtype, d, 1, (cudaStream_t)stream);
}

}  // extern "C"

// 2-CTA weight gradient (+ ordered slice reduction) of dy^T [x | x2]; x2 may be null.  Returns MMSA_OK, or -1 when the shape is
// not eligible (the caller then takes the generic path).
static int wgrad_pair_path(int64_t M, int64_t N, int64_t K, int64_t K2, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                           const void* x2, int64_t ldx2, float* dw, int64_t lddw, float* db, float* ws, float* ws_colsum,
                           cudaStream_t s) {
  const int64_t Kt = K + (x2 ? K2 : 0);
  if (getenv("MMSA_WGRAD_PAIR_OFF") != nullptr) return -1;
  if (x2 && (K % 256 != 0 || ((uintptr_t)x2 % 16) != 0 || ldx2 % 8 != 0)) return -1;
  const int ks_req = wgrad_pair_ksplit(N, Kt, M, dw, lddw);
  if (ks_req <= 0) return -1;
  GemmDesc d{};
  d.M = N; d.N = Kt; d.K = M; d.K2 = 0;
  d.A = dy; d.lda = lddy; d.a_mn_major = true; d.A2 = nullptr;
  d.B = x; d.ldb = ldx; d.b_mn_major = true;
  d.B2 = x2; d.ldb2 = ldx2; d.N1 = x2 ? K : 0;
  d.bias = nullptr; d.residual = nullptr; d.act = MMSA_ACT_NONE; d.alpha = 1.f; d.out_dtype = MMSA_F32;
  if (!gemm_bf16_sm100_supported(d)) return -1;
  d.C = ks_req > 1 ? (void*)ws : (void*)dw;
  d.ldc = ks_req > 1 ? Kt : lddw;
  d.colsum = db ? (ks_req > 1 ? ws_colsum : db) : nullptr;
  int ks = 1;
  int rc = gemm_bf16_sm100_wgrad_pair(d, ks_req, &ks, s);
  if (rc) return rc;
  if (ks_req > 1) {
    const int64_t total4 = N * (Kt / 4);
    int mat_blocks = (int)ceil_div(total4, 256);
    if (mat_blocks > 148 * 4) mat_blocks = 148 * 4;
    const int cs_blocks = db ? (int)ceil_div(N, 256) : 0;
    ProfScope prof("wgrad_reduce", s, 4.0 * (double)N * Kt * (ks + 1));
    wgrad_reduce_kernel<<<(unsigned)(mat_blocks + cs_blocks), 256, 0, s>>>(ws, ks, N, Kt, dw, lddw, ws_colsum, db, mat_blocks);
    MMSA_LAUNCH_CHECK("wgrad_reduce_kernel");
  }
  return MMSA_OK;
}

extern "C" {

int64_t mmsa_linear_wgrad_workspace(int dtype, int64_t M, int64_t N, int64_t K) {
  (void)dtype; (void)M;
  return (int64_t)sizeof(float) * (kMaxSplits * N * K + (int64_t)(kColsumRowSplits > kMaxSplits ? kColsumRowSplits : kMaxSplits) * N);
}

int mmsa_linear_wgrad(int dtype, int64_t M, int64_t N, int64_t K, const void* dy, int64_t lddy, const void* x,
                      int64_t ldx, float* dw, int64_t lddw, float* db, void* workspace, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(dtype == MMSA_F32 || dtype == MMSA_BF16, "mmsa_linear_wgrad: bad dtype");
  MMSA_REQUIRE(M > 0 && N > 0 && K > 0, "mmsa_linear_wgrad: bad shape");
  MMSA_REQUIRE(workspace != nullptr, "mmsa_linear_wgrad: workspace required");
  cudaStream_t s = (cudaStream_t)stream;
  if (skinny_ok(dtype, N)) {
    if (dw == nullptr && db == nullptr) return MMSA_OK;
    ProfScope prof("skinny_wgrad", s, 2.0 * (double)M * N * K);
    skinny_wgrad_kernel<<<(unsigned)ceil_div(K, 32), 32 * kSkinnyWgRows, 0, s>>>(M, (int)N, (int)K, (const bf16*)dy, lddy, (const bf16*)x, ldx, dw, lddw, db);
    MMSA_LAUNCH_CHECK("skinny_wgrad_kernel");
    return MMSA_OK;
  }
  float* ws = reinterpret_cast<float*>(workspace);
  float* ws_colsum = ws + (int64_t)kMaxSplits * N * K;
  bool db_done = false;
  if (dw != nullptr) {
    GemmDesc d{};
    d.M = N; d.N = K; d.K = M; d.K2 = 0;
    d.A = dy; d.lda = lddy; d.a_mn_major = true;    // dy[M,N]: reduction index m is the row
    d.A2 = nullptr;
    d.B = x; d.ldb = ldx; d.b_mn_major = true;      // x[M,K]:  reduction index m is the row
    d.bias = nullptr; d.residual = nullptr; d.act = MMSA_ACT_NONE; d.alpha = 1.f; d.out_dtype = MMSA_F32;
    d.C = dw; d.ldc = lddw;
    if (small_ok(dtype, d)) {
      d.colsum = db;                                  // bias gradient from the ones-fragment MMA of the same launch
      int rc = gemm_bf16_small(d, s);
      if (rc) return rc;
      db_done = true;
    } else if (use_tc(dtype, d) && wgrad_pair_path(M, N, K, 0, dy, lddy, x, ldx, nullptr, 0, dw, lddw, db, ws, ws_colsum, s) == MMSA_OK) {
      // 2-CTA 256 x 256 tiles, (tile, K-slice) work units filling the 74 CTA pairs, fp32 partials into the (L2-resident)
      // workspace, bias-gradient partials from the ones-tile MMA, then one small kernel sums the slices in order
      db_done = true;
    } else if (use_tc(dtype, d)) {
      // one launch: split-K partials are summed by the last CTA of each tile, and the bias gradient
      // falls out of an extra ones-tile MMA (gemm_sm100.cu)
      d.colsum = db;
      int bn = 0, splits = 1;
      tc_plan_wgrad(N, K, ceil_div(M, 64), &bn, &splits);
      // (The big weight gradients run on a second stream next to the dgrad / attention chain (ops.py); 4 K-slices would
      //  cost 10 % less SM-time than the latency-optimal 6 and shave 1.4 % off the step, but slow the kernel itself by
      //  30 % -- MMSA_WGRAD_SPLITS=4 reproduces it.  The planner keeps the latency-optimal split.)
      if (const char* e = getenv("MMSA_WGRAD_SPLITS")) splits = atoi(e);      // tuning probes only
      if (const char* e = getenv("MMSA_WGRAD_BN")) bn = atoi(e);
      int rc = gemm_bf16_sm100_splits(d, splits, bn, s);
      if (rc) return rc;
      db_done = true;
    } else {
      // split the long reduction (M = B*L) so that the few output tiles still fill the SMs
      int64_t tiles = ceil_div(N, 64) * ceil_div(K, 64);
      int splits = (int)(gemm_num_sms() / (tiles > 0 ? tiles : 1));
      if (splits > kMaxSplits) splits = kMaxSplits;
      if (splits < 1) splits = 1;
      int real = gemm_simt_real_splits(M, splits);
      if (real <= 1) {
        int rc = run_gemm(dtype, d, 1, s);
        if (rc) return rc;
      } else {
        d.C = ws; d.ldc = K;
        int rc = run_gemm(dtype, d, splits, s);
        if (rc) return rc;
        int64_t total = N * K;
        int64_t blocks = ceil_div(total, 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        ProfScope prof("reduce_splits", s, 4.0 * (double)N * K * (real + 1));
        reduce_splits_kernel<<<(unsigned)blocks, 256, 0, s>>>(ws, real, N * K, N, K, K, dw, lddw);
        MMSA_LAUNCH_CHECK("reduce_splits_kernel");
      }
    }
  }
  if (db != nullptr && !db_done) {
    int rs = (int)(M < kColsumRowSplits * 8 ? ceil_div(M, 8) : kColsumRowSplits);
    if (rs < 1) rs = 1;
    int64_t rpb = ceil_div(M, rs);
    rs = (int)ceil_div(M, rpb);
    dim3 block(32, 8), grid((unsigned)ceil_div(N, 32), (unsigned)rs);
    {
      ProfScope prof("colsum_partial", s, (double)M * N * (dtype == MMSA_F32 ? 4 : 2));
      MMSA_DISPATCH_DTYPE(dtype, T, (colsum_partial_kernel<T><<<grid, block, 0, s>>>(M, N, (const T*)dy, lddy, rpb, ws_colsum)));
    }
    MMSA_LAUNCH_CHECK("colsum_partial_kernel");
    ProfScope prof("reduce_splits", s, 4.0 * (double)N * (rs + 1));
    reduce_splits_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, s>>>(ws_colsum, rs, N, 1, N, N, db, N);
    MMSA_LAUNCH_CHECK("reduce_splits_kernel");
  }
  return MMSA_OK;
}

/* dw[N, K + K2] = dy^T [x | x2], db = column sums of dy: the weight gradient of a Linear whose input is the feature-axis
 * concat of two tensors (the gate, MultimodalModel.py:147), without materialising the concat.  Eligible shapes run as ONE
 * 2-CTA launch (B operand from two tensor maps); others as two mmsa_linear_wgrad calls on the column blocks of dw. */
int mmsa_linear_wgrad2(int dtype, int64_t M, int64_t N, int64_t K, int64_t K2, const void* dy, int64_t lddy, const void* x,
                       int64_t ldx, const void* x2, int64_t ldx2, float* dw, int64_t lddw, float* db, void* workspace,
                       void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(dtype == MMSA_F32 || dtype == MMSA_BF16, "mmsa_linear_wgrad2: bad dtype");
  MMSA_REQUIRE(M > 0 && N > 0 && K > 0 && K2 > 0 && x2 != nullptr && dw != nullptr, "mmsa_linear_wgrad2: bad arguments");
  MMSA_REQUIRE(workspace != nullptr, "mmsa_linear_wgrad2: workspace required (mmsa_linear_wgrad_workspace(dtype, M, N, K + K2))");
  float* ws = reinterpret_cast<float*>(workspace);
  float* ws_colsum = ws + (int64_t)kMaxSplits * N * (K + K2);
  if (dtype == MMSA_BF16) {
    GemmDesc probe{};
    probe.M = N; probe.N = K + K2; probe.K = M; probe.A = dy; probe.lda = lddy; probe.a_mn_major = true;
    probe.B = x; probe.ldb = ldx; probe.b_mn_major = true;
    if (use_tc(dtype, probe)) {
      int rc = wgrad_pair_path(M, N, K, K2, dy, lddy, x, ldx, x2, ldx2, dw, lddw, db, ws, ws_colsum, (cudaStream_t)stream);
      if (rc != -1) return rc;
    }
  }
  int rc = mmsa_linear_wgrad(dtype, M, N, K, dy, lddy, x, ldx, dw, lddw, db, workspace, stream);
  if (rc) return rc;
  return mmsa_linear_wgrad(dtype, M, N, K2, dy, lddy, x2, ldx2, dw + K, lddw, nullptr, workspace, stream);
}

}  // extern "C"
