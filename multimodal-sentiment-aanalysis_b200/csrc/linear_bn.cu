// linear_bn.cu -- Linear -> BatchNorm1d (-> GELU, or ReLU first) -> Dropout of the [B,*] tail in ONE launch
// (MultimodalModel.py:179-199 fusion / heads, ME-MHACL/model.py:82-97), forward.
//
// The tail's layers are [B <= 256, K] x [N <= 256, K] products followed by a column-statistics kernel; issued separately
// each pair costs two launches, an fp32 round trip of z through L2 and a three-pass BatchNorm kernel that is pure latency
// (a 256 x 256 tile).  Here ONE CTA owns ALL batch rows of 16 output columns, so the batch statistics of its columns never
// leave the CTA: 8 warps x 32 rows of mma.sync.m16n8k16 (fp32 accumulators in registers), operands staged by a 4-deep
// cp.async ring, the reduction optionally split over a thread-block cluster (<= 8 K-slices, summed in slice order through
// distributed shared memory into rank 0: deterministic, no workspace) and the epilogue does bias, the two-pass column
// statistics (shuffle + fixed-order cross-warp sum), running-stat update, affine, activation and Philox dropout -- the
// same arithmetic and the same Philox indexing (offset + row * N + col) as mmsa_bn_act_fwd, whose backward consumes what
// this kernel saves (z, mean, rstd, keep mask).
#include "common.cuh"

namespace mmsa {

namespace {

constexpr int LB_ROWS = 256;                       // batch rows owned by one CTA (= the largest batch this kernel takes)
constexpr int LB_COLS = 16;                        // output columns per CTA
constexpr int LB_TK = 64, LB_PITCH = LB_TK + 8;    // 144-byte tile rows: ldmatrix rows hit distinct banks
constexpr int LB_STAGES = 4;
constexpr int LB_THREADS = 256;
constexpr int LB_MAX_SLICES = 8;
constexpr int LB_KB_PER_SLICE = 6;                 // target k-blocks per CTA before the reduction is split
constexpr int LB_A_BYTES = LB_ROWS * LB_PITCH * 2;
constexpr int LB_B_BYTES = LB_COLS * LB_PITCH * 2;
constexpr int LB_STAGE_BYTES = LB_A_BYTES + LB_B_BYTES;
constexpr int LB_SMEM = LB_STAGES * LB_STAGE_BYTES;
constexpr int LB_ZP = LB_COLS + 1;                  // pitch (floats) of the epilogue's z tile
static_assert(LB_ROWS * LB_ZP * 4 <= LB_SMEM, "the z tile aliases the operand ring");
static_assert(LB_THREADS * 16 * 4 <= LB_SMEM, "the partial tile of the cluster sum aliases the operand ring");

struct LbParams {
  int M, N, K;
  const bf16* X; long long ldx;
  const bf16* W; long long ldw;
  const float* bias; const float* gamma; const float* beta;
  float* running_mean; float* running_var; int64_t* nbt;
  float momentum, eps; int training, order;
  float dropout_p; uint8_t* keep_mask; int mask_given; uint64_t seed, offset; const uint64_t* rng_state;
  float* z; void* y; int y_is_f32; bf16* y_lp; float* save_mean; float* save_rstd;
  int kb_total, kb_per_slice, slices;
};

__device__ __forceinline__ uint32_t lb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lb_cp16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void lb_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void lb_ldsm(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lb_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int lb_clamp16(int elems) { return elems <= 0 ? 0 : (elems >= 8 ? 16 : elems * 2); }
__device__ __forceinline__ void lb_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(LB_THREADS, 1)
linear_bn_act_kernel(const LbParams p) {
  extern __shared__ __align__(16) uint8_t lb_smem[];
  __shared__ float red[LB_THREADS / 16][LB_COLS];
  __shared__ float stat[2][LB_COLS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * LB_COLS;
  const int slice = blockIdx.z;                              // cluster = (1, 1, slices): rank == blockIdx.z
  const int kb0 = slice * p.kb_per_slice;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
  const int nk = kb1 - kb0;
  const int wm = warp * 32;                                  // the warp's 32 rows x 16 columns
  const bool warp_live = wm < p.M;                           // warps past the batch only help with the loads

  float acc[2][2][4] = {};
  const uint32_t s0 = lb_smem_u32(lb_smem);
  auto issue = [&](int j) {                                  // k-block j of this slice -> ring slot j % LB_STAGES
    const int kg0 = (kb0 + j) * LB_TK;
    const uint32_t sa = s0 + (uint32_t)(j % LB_STAGES) * LB_STAGE_BYTES, sb = sa + LB_A_BYTES;
#pragma unroll
    for (int i = 0; i < LB_ROWS * 8 / LB_THREADS; ++i) {
      const int c = tid + i * LB_THREADS, row = c >> 3, kc = (c & 7) * 8, kg = kg0 + kc;
      const bf16* src = p.X; int bytes = 0;
      if (row < p.M && kg < p.K) { src = p.X + (long long)row * p.ldx + kg; bytes = lb_clamp16(p.K - kg); }
      lb_cp16(sa + (uint32_t)(row * LB_PITCH + kc) * 2u, src, bytes);
    }
    if (tid < LB_COLS * 8) {
      const int row = tid >> 3, kc = (tid & 7) * 8, kg = kg0 + kc, col = n0 + row;
      const bf16* src = p.W; int bytes = 0;
      if (col < p.N && kg < p.K) { src = p.W + (long long)col * p.ldw + kg; bytes = lb_clamp16(p.K - kg); }
      lb_cp16(sb + (uint32_t)(row * LB_PITCH + kc) * 2u, src, bytes);
    }
  };
#pragma unroll
  for (int j = 0; j < LB_STAGES - 1; ++j) {
    if (j < nk) issue(j);
    lb_commit();
  }
  const int l7 = lane & 7, l8 = (lane >> 3) & 1, l16 = lane >> 4;
  const uint32_t a_off = (uint32_t)((wm + l8 * 8 + l7) * LB_PITCH + l16 * 8) * 2u;
  const uint32_t b_off = (uint32_t)((l16 * 8 + l7) * LB_PITCH + l8 * 8) * 2u;
  for (int j = 0; j < nk; ++j) {
    lb_wait<LB_STAGES - 2>();
    __syncthreads();                                         // k-block j has landed; slot (j-1) % LB_STAGES is free
    if (j + LB_STAGES - 1 < nk) issue(j + LB_STAGES - 1);
    lb_commit();
    if (warp_live) {
      const uint32_t aS = s0 + (uint32_t)(j % LB_STAGES) * LB_STAGE_BYTES + a_off;
      const uint32_t bS = s0 + (uint32_t)(j % LB_STAGES) * LB_STAGE_BYTES + LB_A_BYTES + b_off;
#pragma unroll
      for (int ks = 0; ks < LB_TK / 16; ++ks) {
        uint32_t a[2][4], b[4];
        lb_ldsm(aS + ks * 32, a[0]);
        lb_ldsm(aS + ks * 32 + 16 * LB_PITCH * 2, a[1]);
        lb_ldsm(bS + ks * 32, b);                            // b[0..1]: columns 0-7, b[2..3]: columns 8-15
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          lb_mma(acc[mi][0], a[mi], b[0], b[1]);
          lb_mma(acc[mi][1], a[mi], b[2], b[3]);
        }
      }
    }
  }
  lb_wait<0>();

  if (p.slices > 1) {
    // cluster sum: every CTA parks its partial tile over the drained ring; rank 0 adds the others in slice order
    __syncthreads();
    float4* mine = reinterpret_cast<float4*>(lb_smem) + tid;           // [group][thread]: conflict-free 16-byte rows
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        mine[(mi * 2 + nt) * LB_THREADS] = make_float4(acc[mi][nt][0], acc[mi][nt][1], acc[mi][nt][2], acc[mi][nt][3]);
    lb_cluster_sync();
    if (slice == 0) {
      const uint32_t my = lb_smem_u32(mine);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint32_t off = my + (uint32_t)((mi * 2 + nt) * LB_THREADS * 16);
          float4 u[LB_MAX_SLICES - 1];
#pragma unroll
          for (int r = 1; r < LB_MAX_SLICES; ++r) {
            if (r < p.slices) {
              uint32_t ra;
              asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(off), "r"(r));
              asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(u[r - 1].x), "=f"(u[r - 1].y), "=f"(u[r - 1].z), "=f"(u[r - 1].w) : "r"(ra) : "memory");
            }
          }
#pragma unroll
          for (int r = 1; r < LB_MAX_SLICES; ++r) {
            if (r < p.slices) {
              acc[mi][nt][0] += u[r - 1].x; acc[mi][nt][1] += u[r - 1].y; acc[mi][nt][2] += u[r - 1].z; acc[mi][nt][3] += u[r - 1].w;
            }
          }
        }
      }
    }
    lb_cluster_sync();                                       // the peers' shared memory stays valid until rank 0 has read it
    if (slice != 0) return;
  }

  // ---- epilogue (rank 0).  The accumulators go through a shared-memory tile (over the drained ring) so that the statistics
  //      and the element-wise tail are short ROLLED loops: unrolled over the 32 values a thread holds, the Philox + erf code
  //      alone was 118 KB of straight-line SASS executed once per warp -- the kernel spent its time on instruction fetch.
  __syncthreads();                                           // every warp is done with the operand ring
  float* zt = reinterpret_cast<float*>(lb_smem);             // [LB_ROWS][LB_ZP] fp32: z = x w^T + bias
  {
    const int g = lane >> 2, q = (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int cl = nt * 8 + q, col = n0 + cl;
      const bool cok = col < p.N;                            // N % 8 == 0: the pair (col, col + 1) is inside or outside together
      const float b0 = (cok && p.bias) ? __ldg(p.bias + col) : 0.f, b1 = (cok && p.bias) ? __ldg(p.bias + col + 1) : 0.f;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = wm + mi * 16 + g + h * 8;
          zt[row * LB_ZP + cl] = acc[mi][nt][h * 2] + b0;
          zt[row * LB_ZP + cl + 1] = acc[mi][nt][h * 2 + 1] + b1;
        }
      }
    }
  }
  __syncthreads();
  const bool pre_relu = p.order == MMSA_RELU_THEN_BN;
  const int sc = tid & (LB_COLS - 1), sg = tid >> 4;         // statistics: column sc, rows [16 sg, 16 sg + 16)
  if (p.training) {
    const float m_f = (float)p.M;
    const int r_lo = sg * 16, r_hi = min(r_lo + 16, p.M);
    float s = 0.f;
    for (int r = r_lo; r < r_hi; ++r) { const float z = zt[r * LB_ZP + sc]; s += pre_relu ? fmaxf(z, 0.f) : z; }
    red[sg][sc] = s;
    __syncthreads();
    if (tid < LB_COLS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LB_THREADS / 16; ++w) t += red[w][tid];
      stat[0][tid] = t / m_f;
    }
    __syncthreads();
    const float mean = stat[0][sc];
    s = 0.f;
    for (int r = r_lo; r < r_hi; ++r) {
      const float z = zt[r * LB_ZP + sc];
      const float d = (pre_relu ? fmaxf(z, 0.f) : z) - mean;
      s += d * d;
    }
    red[sg][sc] = s;                                         // pass 1's reads of `red` finished before the last barrier
    __syncthreads();
    if (tid < LB_COLS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LB_THREADS / 16; ++w) t += red[w][tid];
      const float var = t / m_f;
      const float rstd = 1.f / sqrtf(var + p.eps);
      stat[1][tid] = rstd;
      const int col = n0 + tid;
      if (col < p.N) {
        const float mu = stat[0][tid];
        p.save_mean[col] = mu;
        p.save_rstd[col] = rstd;
        if (p.running_mean != nullptr) {
          const float unbiased = p.M > 1 ? t / (float)(p.M - 1) : var;
          p.running_mean[col] = (1.f - p.momentum) * p.running_mean[col] + p.momentum * mu;
          p.running_var[col] = (1.f - p.momentum) * p.running_var[col] + p.momentum * unbiased;
        }
      }
    }
    if (p.nbt != nullptr && blockIdx.x == 0 && tid == 0) *p.nbt += 1;
  } else if (tid < LB_COLS) {
    const int col = n0 + tid;
    const bool cok = col < p.N;
    const float mu = cok ? p.running_mean[col] : 0.f;
    const float rstd = cok ? 1.f / sqrtf(p.running_var[col] + p.eps) : 0.f;
    stat[0][tid] = mu;
    stat[1][tid] = rstd;
    if (cok) { p.save_mean[col] = mu; p.save_rstd[col] = rstd; }
  }
  __syncthreads();

  // affine, activation, dropout, stores: one column pair per thread and iteration (8 threads cover a row's 16 columns)
  uint64_t seed = p.seed, offset = p.offset;
  if (p.rng_state != nullptr) { seed = p.rng_state[0]; offset += p.rng_state[1]; }     // device-resident stream position
  const bool drop = p.training && p.dropout_p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
  const int cl = (tid & 7) * 2, col = n0 + cl;
  if (col >= p.N) return;
  const float mean0 = stat[0][cl], mean1 = stat[0][cl + 1], rstd0 = stat[1][cl], rstd1 = stat[1][cl + 1];
  const float g0 = __ldg(p.gamma + col), g1 = __ldg(p.gamma + col + 1);
  const float t0 = __ldg(p.beta + col), t1 = __ldg(p.beta + col + 1);
#pragma unroll 1
  for (int row = tid >> 3; row < p.M; row += LB_THREADS / 8) {
    const float z0 = zt[row * LB_ZP + cl], z1 = zt[row * LB_ZP + cl + 1];
    const long long idx = (long long)row * p.N + col;
    *reinterpret_cast<float2*>(p.z + idx) = make_float2(z0, z1);
    float y0 = ((pre_relu ? fmaxf(z0, 0.f) : z0) - mean0) * rstd0 * g0 + t0;
    float y1 = ((pre_relu ? fmaxf(z1, 0.f) : z1) - mean1) * rstd1 * g1 + t1;
    if (p.order == MMSA_BN_THEN_GELU) { y0 = gelu_erf(y0); y1 = gelu_erf(y1); }
    if (drop) {
      uint8_t k0, k1;
      if (p.mask_given) { k0 = p.keep_mask[idx]; k1 = p.keep_mask[idx + 1]; }
      else {
        const uint32_t r0 = philox_first(seed, offset + (uint64_t)idx);
        const uint32_t r1 = philox_first(seed, offset + (uint64_t)idx + 1u);
        k0 = ((float)(r0 >> 8) * (1.f / 16777216.f)) >= p.dropout_p ? 1 : 0;
        k1 = ((float)(r1 >> 8) * (1.f / 16777216.f)) >= p.dropout_p ? 1 : 0;
        p.keep_mask[idx] = k0; p.keep_mask[idx + 1] = k1;
      }
      y0 = k0 ? y0 * keep_scale : 0.f;
      y1 = k1 ? y1 * keep_scale : 0.f;
    }
    if (p.y_is_f32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.y) + idx) = make_float2(y0, y1);
    else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(p.y) + idx) = __floats2bfloat162_rn(y0, y1);
    if (p.y_lp != nullptr) *reinterpret_cast<__nv_bfloat162*>(p.y_lp + idx) = __floats2bfloat162_rn(y0, y1);
  }
}

bool lb_shape_ok(int64_t M, int64_t N, int64_t K, int64_t ldx, int64_t ldw) {
  return M > 0 && M <= LB_ROWS && N > 0 && N % 8 == 0 && N <= 65535LL * LB_COLS && K > 0 && K % 8 == 0 && K <= (1 << 20) &&
         ldx % 8 == 0 && ldw % 8 == 0;
}

}  // namespace

}  // namespace mmsa

using namespace mmsa;

extern "C" {

int mmsa_linear_bn_act_supported(int64_t M, int64_t N, int64_t K, int64_t ldx, int64_t ldw) {
  return lb_shape_ok(M, N, K, ldx, ldw) ? 1 : 0;
}

int mmsa_linear_bn_act_fwd(int64_t M, int64_t N, int64_t K, const void* x, int64_t ldx, const void* w, int64_t ldw,
                           const float* bias, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float momentum, float eps, int training,
                           int order, float dropout_p, uint8_t* keep_mask, int mask_given, uint64_t seed, uint64_t offset,
                           const uint64_t* rng_state, float* z, int out_dtype, void* y, void* y_lp, float* save_mean,
                           float* save_rstd, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(lb_shape_ok(M, N, K, ldx, ldw), "mmsa_linear_bn_act_fwd: shape M=%lld N=%lld K=%lld (ldx %lld, ldw %lld) not "
               "supported (M <= %d, N %% 8 == 0, K %% 8 == 0); use mmsa_linear_fwd + mmsa_bn_act_fwd",
               (long long)M, (long long)N, (long long)K, (long long)ldx, (long long)ldw, LB_ROWS);
  MMSA_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0, "mmsa_linear_bn_act_fwd: operands must be 16-byte aligned");
  MMSA_REQUIRE(((uintptr_t)z % 8) == 0 && ((uintptr_t)y % 8) == 0 && ((uintptr_t)y_lp % 4) == 0,
               "mmsa_linear_bn_act_fwd: outputs must be 8-byte aligned");
  MMSA_REQUIRE(x && w && gamma && beta && z && y && save_mean && save_rstd, "mmsa_linear_bn_act_fwd: null argument");
  MMSA_REQUIRE(order == MMSA_BN_THEN_GELU || order == MMSA_RELU_THEN_BN || order == MMSA_BN_ONLY, "mmsa_linear_bn_act_fwd: bad order %d", order);
  MMSA_REQUIRE(training || (running_mean && running_var), "mmsa_linear_bn_act_fwd: eval mode needs running stats");
  MMSA_REQUIRE(!(training && dropout_p > 0.f) || keep_mask != nullptr, "mmsa_linear_bn_act_fwd: dropout needs keep_mask storage");
  MMSA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "mmsa_linear_bn_act_fwd: dropout_p out of [0,1)");
  MMSA_REQUIRE(out_dtype == MMSA_F32 || out_dtype == MMSA_BF16, "mmsa_linear_bn_act_fwd: bad out_dtype");
  cudaStream_t s = (cudaStream_t)stream;
  LbParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.X = (const bf16*)x; p.ldx = ldx; p.W = (const bf16*)w; p.ldw = ldw;
  p.bias = bias; p.gamma = gamma; p.beta = beta;
  p.running_mean = running_mean; p.running_var = running_var; p.nbt = num_batches_tracked;
  p.momentum = momentum; p.eps = eps; p.training = training; p.order = order;
  p.dropout_p = dropout_p; p.keep_mask = keep_mask; p.mask_given = mask_given; p.seed = seed; p.offset = offset;
  p.rng_state = rng_state;
  p.z = z; p.y = y; p.y_is_f32 = out_dtype == MMSA_F32; p.y_lp = (bf16*)(out_dtype == MMSA_F32 ? y_lp : nullptr);
  p.save_mean = save_mean; p.save_rstd = save_rstd;
  p.kb_total = (p.K + LB_TK - 1) / LB_TK;
  int slices = (p.kb_total + LB_KB_PER_SLICE - 1) / LB_KB_PER_SLICE;
  if (slices > LB_MAX_SLICES) slices = LB_MAX_SLICES;
  p.kb_per_slice = (p.kb_total + slices - 1) / slices;
  p.slices = (p.kb_total + p.kb_per_slice - 1) / p.kb_per_slice;
  static PerDeviceOnce attr_set;
  if (attr_set.pending()) {
    cudaError_t e = cudaFuncSetAttribute(linear_bn_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM);
    if (e != cudaSuccess) {
      set_error("mmsa: cudaFuncSetAttribute(linear_bn_act_kernel, smem=%d) failed: %s", LB_SMEM, cudaGetErrorString(e));
      return MMSA_ERR_CUDA;
    }
    attr_set.mark();
  }
  char nm[64];
  snprintf(nm, sizeof nm, "linear_bn_act_%dx%dx%d", p.M, p.N, p.K);
  ProfScope prof(nm, s, 2.0 * (double)M * (double)N * (double)K);
  dim3 grid((unsigned)((p.N + LB_COLS - 1) / LB_COLS), 1, (unsigned)p.slices);
  if (p.slices == 1) {
    linear_bn_act_kernel<<<grid, LB_THREADS, LB_SMEM, s>>>(p);
  } else {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p.slices;
    cfg.gridDim = grid; cfg.blockDim = dim3(LB_THREADS); cfg.dynamicSmemBytes = LB_SMEM; cfg.stream = s;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, linear_bn_act_kernel, p);
    if (e != cudaSuccess) {
      set_error("mmsa: cluster launch of linear_bn_act_kernel (cluster %d) failed: %s", p.slices, cudaGetErrorString(e));
      return MMSA_ERR_CUDA;
    }
  }
  MMSA_LAUNCH_CHECK("linear_bn_act_kernel");
  return MMSA_OK;
}

}  // extern "C"
