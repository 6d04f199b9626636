// gemm_small.cu -- latency-optimised bf16 GEMM for the [B,*] side of the path (modality-weight MLP, fusion MLP,
// heads, InfoNCE similarity block and their dgrad / wgrad: MultimodalModel.py:171-198,232-260), where one product is
// 0.005-0.13 GFLOP and what counts is the time from launch to last store, not tensor-pipe utilisation.
//
//   C[M,N] = act( alpha * A[M,Kt] * B[N,Kt]^T + bias[N] + residual[M,N] ),  bf16 operands, fp32 accumulation
//
// The tcgen05 kernel (gemm_sm100.cu) pays ~8 us per launch in fixed cost here (227 KB shared-memory carve-out, TMEM
// allocation, tensor-map fetches, a 3-role pipeline that needs tens of k-blocks to fill).  This kernel instead uses
// many tiny CTAs: 32 x 32 output tile, 4 warps of mma.sync.m16n8k16 (fp32 accumulators in registers), operands
// staged by a 4-deep cp.async ring so that a CTA's whole K-slice is in flight at once, and the reduction split over a
// thread-block cluster (<= 8 K-slices per tile, summed in slice order through distributed shared memory --
// deterministic, no workspace).  Either operand may be K-major or MN-major (ldmatrix / ldmatrix.trans), so fwd, dgrad
// and wgrad read the tensors where they lie; the wgrad bias gradient (row sums of A) comes from one extra MMA per
// k-step against a fragment of ones.
#include "common.cuh"

namespace mmsa {

namespace {

constexpr int TK = 64;
constexpr int kStages = 4;
constexpr int kThreads = 128;
constexpr int kPitchK = TK + 8;      // K-major tile row: 64 k + 8 pad elements (144 B: ldmatrix rows hit distinct banks)
constexpr int kMaxSlices = 8;

// WT = warp tile edge; the CTA tile is 2 WT x 2 WT (4 warps, 2 x 2).  WT = 16 is what ships (WT = 32 measured slower)
template <int WT>
struct SmallCfg {
  static constexpr int TM = 2 * WT, TN = 2 * WT;
  static constexpr int PITCH_MN = TM + 8;         // MN-major tile row: TM m/n + 8 pad elements (80 / 144 B)
  static constexpr int OPER_ELEMS = (TM * kPitchK > TK * PITCH_MN) ? TM * kPitchK : TK * PITCH_MN;
  static constexpr int ACC = (WT / 16) * (WT / 8) * 4;      // fp32 accumulators per thread: 8 or 32
  static constexpr int PART_FLOATS = kThreads * ACC + TM;   // partial tile + row sums, aliases the operand ring
  static constexpr int SMEM_BYTES = 2 * kStages * OPER_ELEMS * 2;
  static_assert(PART_FLOATS * 4 <= SMEM_BYTES, "partial tile fits in the ring");
};

struct SmallParams {
  int M, N, K, K2;
  const bf16* A; long long lda;
  const bf16* A2; long long lda2;
  const bf16* B; long long ldb;
  const float* bias;
  const bf16* residual; long long ldr;
  int act;
  void* C; long long ldc; int out_is_f32;
  float alpha;
  float* colsum;
  int kb_total, kb_per_slice, slices;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int clamp16(long long elems) { return elems <= 0 ? 0 : (elems >= 8 ? 16 : (int)elems * 2); }

// one (ROWS x 64 k) operand tile into a shared-memory slot; OOB rows / k are zero-filled by cp.async
template <bool MN, int ROWS>
__device__ __forceinline__ void load_tile(uint32_t slot, const bf16* __restrict__ P, long long ld, const bf16* __restrict__ P2,
                                          long long ld2, int K, int Kt, int rows_total, int r0, int kg0, int tid) {
  constexpr int PITCH_MN = ROWS + 8;
#pragma unroll
  for (int i = 0; i < ROWS * 8 / kThreads; ++i) {
    const int c = tid + i * kThreads;                 // ROWS * 8 chunks of 8 elements
    if (!MN) {
      const int row = c >> 3, kc = (c & 7) * 8, kg = kg0 + kc, r = r0 + row;
      const bf16* src = P; int bytes = 0;
      if (r < rows_total && kg < Kt) {
        if (kg < K) { src = P + (long long)r * ld + kg; bytes = clamp16(K - kg); }
        else { src = P2 + (long long)r * ld2 + (kg - K); bytes = clamp16(Kt - kg); }
      }
      cp_async16(slot + (uint32_t)(row * kPitchK + kc) * 2u, src, bytes);
    } else {
      const int krow = c / (ROWS / 8), mc = (c % (ROWS / 8)) * 8, kg = kg0 + krow, r = r0 + mc;
      const bf16* src = P; int bytes = 0;
      if (kg < Kt && r < rows_total) { src = P + (long long)kg * ld + r; bytes = clamp16(rows_total - r); }
      cp_async16(slot + (uint32_t)(krow * PITCH_MN + mc) * 2u, src, bytes);
    }
  }
}

template <int WT, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads)
gemm_small_kernel(const SmallParams p) {
  using Cfg = SmallCfg<WT>;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, PITCH_MN = Cfg::PITCH_MN;
  constexpr int MI = WT / 16, NJ = WT / 16;             // m16 blocks and n16 blocks per warp
  extern __shared__ __align__(16) uint8_t smem_small[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int slice = blockIdx.z;                               // cluster = (1, 1, slices): rank == blockIdx.z
  const int Kt = p.K + p.K2;
  const int kb0 = slice * p.kb_per_slice;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
  const int nk = kb1 - kb0;
  const int wm = (warp >> 1) * WT, wn = (warp & 1) * WT;      // warp tile WT x WT
  const bool want_cs = p.colsum != nullptr && blockIdx.x == 0 && (warp & 1) == 0;

  float acc[MI][2 * NJ][4] = {};
  float cs[MI][4] = {};
  constexpr uint32_t kSlotBytes = Cfg::OPER_ELEMS * 2;
  const uint32_t sA0 = smem_u32(smem_small), sB0 = sA0 + kStages * kSlotBytes;

  auto issue = [&](int j) {     // k-block j of this slice -> ring slot j % kStages
    const int st = j % kStages, kg0 = (kb0 + j) * TK;
    load_tile<A_MN, TM>(sA0 + st * kSlotBytes, p.A, p.lda, p.A2, p.lda2, p.K, Kt, p.M, m0, kg0, tid);
    load_tile<B_MN, TN>(sB0 + st * kSlotBytes, p.B, p.ldb, nullptr, 0, Kt, Kt, p.N, n0, kg0, tid);
  };
#pragma unroll
  for (int j = 0; j < kStages - 1; ++j) {
    if (j < nk) issue(j);
    cp_async_commit();
  }
  // per-lane ldmatrix offsets (elements) inside a slot, for k-step 0 and the warp's first 16 x 16 block
  const int l7 = lane & 7, l8 = (lane >> 3) & 1, l16 = lane >> 4;
  const uint32_t a_off = A_MN ? (uint32_t)((l16 * 8 + l7) * PITCH_MN + wm + l8 * 8) : (uint32_t)((wm + l8 * 8 + l7) * kPitchK + l16 * 8);
  const uint32_t b_off = B_MN ? (uint32_t)((l8 * 8 + l7) * PITCH_MN + wn + l16 * 8) : (uint32_t)((wn + l16 * 8 + l7) * kPitchK + l8 * 8);
  constexpr uint32_t a_step = (A_MN ? 16 * PITCH_MN : 16) * 2, b_step = (B_MN ? 16 * PITCH_MN : 16) * 2;   // bytes per k16 step
  constexpr uint32_t a_blk = (A_MN ? 16 : 16 * kPitchK) * 2, b_blk = (B_MN ? 16 : 16 * kPitchK) * 2;       // bytes per 16 rows of m / n

  for (int j = 0; j < nk; ++j) {
    cp_async_wait<kStages - 2>();
    __syncthreads();                                  // k-block j has landed; everyone is done with slot (j-1) % kStages
    if (j + kStages - 1 < nk) issue(j + kStages - 1);
    cp_async_commit();
    const uint32_t aS = sA0 + (j % kStages) * kSlotBytes + a_off * 2u;
    const uint32_t bS = sB0 + (j % kStages) * kSlotBytes + b_off * 2u;
#pragma unroll
    for (int ks = 0; ks < TK / 16; ++ks) {
      uint32_t a[MI][4], b[NJ][4];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        if (A_MN) ldsm_x4_t(aS + ks * a_step + mi * a_blk, a[mi]); else ldsm_x4(aS + ks * a_step + mi * a_blk, a[mi]);
      }
#pragma unroll
      for (int nj = 0; nj < NJ; ++nj) {
        if (B_MN) ldsm_x4_t(bS + ks * b_step + nj * b_blk, b[nj]); else ldsm_x4(bS + ks * b_step + nj * b_blk, b[nj]);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
        for (int nj = 0; nj < NJ; ++nj) {
          mma_bf16(acc[mi][2 * nj], a[mi], b[nj][0], b[nj][1]);
          mma_bf16(acc[mi][2 * nj + 1], a[mi], b[nj][2], b[nj][3]);
        }
        if (want_cs) mma_bf16(cs[mi], a[mi], 0x3F803F80u, 0x3F803F80u);      // row sums of A: B fragment of bf16 ones
      }
    }
  }
  cp_async_wait<0>();

  const int g = lane >> 2, q = (lane & 3) * 2;
  if (p.slices > 1) {
    // ---- cluster sum: every CTA parks its partial (over the drained ring); rank 0 adds the others in slice
    //      order through distributed shared memory ----
    __syncthreads();                                  // all warps are done reading the ring
    float* part = reinterpret_cast<float*>(smem_small);
    float4* mine = reinterpret_cast<float4*>(part) + tid;          // [group][thread]: conflict-free 16-byte rows
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int nt = 0; nt < 2 * NJ; ++nt)
        mine[(mi * 2 * NJ + nt) * kThreads] = make_float4(acc[mi][nt][0], acc[mi][nt][1], acc[mi][nt][2], acc[mi][nt][3]);
    float* part_cs = part + kThreads * Cfg::ACC;
    if (want_cs && (lane & 3) == 0) {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) { part_cs[wm + mi * 16 + g] = cs[mi][0]; part_cs[wm + mi * 16 + g + 8] = cs[mi][2]; }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (slice == 0) {
      const uint32_t my = smem_u32(mine);
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
        for (int nt = 0; nt < 2 * NJ; ++nt) {
          const uint32_t off = my + (uint32_t)((mi * 2 * NJ + nt) * kThreads * 16);
          float4 u[kMaxSlices - 1];
#pragma unroll
          for (int r = 1; r < kMaxSlices; ++r) {      // all remote loads of this group in flight before the first add
            if (r < p.slices) {
              uint32_t ra;
              asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(off), "r"(r));
              asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(u[r - 1].x), "=f"(u[r - 1].y), "=f"(u[r - 1].z), "=f"(u[r - 1].w) : "r"(ra) : "memory");
            }
          }
#pragma unroll
          for (int r = 1; r < kMaxSlices; ++r) {      // slice order: deterministic
            if (r < p.slices) { acc[mi][nt][0] += u[r - 1].x; acc[mi][nt][1] += u[r - 1].y; acc[mi][nt][2] += u[r - 1].z; acc[mi][nt][3] += u[r - 1].w; }
          }
        }
      }
      if (want_cs && (lane & 3) == 0) {
        const uint32_t my_cs = smem_u32(part_cs + wm + g);
        for (int r = 1; r < p.slices; ++r) {
          uint32_t rc;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rc) : "r"(my_cs), "r"(r));
#pragma unroll
          for (int mi = 0; mi < MI; ++mi) {
            float x, y;
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x) : "r"(rc + (uint32_t)(mi * 64)) : "memory");
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(y) : "r"(rc + (uint32_t)(mi * 64 + 32)) : "memory");
            cs[mi][0] += x; cs[mi][2] += y;
          }
        }
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' smem stays valid until read
    if (slice != 0) return;
  }

  // ---- epilogue: alpha, bias, residual, activation, store (thread owns rows g, g+8 and column pairs) ----
  if (want_cs && (lane & 3) == 0) {
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const int r = m0 + wm + mi * 16 + g;
      if (r < p.M) p.colsum[r] = cs[mi][0];
      if (r + 8 < p.M) p.colsum[r + 8] = cs[mi][2];
    }
  }
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
    for (int nt = 0; nt < 2 * NJ; ++nt) {
      const int col = n0 + wn + nt * 8 + q;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + wm + mi * 16 + g + h * 8;
        if (row >= p.M) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = col + e;
          if (c >= p.N) continue;
          float v = acc[mi][nt][h * 2 + e] * p.alpha;
          if (p.bias) v += __ldg(p.bias + c);
          if (p.residual) v += to_f(p.residual[(long long)row * p.ldr + c]);
          v = apply_act(v, p.act);
          if (p.out_is_f32) reinterpret_cast<float*>(p.C)[(long long)row * p.ldc + c] = v;
          else reinterpret_cast<bf16*>(p.C)[(long long)row * p.ldc + c] = __float2bfloat16_rn(v);
        }
      }
    }
  }
}

template <int WT, bool A_MN, bool B_MN>
int launch_small(const GemmDesc& d, cudaStream_t s) {
  using Cfg = SmallCfg<WT>;
  SmallParams p{};
  p.M = (int)d.M; p.N = (int)d.N; p.K = (int)d.K; p.K2 = d.A2 ? (int)d.K2 : 0;
  p.A = (const bf16*)d.A; p.lda = d.lda; p.A2 = (const bf16*)d.A2; p.lda2 = d.lda2;
  p.B = (const bf16*)d.B; p.ldb = d.ldb; p.bias = d.bias; p.residual = (const bf16*)d.residual; p.ldr = d.ldr;
  p.act = d.act; p.C = d.C; p.ldc = d.ldc; p.out_is_f32 = d.out_dtype == MMSA_F32; p.alpha = d.alpha; p.colsum = d.colsum;
  const int Kt = p.K + p.K2;
  p.kb_total = (Kt + TK - 1) / TK;
  int slices = (p.kb_total + 3) / 4;                   // <= 4 k-blocks per CTA: the whole slice fits the cp.async ring
  if (slices > kMaxSlices) slices = kMaxSlices;
  p.kb_per_slice = (p.kb_total + slices - 1) / slices;
  p.slices = (p.kb_total + p.kb_per_slice - 1) / p.kb_per_slice;
  dim3 grid((unsigned)((p.N + Cfg::TN - 1) / Cfg::TN), (unsigned)((p.M + Cfg::TM - 1) / Cfg::TM), (unsigned)p.slices);
  auto kern = gemm_small_kernel<WT, A_MN, B_MN>;
  if (Cfg::SMEM_BYTES > 48 * 1024) {      // per device, so not cached in a static (WT = 16 stays below the opt-in limit)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("mmsa: cudaFuncSetAttribute(gemm_small_kernel, smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return MMSA_ERR_CUDA; }
  }
  char nm[64];
  snprintf(nm, sizeof nm, "gemm_small_%c%c_%dx%dx%d", A_MN ? 'm' : 'k', B_MN ? 'm' : 'k', p.M, p.N, Kt);
  ProfScope prof(nm, s, 2.0 * (double)d.M * (double)d.N * (double)Kt);
  if (p.slices == 1) {
    kern<<<grid, kThreads, Cfg::SMEM_BYTES, s>>>(p);
  } else {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p.slices;
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = s;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) { set_error("mmsa: cluster launch of gemm_small_kernel (cluster %d) failed: %s", p.slices, cudaGetErrorString(e)); return MMSA_ERR_CUDA; }
  }
  MMSA_LAUNCH_CHECK("gemm_small_kernel");
  return MMSA_OK;
}

template <int WT>
int dispatch_small(const GemmDesc& d, cudaStream_t s) {
  if (d.a_mn_major) return d.b_mn_major ? launch_small<WT, true, true>(d, s) : launch_small<WT, true, false>(d, s);
  return d.b_mn_major ? launch_small<WT, false, true>(d, s) : launch_small<WT, false, false>(d, s);
}

}  // namespace

// Products small enough that launch-to-finish latency, not throughput, is what matters.  Measured in-graph on B200
// (scripts/small_gemm_probe.py): 256x128x256 4.6 us vs 5.7 us on the tcgen05 kernel, 256x64x1536 6.3 vs 9.5,
// 256x768x64 3.7 vs 5.5; from ~0.13 GFLOP up (fusion.0: 256x256x2304, 11.6 vs 10.2 us) the operand re-reads of 32 x 32
// tiles through L2 (M N K / 16 bytes) outweigh the extra CTAs, and 64 x 64 tiles are slower still (fewer, longer
// CTAs: 16.5 us), so those stay on the tcgen05 split-K cluster path.
bool gemm_bf16_small_ok(const GemmDesc& d) {
  auto al16 = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return false;
  if (!al16(d.A) || !al16(d.B) || (d.A2 && !al16(d.A2))) return false;
  if (d.lda % 8 || d.ldb % 8 || (d.A2 && d.lda2 % 8)) return false;
  if (d.A2 && (d.a_mn_major || d.K % 8 != 0)) return false;
  const int64_t Kt = d.K + (d.A2 ? d.K2 : 0);
  if (d.M > 2048 || d.N > 65535 * 32LL || Kt > (1 << 20)) return false;
  return (double)d.M * (double)d.N * (double)Kt < 67108864.0;        // < 0.13 GFLOP
}

int gemm_bf16_small(const GemmDesc& d, cudaStream_t s) { return dispatch_small<16>(d, s); }

}  // namespace mmsa
