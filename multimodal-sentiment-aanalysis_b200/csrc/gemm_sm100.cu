// gemm_sm100.cu -- bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.
//
//   C[M,N] = act( alpha * A[M,Kt] * B[N,Kt]^T + bias[N] + residual[M,N] ),  fp32 accumulation
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0 (lane 0)  TMA producer      -- fills the smem ring, arrives on full[stage] with expect_tx
//   warp 1 (lane 0)  MMA issuer        -- tcgen05.mma 128 x BN x 16, tcgen05.commit -> empty[stage];
//                                         after the last k-block commit -> tmem_full[acc]
//   warps 2..5       epilogue          -- tcgen05.ld 32 lanes x 32 columns, bias/residual/activation,
//                                         vector stores; arrive tmem_empty[acc]
// Two accumulator buffers in TMEM (2 x BN columns) let the epilogue of tile i overlap the main loop
// of tile i+1.  Either operand may be K-major (row = M/N index, K contiguous: activations x,
// weights W[N,K]) or MN-major (row = K index, M/N contiguous: dY and X in wgrad, W in dgrad), so the
// three Linear products (fwd, dgrad, wgrad) need no transposed copies.  A may be split in two
// K-segments taken from two tensors (the gate's cat[q, attn], MultimodalModel.py:147).
// Split-K (for wgrad, whose output is small and whose reduction dim is B*L) writes fp32 partials.
#include <cuda.h>
#include "common.cuh"

namespace mmsa {

static constexpr int BM = 128;
static constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int kGemmThreads = 192;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (SWIZZLE_128B, Blackwell version 1)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct GemmParams {
  int M, N;                // output extents
  int kb_a1;               // k-blocks taken from A (first segment)
  int kb_total;            // total k-blocks (A + A2)
  int kb_per_split, splits;
  int tiles_m, tiles_n;
  const float* bias;
  const void* residual; long long ldr; int res_is_f32;
  void* C; long long ldc; int out_is_f32; long long split_stride;   // elements between split partials
  int act; float alpha;
};

template <int BN, bool A_MN, bool B_MN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN, A_MN, B_MN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_al + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_tiles = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (p.tiles_m * p.tiles_n);
        const int rem = tile - split * (p.tiles_m * p.tiles_n);
        const int m0 = (rem / p.tiles_n) * BM, n0 = (rem % p.tiles_n) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const bool second = kb >= p.kb_a1;
          const CUtensorMap* ma = second ? &tmA2 : &tmA;
          const int ka = (second ? kb - p.kb_a1 : kb) * BK;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, ma, m0 + c * 64, ka, full_bar(stage));
          } else {
            tma_load_2d(sa, ma, ka, m0, full_bar(stage));
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n0 + c * 64, kb * BK, full_bar(stage));
          } else {
            tma_load_2d(sb, &tmB, kb * BK, n0, full_bar(stage));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (p.tiles_m * p.tiles_n);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 32 bytes inside the 128B swizzle row; MN-major: advance 16 rows of 128B
            const uint64_t ad = A_MN ? make_sdesc(sa + k * (UMMA_K * 128), 8192, 1024) : make_sdesc(sa + k * (UMMA_K * 2), 0, 1024);
            const uint64_t bd = B_MN ? make_sdesc(sb + k * (UMMA_K * 128), 8192, 1024) : make_sdesc(sb + k * (UMMA_K * 2), 0, 1024);
            tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));                 // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / (p.tiles_m * p.tiles_n);
      const int rem = tile - split * (p.tiles_m * p.tiles_n);
      const int m0 = (rem / p.tiles_n) * BM, n0 = (rem % p.tiles_n) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          const bool full = (col0 + 32 <= p.N);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (full || col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
          if (p.residual) {
            if (p.res_is_f32) {
              const float* rp = reinterpret_cast<const float*>(p.residual) + (long long)row * p.ldr + col0;
              if (full && (p.ldr % 4 == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) { float t[4]; load_vec<float>(rp + j, t); v[j] += t[0]; v[j + 1] += t[1]; v[j + 2] += t[2]; v[j + 3] += t[3]; }
              } else {
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += rp[j];
              }
            } else {
              const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + (long long)row * p.ldr + col0;
              if (full && (p.ldr % 8 == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) { float t[8]; load_vec<bf16>(rp + j, t);
#pragma unroll
                  for (int q = 0; q < 8; ++q) v[j + q] += t[q]; }
              } else {
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += to_f(rp[j]);
              }
            }
          }
          if (p.act != MMSA_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          }
          if (p.out_is_f32) {
            float* cp = reinterpret_cast<float*>(p.C) + (long long)split * p.split_stride + (long long)row * p.ldc + col0;
            if (full && (p.ldc % 4 == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) store_vec<float>(cp + j, v + j);
            } else {
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) cp[j] = v[j];
            }
          } else {
            bf16* cp = reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc + col0;
            if (full && (p.ldc % 8 == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) store_vec<bf16>(cp + j, v + j);
            } else {
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) cp[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr)
    return nullptr;
  fn = (EncodeTiledFn)ptr;
  return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) extent `inner`, dim1 extent `outer`, row stride `ld` elements.
static bool make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld,
                     int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("mmsa: cuTensorMapEncodeTiled entry point not found"); return false; }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mmsa: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%dx%d base=%p", (int)r,
              (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer, base);
    return false;
  }
  return true;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

bool gemm_bf16_sm100_supported(const GemmDesc& d) {
  auto al16 = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  if (!al16(d.A) || !al16(d.B) || (d.A2 && !al16(d.A2))) return false;
  if (d.lda % 8 || d.ldb % 8 || (d.A2 && d.lda2 % 8)) return false;
  if (d.A2 && (d.a_mn_major || d.K % BK != 0)) return false;
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return false;
  return true;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const GemmDesc& d, int splits_req, cudaStream_t s) {
  using Cfg = GemmCfg<BN, A_MN, B_MN>;
  CUtensorMap tmA, tmA2, tmB;
  // K-major operand [rows, K]: inner = K, outer = rows, box = {64, rows_tile}
  // MN-major operand [K, mn]: inner = mn, outer = K, box = {64, 64}
  if (A_MN) { if (!make_map(&tmA, d.A, d.M, d.K, d.lda, 64, BK)) return MMSA_ERR_CUDA; }
  else      { if (!make_map(&tmA, d.A, d.K, d.M, d.lda, BK, BM)) return MMSA_ERR_CUDA; }
  if (d.A2) { if (!make_map(&tmA2, d.A2, d.K2, d.M, d.lda2, BK, BM)) return MMSA_ERR_CUDA; }
  else tmA2 = tmA;
  const int64_t Kt = d.K + (d.A2 ? d.K2 : 0);
  if (B_MN) { if (!make_map(&tmB, d.B, d.N, Kt, d.ldb, 64, BK)) return MMSA_ERR_CUDA; }
  else      { if (!make_map(&tmB, d.B, Kt, d.N, d.ldb, BK, BN)) return MMSA_ERR_CUDA; }

  GemmParams p{};
  p.M = (int)d.M; p.N = (int)d.N;
  p.kb_a1 = (int)ceil_div(d.K, BK);
  p.kb_total = p.kb_a1 + (d.A2 ? (int)ceil_div(d.K2, BK) : 0);
  p.tiles_m = (int)ceil_div(d.M, BM); p.tiles_n = (int)ceil_div(d.N, BN);
  int splits = splits_req < 1 ? 1 : splits_req;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (int)ceil_div(p.kb_total, splits);
  p.splits = (int)ceil_div(p.kb_total, p.kb_per_split);
  p.bias = d.bias; p.residual = d.residual; p.ldr = d.ldr; p.res_is_f32 = 0;
  p.C = d.C; p.ldc = d.ldc; p.out_is_f32 = (d.out_dtype == MMSA_F32);
  p.split_stride = (long long)d.M * d.ldc;
  p.act = d.act; p.alpha = d.alpha;
  int total = p.tiles_m * p.tiles_n * p.splits;
  int grid = total < num_sms() ? total : num_sms();
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("mmsa: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return MMSA_ERR_CUDA; }
    attr_set = true;
  }
  {
    ProfScope prof("gemm_tcgen05", s, 2.0 * (double)d.M * (double)d.N * (double)Kt);
    kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, s>>>(tmA, tmA2, tmB, p);
  }
  MMSA_LAUNCH_CHECK("gemm_tcgen05_kernel");
  return MMSA_OK;
}

int gemm_bf16_sm100_splits(const GemmDesc& d, int splits, cudaStream_t s);

template <bool A_MN, bool B_MN>
static int dispatch_bn(const GemmDesc& d, int splits, cudaStream_t s) {
  // tile width: 192 divides the E=768 family exactly (6.9 waves of 148 at M=32768 instead of 5.2 for 256)
  const int64_t N = d.N;
  if (N <= 64) return launch_gemm<64, A_MN, B_MN>(d, splits, s);
  if (N <= 128) return launch_gemm<128, A_MN, B_MN>(d, splits, s);
  if (N % 192 == 0) return launch_gemm<192, A_MN, B_MN>(d, splits, s);
  if (N % 256 == 0 || N > 512) return launch_gemm<256, A_MN, B_MN>(d, splits, s);
  if (N <= 192) return launch_gemm<192, A_MN, B_MN>(d, splits, s);
  return launch_gemm<256, A_MN, B_MN>(d, splits, s);
}

int gemm_bf16_sm100_splits(const GemmDesc& d, int splits, cudaStream_t s) {
  if (!gemm_bf16_sm100_supported(d)) {
    set_error("mmsa: bf16 GEMM operands must be 16B aligned with leading dims multiple of 8 (M=%lld N=%lld K=%lld lda=%lld ldb=%lld)",
              (long long)d.M, (long long)d.N, (long long)d.K, (long long)d.lda, (long long)d.ldb);
    return MMSA_ERR_ARG;
  }
  if (d.a_mn_major) return d.b_mn_major ? dispatch_bn<true, true>(d, splits, s) : dispatch_bn<true, false>(d, splits, s);
  return d.b_mn_major ? dispatch_bn<false, true>(d, splits, s) : dispatch_bn<false, false>(d, splits, s);
}

int gemm_bf16_sm100(const GemmDesc& d, cudaStream_t s) { return gemm_bf16_sm100_splits(d, 1, s); }

int gemm_num_sms() { return num_sms(); }

}  // namespace mmsa
