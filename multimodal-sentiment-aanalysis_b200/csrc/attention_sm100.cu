// attention_sm100.cu -- bf16 multi-head attention core (d_h = 64) on the 5th-gen tensor cores, forward and backward.
//   S = Q K^T (tcgen05.mma, accumulator in TMEM) -> softmax in registers (one thread per query row, no shuffles)
//   -> P (bf16) staged in shared memory as the A operand -> O = P V (tcgen05.mma, accumulator in TMEM).
// Replaces torch's bmm + softmax + bmm of F.multi_head_attention_forward (need_weights branch, reached from
// MultimodalModel.py:139-143): q is scaled by 1/sqrt(d_h) before the product there; 1/8 is a power of two, so scaling
// the fp32 scores instead is bit-identical.  The B*H*Lq*Lk probability matrix never leaves the SM; only the per-row
// log-sum-exp is saved for the backward.
//
// Stand-alone the core is HBM-bound (AI 28-45 FLOP/B), so the kernel is organised around keeping loads in flight:
// persistent CTAs (two per SM) walk (sample, head, query-tile) units; warp 0 streams the Q tile (128 x 64) and the
// K / V tiles (64 keys x 64) of the NEXT units by TMA into mbarrier rings while warp 1 issues the MMAs of the current
// unit and warps 2-5 (128 threads = 128 TMEM lanes = 128 query rows) do the softmax and the output rows.
// Keys are walked in tiles of 64 with an online softmax; Lk = 49 (image regions) is one masked tile.  With more than
// one key tile the O accumulator stays in TMEM (accumulating MMAs) and is rescaled in place when a row maximum moves.
#include <cuda.h>
#include "common.cuh"

namespace mmsa {

bool tc_make_map3_bf16(CUtensorMap* map, const void* base, int64_t cols, int64_t L, int64_t B, int64_t ld, int box_inner,
                       int box_rows);

namespace {

constexpr int TQ = 128, TK = 64, HD = 64;
constexpr int QST = 2, KST = 2;      // ring depths: 2 x (16 + 16) KB + P + O staging = 97 KB -> two CTAs per SM
constexpr int Q_BYTES = TQ * HD * 2, K_BYTES = TK * HD * 2, P_BYTES = TQ * TK * 2;
constexpr int O_BYTES = TQ * HD * 2;    // output staging: one 32-row x 128 B slab per softmax warp, stored by TMA
constexpr int SMEM_Q = 0, SMEM_KV = SMEM_Q + QST * Q_BYTES, SMEM_P = SMEM_KV + KST * 2 * K_BYTES,
              SMEM_O = SMEM_P + P_BYTES, SMEM_BAR = SMEM_O + O_BYTES, SMEM_TOTAL = SMEM_BAR + 256 + 1024 /*align*/;
constexpr int kThreads = 192;
constexpr uint32_t TMEM_COLS = 128;      // S: columns [0,64), O: columns [64,128)

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
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte chunk `c` (0..7) of row `r` of a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(uint32_t tile, int r, int c) { return tile + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ float fast_exp2(float x) {      // MUFU.EX2; exp2(-inf) = 0, denormal results flush to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h2);
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_128B (same encoding as gemm_sm100.cu)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AttnParams {
  int H, Lq, Lk, nq, nkv;
  long long units;
  bf16* o; long long ldo;
  float* lse;
  float scale, scale_log2;     // 1/sqrt(d_h), and the same times log2(e)
};

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + SMEM_BAR;
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_empty = [&](int s) { return bar0 + 8u * (QST + s); };
  auto kv_full = [&](int s) { return bar0 + 8u * (2 * QST + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (2 * QST + KST + s); };
  const uint32_t s_full = bar0 + 8u * (2 * QST + 2 * KST), s_free = s_full + 8, p_full = s_full + 16,
                 o_full = s_full + 24, o_free = s_full + 32;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_al + SMEM_BAR + 8 * (2 * QST + 2 * KST + 5));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    for (int s = 0; s < QST; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int s = 0; s < KST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64u;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int qs = 0, ks = 0; uint32_t qph = 0, kph = 0;
      for (long long unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const int qt = (int)(unit % p.nq);
        const long long bh = unit / p.nq;
        const int h = (int)(bh % p.H);
        const long long b = bh / p.H;
        mbar_wait(q_empty(qs), qph ^ 1u);
        mbar_expect_tx(q_full(qs), Q_BYTES);
        tma_load_3d(base + SMEM_Q + qs * Q_BYTES, &tmQ, h * HD, qt * TQ, (int)b, q_full(qs));
        if (++qs == QST) { qs = 0; qph ^= 1u; }
        for (int j = 0; j < p.nkv; ++j) {
          mbar_wait(kv_empty(ks), kph ^ 1u);
          mbar_expect_tx(kv_full(ks), 2 * K_BYTES);
          const uint32_t dst = base + SMEM_KV + ks * 2 * K_BYTES;
          tma_load_3d(dst, &tmK, h * HD, j * TK, (int)b, kv_full(ks));
          tma_load_3d(dst + K_BYTES, &tmV, h * HD, j * TK, (int)b, kv_full(ks));
          if (++ks == KST) { ks = 0; kph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      // S = Q K^T: A = Q (K-major), B = K (K-major), M = 128, N = 64.   O = P V: A = P (K-major), B = V (MN-major).
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      // Two cursors over the (unit, key-tile) steps: S of step t+1 is issued as soon as the softmax warps have READ
      // S of step t (s_free), i.e. while they are still exponentiating it, so the next scores are already waiting
      // in TMEM when they come back; P V of step t follows once P is staged.
      struct Cur { long long unit; int j, qs, ks; uint32_t qph, kph; };
      Cur a{(long long)blockIdx.x, 0, 0, 0, 0u, 0u}, c = a;
      uint32_t sfree_ph = 0, pfull_ph = 0, ofree_ph = 0;
      auto advance = [&](Cur& x) {
        if (++x.ks == KST) { x.ks = 0; x.kph ^= 1u; }
        if (++x.j == p.nkv) { x.j = 0; x.unit += gridDim.x; if (++x.qs == QST) { x.qs = 0; x.qph ^= 1u; } }
      };
      auto issue_s = [&](Cur& x) {
        if (x.j == 0) mbar_wait(q_full(x.qs), x.qph);
        mbar_wait(kv_full(x.ks), x.kph);
        mbar_wait(s_free, sfree_ph ^ 1u); sfree_ph ^= 1u;        // the softmax warps have read the previous S
        tc_fence_after();
        const uint32_t sq = base + SMEM_Q + x.qs * Q_BYTES, sk = base + SMEM_KV + x.ks * 2 * K_BYTES;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc_mma(tmem_S, make_sdesc(sq + k * 32, 0, 1024), make_sdesc(sk + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(s_full);
      };
      if (a.unit < p.units) { issue_s(a); advance(a); }
      while (c.unit < p.units) {
        if (a.unit < p.units) { issue_s(a); advance(a); }
        mbar_wait(p_full, pfull_ph); pfull_ph ^= 1u;             // P staged (and O rescaled) by the softmax warps
        if (c.j == 0) { mbar_wait(o_free, ofree_ph ^ 1u); ofree_ph ^= 1u; }   // previous unit's O has been read out
        tc_fence_after();
        const uint32_t sp = base + SMEM_P, sv = base + SMEM_KV + c.ks * 2 * K_BYTES + K_BYTES;
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)
          tc_mma(tmem_O, make_sdesc(sp + k * 32, 0, 1024), make_sdesc(sv + k * 2048, 8192, 1024), idesc_o,
                 (c.j > 0 || k > 0) ? 1u : 0u);
        tc_commit(kv_empty(c.ks));
        tc_commit(o_full);
        if (c.j == p.nkv - 1) tc_commit(q_empty(c.qs));
        advance(c);
      }
    }
  } else {
    // ===================== softmax / output warps (2..5): thread <-> query row =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                        // row of the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t sfull_ph = 0, ofull_ph = 0;
    const uint32_t p_tile = base + SMEM_P;
    const uint32_t o_slab = base + SMEM_O + (uint32_t)quad * 4096u;      // this warp's 32 x 128 B output slab
    // The output of a unit (O / l through this warp's staging slab and a TMA store -- rows past Lq are clipped by the
    // 3-D tensor map -- and the row log-sum-exp) is DEFERRED into the next unit, after its scores have been
    // exponentiated: the P V product is then never waited for.
    struct FwdPending { bool on; float m, l; int h, b, qt; } pend{false, 0.f, 1.f, 0, 0, 0};
    auto finish_prev = [&]() {
      if (!pend.on) return;
      pend.on = false;
      mbar_wait(o_full, ofull_ph); ofull_ph ^= 1u;
      tc_fence_after();
      const float inv = 1.f / pend.l;
      if (lane == 0) tma_store_wait_read0();               // the previous store of this slab has been read out
      __syncwarp();
#pragma unroll 1
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t orr[32];
        tmem_ld32(tmem_O + lane_base + cb * 32u, orr);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            w[t] = pack_bf16(__uint_as_float(orr[c * 8 + 2 * t]) * inv, __uint_as_float(orr[c * 8 + 2 * t + 1]) * inv);
          sts128(sw128(o_slab, lane, cb * 4 + c), w[0], w[1], w[2], w[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_3d(&tmO, o_slab, pend.h * HD, pend.qt * TQ + quad * 32, pend.b); tma_store_commit(); }
      const int i = pend.qt * TQ + r;
      if (i < p.Lq) p.lse[((long long)pend.b * p.H + pend.h) * (long long)p.Lq + i] = pend.m * p.scale + logf(pend.l);
    };
    for (long long unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int qt = (int)(unit % p.nq);
      const long long bh = unit / p.nq;
      const int h = (int)(bh % p.H);
      const long long b = bh / p.H;
      // a warp whose 32 rows all lie past the end of the sample (Lq = 49: warps 2 and 3) only keeps the barrier
      // protocol going; its P rows stay stale, which only feeds accumulator rows nobody reads
      const bool live = qt * TQ + quad * 32 < p.Lq;
      if (!live) {
        finish_prev();
        for (int j = 0; j < p.nkv; ++j) {
          mbar_wait(s_full, sfull_ph); sfull_ph ^= 1u;
          if (lane == 0) mbar_arrive(s_free);
          if (j > 0) { mbar_wait(o_full, ofull_ph); ofull_ph ^= 1u; }
          if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(o_full, ofull_ph); ofull_ph ^= 1u;
        if (lane == 0) mbar_arrive(o_free);
        continue;
      }
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < p.nkv; ++j) {
        mbar_wait(s_full, sfull_ph); sfull_ph ^= 1u;
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld32(tmem_S + lane_base, sr);
        tmem_ld32(tmem_S + lane_base + 32u, sr + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        const int valid = min(TK, p.Lk - j * TK);
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; ++c) if (c < valid) mx = fmaxf(mx, __uint_as_float(sr[c]));
        const float m_new = fmaxf(m, mx);
        const float alpha = fast_exp2((m - m_new) * p.scale_log2);     // first tile: exp2(-inf) = 0
        float rowsum = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float p0 = c < valid ? fast_exp2((__uint_as_float(sr[c]) - m_new) * p.scale_log2) : 0.f;
          const float p1 = c + 1 < valid ? fast_exp2((__uint_as_float(sr[c + 1]) - m_new) * p.scale_log2) : 0.f;
          rowsum += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
        if (j > 0) {
          // P V of the previous tile has retired: P may be overwritten, and O (accumulated so far against the old
          // row maxima) is rescaled in place where a maximum moved
          mbar_wait(o_full, ofull_ph); ofull_ph ^= 1u;
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
            for (int cb = 0; cb < 2; ++cb) {
              uint32_t orr[32];
              tmem_ld32(tmem_O + lane_base + cb * 32u, orr);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * alpha);
              tmem_st32(tmem_O + lane_base + cb * 32u, orr);
            }
            tmem_st_wait();
          }
        } else {
          finish_prev();       // the previous unit's O: its P V ran while this unit's scores were exponentiated
        }
        l = l * alpha + rowsum;
        m = m_new;
        // stage P (bf16) as the K-major, 128B-swizzled A operand
#pragma unroll
        for (int c = 0; c < 8; ++c) sts128(sw128(p_tile, r, c), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      pend = FwdPending{true, m, l, h, (int)b, qt};
    }
    finish_prev();
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- backward
// One fused kernel per direction: Q, K, V, O, dO and the row log-sum-exp are read once, dQ, dK, dV written once; the
// probabilities are recomputed tile by tile and never leave the SM, and delta = rowsum(dO * O) is formed from the
// staged tiles (no separate kernel, no HBM round trip).  Per (query tile of 128, key tile of 64):
//   S = Q K^T, dP = dO V^T                         (tcgen05.mma, M = 128, N = 64)
//   P = exp(S/8 - lse), dS = P * (dP - delta) / 8   (one thread per query row; bf16 into shared memory, side by side)
//   dQ = dS K                                       (A = dS K-major, B = K MN-major)
//   [dV ; dK] = [P | dS]^T [dO | Q]                 (ONE M = 128, N = 128 MMA chain over the 128 query rows: the
//                                                    diagonal blocks are dV (TMEM lanes 0-63, columns 0-63) and dK
//                                                    (lanes 64-127, columns 64-127); every product keeps M = 128)
// Works when one side fits a single tile (all BASELINE shapes: 49 image regions): with one key tile the kernel walks
// the query tiles of a (sample, head) and accumulates dK/dV in TMEM; with one query tile it walks the key tiles and
// accumulates dQ.  Other shapes use the mma.sync engine (attention_mma.cu).
constexpr int B_QR_BYTES = 3 * Q_BYTES;       // dO | Q | O tiles of one query tile
constexpr int B_KR_BYTES = 2 * K_BYTES;       // K | V tiles of one key tile
constexpr int B_PS_BYTES = 2 * P_BYTES;       // P | dS, two 128 x 128 B atoms
constexpr int B_SMEM_QR = 0, B_SMEM_KR = B_SMEM_QR + 2 * B_QR_BYTES, B_SMEM_PS = B_SMEM_KR + 2 * B_KR_BYTES,
              B_SMEM_DQ = B_SMEM_PS + B_PS_BYTES, B_SMEM_DKV = B_SMEM_DQ + 16384, B_SMEM_BAR = B_SMEM_DKV + 16384,
              B_SMEM_TOTAL = B_SMEM_BAR + 256 + 1024;

struct AttnBwdParams {
  int H, Lq, Lk, nsteps, q_outer;    // q_outer: one key tile, steps walk query tiles; else one query tile, steps walk key tiles
  long long units;                   // B * H
  const float* lse;
  float scale, scale_log2;
};

constexpr int kThreadsBwd = 64 + 256;     // TMA warp, MMA warp, 8 compute warps (two per TMEM lane quadrant)

__global__ void __launch_bounds__(kThreadsBwd, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmdQ,
                   const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + B_SMEM_BAR;
  auto qr_full = [&](int s) { return bar0 + 8u * s; };
  auto qr_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto kr_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto kr_empty = [&](int s) { return bar0 + 8u * (6 + s); };
  const uint32_t sdp_full = bar0 + 64, sdp_free = bar0 + 72, ps_full = bar0 + 80, ps_free = bar0 + 88,
                 dq_full = bar0 + 96, dq_free = bar0 + 104, dkv_full = bar0 + 112, dkv_free = bar0 + 120;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_al + B_SMEM_BAR + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsteps = p.nsteps;
  const bool q_outer = p.q_outer != 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(qr_full(s), 1); mbar_init(qr_empty(s), 1); mbar_init(kr_full(s), 1); mbar_init(kr_empty(s), 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_free, 8); mbar_init(ps_full, 8); mbar_init(ps_free, 1);
    mbar_init(dq_full, 1); mbar_init(dq_free, 8); mbar_init(dkv_full, 1); mbar_init(dkv_free, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64u, tmem_dQ = tmem_base + 128u, tmem_dKV = tmem_base + 192u;

  // every role walks the same (unit, step) sequence; a cursor tracks which query-side / key-side tile set is current
  struct Cur { long long unit; int st; int qn, kn, qi, ki; };
  auto enter = [&](Cur& x) {              // call at the start of a step: picks up a new tile set where the step needs one
    if (q_outer || x.st == 0) x.qi = x.qn++;
    if (!q_outer || x.st == 0) x.ki = x.kn++;
  };
  auto advance = [&](Cur& x) { if (++x.st == nsteps) { x.st = 0; x.unit += gridDim.x; } };
  auto new_qr = [&](const Cur& x) { return q_outer || x.st == 0; };
  auto new_kr = [&](const Cur& x) { return !q_outer || x.st == 0; };
  auto last_qr = [&](const Cur& x) { return q_outer || x.st == nsteps - 1; };
  auto last_kr = [&](const Cur& x) { return !q_outer || x.st == nsteps - 1; };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      Cur x{(long long)blockIdx.x, 0, 0, 0, 0, 0};
      while (x.unit < p.units) {
        const int h = (int)(x.unit % p.H);
        const int b = (int)(x.unit / p.H);
        const int qt = q_outer ? x.st : 0, kt = q_outer ? 0 : x.st;
        enter(x);
        if (new_qr(x)) {
          const int s = x.qi & 1;
          mbar_wait(qr_empty(s), (((uint32_t)x.qi >> 1) & 1u) ^ 1u);
          mbar_expect_tx(qr_full(s), B_QR_BYTES);
          const uint32_t dst = base + B_SMEM_QR + s * B_QR_BYTES;
          tma_load_3d(dst, &tmdO, h * HD, qt * TQ, b, qr_full(s));
          tma_load_3d(dst + Q_BYTES, &tmQ, h * HD, qt * TQ, b, qr_full(s));
          tma_load_3d(dst + 2 * Q_BYTES, &tmO, h * HD, qt * TQ, b, qr_full(s));
        }
        if (new_kr(x)) {
          const int s = x.ki & 1;
          mbar_wait(kr_empty(s), (((uint32_t)x.ki >> 1) & 1u) ^ 1u);
          mbar_expect_tx(kr_full(s), B_KR_BYTES);
          const uint32_t dst = base + B_SMEM_KR + s * B_KR_BYTES;
          tma_load_3d(dst, &tmK, h * HD, kt * TK, b, kr_full(s));
          tma_load_3d(dst + K_BYTES, &tmV, h * HD, kt * TK, b, kr_full(s));
        }
        advance(x);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      constexpr uint32_t idesc_dq = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      constexpr uint32_t idesc_dkv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      Cur a{(long long)blockIdx.x, 0, 0, 0, 0, 0}, c = a;
      uint32_t sdpfree_ph = 0, psfull_ph = 0, dqfree_ph = 0, dkvfree_ph = 0;
      auto issue_sdp = [&](Cur& x) {       // S = Q K^T and dP = dO V^T of step x
        enter(x);
        if (new_qr(x)) mbar_wait(qr_full(x.qi & 1), ((uint32_t)x.qi >> 1) & 1u);
        if (new_kr(x)) mbar_wait(kr_full(x.ki & 1), ((uint32_t)x.ki >> 1) & 1u);
        mbar_wait(sdp_free, sdpfree_ph ^ 1u); sdpfree_ph ^= 1u;      // the previous S / dP have been read out
        tc_fence_after();
        const uint32_t qr = base + B_SMEM_QR + (x.qi & 1) * B_QR_BYTES, kr = base + B_SMEM_KR + (x.ki & 1) * B_KR_BYTES;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc_mma(tmem_S, make_sdesc(qr + Q_BYTES + k * 32, 0, 1024), make_sdesc(kr + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc_mma(tmem_dP, make_sdesc(qr + k * 32, 0, 1024), make_sdesc(kr + K_BYTES + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(sdp_full);
      };
      if (a.unit < p.units) { issue_sdp(a); advance(a); }
      while (c.unit < p.units) {
        if (a.unit < p.units) { issue_sdp(a); advance(a); }      // next step's scores while this step's P / dS are formed
        enter(c);
        const bool dq_fresh = q_outer || c.st == 0, dq_done = q_outer || c.st == nsteps - 1;
        const bool dkv_fresh = !q_outer || c.st == 0, dkv_done = !q_outer || c.st == nsteps - 1;
        mbar_wait(ps_full, psfull_ph); psfull_ph ^= 1u;
        if (dq_fresh) { mbar_wait(dq_free, dqfree_ph ^ 1u); dqfree_ph ^= 1u; }
        if (dkv_fresh) { mbar_wait(dkv_free, dkvfree_ph ^ 1u); dkvfree_ph ^= 1u; }
        tc_fence_after();
        const uint32_t qr = base + B_SMEM_QR + (c.qi & 1) * B_QR_BYTES, kr = base + B_SMEM_KR + (c.ki & 1) * B_KR_BYTES;
        const uint32_t ps = base + B_SMEM_PS;
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)        // dQ (+)= dS K: reduction over the 64 keys
          tc_mma(tmem_dQ, make_sdesc(ps + P_BYTES + k * 32, 0, 1024), make_sdesc(kr + k * 2048, 8192, 1024), idesc_dq,
                 (!dq_fresh || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < TQ / 16; ++k)        // [dV ; dK] (+)= [P | dS]^T [dO | Q]: reduction over the 128 query rows
          tc_mma(tmem_dKV, make_sdesc(ps + k * 2048, P_BYTES, 1024), make_sdesc(qr + k * 2048, Q_BYTES, 1024), idesc_dkv,
                 (!dkv_fresh || k > 0) ? 1u : 0u);
        tc_commit(ps_free);
        if (last_qr(c)) tc_commit(qr_empty(c.qi & 1));
        if (last_kr(c)) tc_commit(kr_empty(c.ki & 1));
        if (dq_done) tc_commit(dq_full);
        if (dkv_done) tc_commit(dkv_full);
        advance(c);
      }
    }
  } else {
    // ===================== compute / output warps (2..9) =====================
    // Two warps per TMEM lane quadrant: thread (row r, half hh) owns columns [32 hh, 32 hh + 32) of row r of every
    // accumulator -- half the exponentials / packing per thread, and twice the warps to hide TMEM and MUFU latency.
    const int quad = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t ps = base + B_SMEM_PS;
    const uint32_t dq_slab = base + B_SMEM_DQ + (uint32_t)quad * 4096u, dkv_slab = base + B_SMEM_DKV + (uint32_t)quad * 4096u;
    const float log2e = 1.4426950408889634f;
    const bool issuer = hh == 0 && lane == 0;           // issues the TMA stores of this quadrant's slabs
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };   // the two warps of a quadrant
    uint32_t sdpfull_ph = 0, psfree_ph = 0, dqfull_ph = 0, dkvfull_ph = 0;
    Cur x{(long long)blockIdx.x, 0, 0, 0, 0, 0};
    float lse_l2 = 0.f, delta = 0.f;
    bool row_valid = false;
    bool ps_dirty = true;      // this thread's P / dS chunks may hold non-zero (or uninitialised) data
    // The read-out of a step's dQ / dK / dV is DEFERRED until the next step's P / dS have been staged: the second MMA
    // chain of step t (dQ, [dV;dK]) then runs while these warps already exponentiate step t+1, instead of being waited for.
    struct Pending { bool dq, dkv; int h, b, qt, kt; } pend{false, false, 0, 0, 0, 0};
    auto readout = [&]() {
      if (pend.dq) {
        mbar_wait(dq_full, dqfull_ph); dqfull_ph ^= 1u;
        tc_fence_after();
        if (issuer) tma_store_wait_read0();
        pair_sync();
        {
          uint32_t orr[32];
          tmem_ld32(tmem_dQ + lane_base + hh * 32u, orr);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_free);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            sts128(sw128(dq_slab, lane, hh * 4 + c4),
                   pack_bf16(__uint_as_float(orr[c4 * 8]), __uint_as_float(orr[c4 * 8 + 1])),
                   pack_bf16(__uint_as_float(orr[c4 * 8 + 2]), __uint_as_float(orr[c4 * 8 + 3])),
                   pack_bf16(__uint_as_float(orr[c4 * 8 + 4]), __uint_as_float(orr[c4 * 8 + 5])),
                   pack_bf16(__uint_as_float(orr[c4 * 8 + 6]), __uint_as_float(orr[c4 * 8 + 7])));
        }
        fence_async_smem();
        pair_sync();
        if (issuer) { tma_store_3d(&tmdQ, dq_slab, pend.h * HD, pend.qt * TQ + quad * 32, pend.b); tma_store_commit(); }
      }
      if (pend.dkv) {
        mbar_wait(dkv_full, dkvfull_ph); dkvfull_ph ^= 1u;
        tc_fence_after();
        if (issuer) tma_store_wait_read0();
        pair_sync();
        // TMEM lanes 0-63 x columns 0-63 hold dV, lanes 64-127 x columns 64-127 hold dK (already scaled through dS)
        const uint32_t col0 = quad < 2 ? 0u : 64u;
        {
          uint32_t orr[32];
          tmem_ld32(tmem_dKV + lane_base + col0 + hh * 32u, orr);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dkv_free);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            sts128(sw128(dkv_slab, lane, hh * 4 + c4),
                   pack_bf16(__uint_as_float(orr[c4 * 8]), __uint_as_float(orr[c4 * 8 + 1])),
                   pack_bf16(__uint_as_float(orr[c4 * 8 + 2]), __uint_as_float(orr[c4 * 8 + 3])),
                   pack_bf16(__uint_as_float(orr[c4 * 8 + 4]), __uint_as_float(orr[c4 * 8 + 5])),
                   pack_bf16(__uint_as_float(orr[c4 * 8 + 6]), __uint_as_float(orr[c4 * 8 + 7])));
        }
        fence_async_smem();
        pair_sync();
        if (issuer) {
          tma_store_3d(quad < 2 ? &tmdV : &tmdK, dkv_slab, pend.h * HD, pend.kt * TK + (quad & 1) * 32, pend.b);
          tma_store_commit();
        }
      }
      pend.dq = pend.dkv = false;
    };
    while (x.unit < p.units) {
      const int h = (int)(x.unit % p.H);
      const int b = (int)(x.unit / p.H);
      const int qt = q_outer ? x.st : 0, kt = q_outer ? 0 : x.st;
      enter(x);
      const bool dq_done = q_outer || x.st == nsteps - 1, dkv_done = !q_outer || x.st == nsteps - 1;
      if (new_qr(x)) {
        // row statistics of this query tile: lse from HBM, delta = <dO_i, O_i> from the staged (swizzled) tiles
        mbar_wait(qr_full(x.qi & 1), ((uint32_t)x.qi >> 1) & 1u);
        const uint32_t t_do = base + B_SMEM_QR + (x.qi & 1) * B_QR_BYTES, t_o = t_do + 2 * Q_BYTES;
        const int i = qt * TQ + r;
        row_valid = i < p.Lq;
        lse_l2 = row_valid ? p.lse[((long long)b * p.H + h) * p.Lq + i] * log2e : 0.f;
        float acc = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 u = lds128(sw128(t_do, r, c8)), w = lds128(sw128(t_o, r, c8));
          const __nv_bfloat162* uh = reinterpret_cast<const __nv_bfloat162*>(&u);
          const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
          for (int t = 0; t < 4; ++t) { const float2 f = __bfloat1622float2(uh[t]), g = __bfloat1622float2(wh[t]); acc += f.x * g.x + f.y * g.y; }
        }
        delta = acc;
      }
      mbar_wait(sdp_full, sdpfull_ph); sdpfull_ph ^= 1u;
      tc_fence_after();
      if (qt * TQ + quad * 32 >= p.Lq) {
        // every query row of this warp lies past the end of the sample (Lq = 49: the two upper lane quadrants): no
        // scores to read, and P = dS = 0 -- written once, the rows stay zero until the warp is live again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
        mbar_wait(ps_free, psfree_ph ^ 1u); psfree_ph ^= 1u;
        if (ps_dirty) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            sts128(sw128(ps, r, hh * 4 + c4), 0u, 0u, 0u, 0u);
            sts128(sw128(ps + P_BYTES, r, hh * 4 + c4), 0u, 0u, 0u, 0u);
          }
          ps_dirty = false;
          fence_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ps_full);
        readout();
        pend = Pending{dq_done, dkv_done, h, b, qt, kt};
        advance(x);
        continue;
      }
      ps_dirty = true;
      const int kvalid = min(TK, p.Lk - kt * TK);
      uint32_t pw[16], dw[16];
      {
        uint32_t sr[32], dr[32];
        tmem_ld32(tmem_S + lane_base + hh * 32u, sr);
        tmem_ld32(tmem_dP + lane_base + hh * 32u, dr);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int cc = 2 * t + e;
            const bool ok = row_valid && (hh * 32 + cc < kvalid);
            const float pp = ok ? fast_exp2(__uint_as_float(sr[cc]) * p.scale_log2 - lse_l2) : 0.f;
            pv[e] = pp;
            dv[e] = pp * (__uint_as_float(dr[cc]) - delta) * p.scale;
          }
          pw[t] = pack_bf16(pv[0], pv[1]);
          dw[t] = pack_bf16(dv[0], dv[1]);
        }
      }
      mbar_wait(ps_free, psfree_ph ^ 1u); psfree_ph ^= 1u;       // the MMAs that read the previous P / dS have retired
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        sts128(sw128(ps, r, hh * 4 + c4), pw[c4 * 4], pw[c4 * 4 + 1], pw[c4 * 4 + 2], pw[c4 * 4 + 3]);
        sts128(sw128(ps + P_BYTES, r, hh * 4 + c4), dw[c4 * 4], dw[c4 * 4 + 1], dw[c4 * 4 + 2], dw[c4 * 4 + 3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ps_full);
      readout();                                                 // the PREVIOUS step's results
      pend = Pending{dq_done, dkv_done, h, b, qt, kt};
      advance(x);
    }
    readout();
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

bool attn_tc_supported(int64_t D, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k,
                       const void* v, const void* o) {
  auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  return D == 64 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && al(q) && al(k) && al(v) && al(o);
}

int attn_fwd_tc_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                     const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s) {
  CUtensorMap tmQ, tmK, tmV, tmO;
  // [B, L, H*64] views; a load box is one head (64 columns = one 128-byte swizzle row) of a row tile of ONE sample,
  // the store box one warp's 32 output rows
  if (!tc_make_map3_bf16(&tmQ, q, H * HD, Lq, B, ldq, HD, TQ)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmK, k, H * HD, Lk, B, ldk, HD, TK)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmV, v, H * HD, Lk, B, ldv, HD, TK)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmO, o, H * HD, Lq, B, ldo, HD, 32)) return MMSA_ERR_CUDA;
  AttnParams p{};
  p.H = (int)H; p.Lq = (int)Lq; p.Lk = (int)Lk;
  p.nq = (int)ceil_div(Lq, TQ); p.nkv = (int)ceil_div(Lk, TK);
  p.units = B * H * p.nq;
  p.o = (bf16*)o; p.ldo = ldo; p.lse = lse;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  static PerDeviceOnce attr_set;
  if (attr_set.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) { set_error("mmsa: cudaFuncSetAttribute(attn_fwd_tc, smem=%d) failed: %s", SMEM_TOTAL, cudaGetErrorString(e)); return MMSA_ERR_CUDA; }
    attr_set.mark();
  }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const long long grid = p.units < 2LL * sms ? p.units : 2LL * sms;
  ProfScope prof("attn_fwd_tc", s, 2.0 * 64 * (double)B * H * (2.0 * Lq + 2.0 * Lk));
  attn_fwd_tc_kernel<<<(unsigned)grid, kThreads, SMEM_TOTAL, s>>>(tmQ, tmK, tmV, tmO, p);
  MMSA_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return MMSA_OK;
}

// fused backward: usable when one side is a single tile (Lk <= 64 or Lq <= 128)
bool attn_bwd_tc_supported(int64_t Lq, int64_t Lk) { return Lk <= TK || Lq <= TQ; }

int attn_bwd_tc_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                     const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo,
                     const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     cudaStream_t s) {
  CUtensorMap tmQ, tmK, tmV, tmO, tmdO, tmdQ, tmdK, tmdV;
  if (!tc_make_map3_bf16(&tmQ, q, H * HD, Lq, B, ldq, HD, TQ)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmO, o, H * HD, Lq, B, ldo, HD, TQ)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmdO, dout, H * HD, Lq, B, lddo, HD, TQ)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmK, k, H * HD, Lk, B, ldk, HD, TK)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmV, v, H * HD, Lk, B, ldv, HD, TK)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmdQ, dq, H * HD, Lq, B, lddq, HD, 32)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmdK, dk, H * HD, Lk, B, lddk, HD, 32)) return MMSA_ERR_CUDA;
  if (!tc_make_map3_bf16(&tmdV, dv, H * HD, Lk, B, lddv, HD, 32)) return MMSA_ERR_CUDA;
  AttnBwdParams p{};
  p.H = (int)H; p.Lq = (int)Lq; p.Lk = (int)Lk;
  const int nq = (int)ceil_div(Lq, TQ), nkv = (int)ceil_div(Lk, TK);
  p.q_outer = nkv == 1 ? 1 : 0;
  p.nsteps = p.q_outer ? nq : nkv;
  p.units = B * H;
  p.lse = lse;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  static PerDeviceOnce attr_set;
  if (attr_set.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_TOTAL);
    if (e != cudaSuccess) { set_error("mmsa: cudaFuncSetAttribute(attn_bwd_tc, smem=%d) failed: %s", B_SMEM_TOTAL, cudaGetErrorString(e)); return MMSA_ERR_CUDA; }
    attr_set.mark();
  }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const long long grid = p.units < sms ? p.units : sms;
  // algorithmic bytes: Q, O, dO, dQ over Lq rows and K, V, dK, dV over Lk rows, 128 B per (row, head)
  ProfScope prof("attn_bwd_tc", s, 2.0 * 64 * (double)B * H * (4.0 * Lq + 4.0 * Lk));
  attn_bwd_tc_kernel<<<(unsigned)grid, kThreadsBwd, B_SMEM_TOTAL, s>>>(tmQ, tmK, tmV, tmO, tmdO, tmdQ, tmdK, tmdV, p);
  MMSA_LAUNCH_CHECK("attn_bwd_tc_kernel");
  return MMSA_OK;
}

}  // namespace mmsa
