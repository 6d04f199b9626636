// attention_mma.cu -- bf16 multi-head attention core (D = 64) on tensor cores, flash style.
// The core is ~2.4% of the path's FLOPs and, stand-alone, HBM-bound (arithmetic intensity 28-45
// FLOP/B with 49 image regions): it is judged on the HBM roofline.  What matters is
// that the B*H*Lq*Lk probability matrix the reference materialises (77 MB at B=256, L=128) never
// leaves the SM: S/P live in registers, only Q, K, V, O and one log-sum-exp per row touch HBM.
// Tiles are 64 query rows x 64 keys, 4 warps, warp-level mma.sync m16n8k16 (bf16 in, fp32
// accumulate) fed by ldmatrix from padded shared memory.
// Backward is two kernels (no atomics, deterministic): dQ walks key chunks per query tile,
// dK/dV walks query chunks per key tile using the transposed products S^T = K Q^T, dP^T = V dO^T.
#include "common.cuh"

namespace mmsa {

constexpr int AT = 64;         // tile rows (queries or keys)
constexpr int AD = 64;         // head dim
constexpr int ALD = AD + 8;    // padded smem row (144 B): conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t* r, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 64 x 64 bf16 tile, global (row stride ld) -> padded shared; rows >= valid are zero-filled
__device__ __forceinline__ void load_tile(bf16 (*dst)[ALD], const bf16* src, int64_t ld, int valid, float scale_unused = 1.f) {
  (void)scale_unused;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int idx = threadIdx.x + i * 128;
    int r = idx >> 3, c = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < valid) v = *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + c * 8);
    *reinterpret_cast<uint4*>(&dst[r][c * 8]) = v;
  }
}

// A-operand fragments (16 rows x 64 cols, 4 k-steps) for this warp's 16 rows of a tile
__device__ __forceinline__ void load_a_frags(uint32_t (*f)[4], bf16 (*tile)[ALD], int row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(f[ks], &tile[row0 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]);
}

// acc[8][4] (16 x 64) += A(16 x 64 over k) * Bt^T where tile `bt` is [n rows][k cols] row-major
__device__ __forceinline__ void mma_a_bt(float (*acc)[4], const uint32_t (*af)[4], bf16 (*bt)[ALD], int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, &bt[np * 16 + (lane & 7) + (lane >> 4) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
      mma16816(acc[2 * np], af[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], af[ks], b[2], b[3]);
    }
}

// acc[8][4] (16 x 64) += P(16 x 64 over k, given as accumulator-layout floats) * Bk where tile `bk`
// is [k rows][n cols] row-major (transposed ldmatrix)
__device__ __forceinline__ void mma_p_b(float (*acc)[4], const float (*p)[4], bf16 (*bk)[ALD], int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = pack2(p[2 * ks][0], p[2 * ks][1]);
    a[1] = pack2(p[2 * ks][2], p[2 * ks][3]);
    a[2] = pack2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
    a[3] = pack2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, &bk[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (*a)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) a[i][e] = 0.f;
}

constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128)
attn_fwd_mma_kernel(int H, int Lq, int Lk, const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k,
                    int64_t ldk, const bf16* __restrict__ v, int64_t ldv, bf16* __restrict__ o, int64_t ldo,
                    float* __restrict__ lse, float scale) {
  __shared__ __align__(16) bf16 Qs[AT][ALD];
  __shared__ __align__(16) bf16 Ks[AT][ALD];
  __shared__ __align__(16) bf16 Vs[AT][ALD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * AT;
  const float sl2 = scale * kLog2e;
  load_tile(Qs, q + ((int64_t)b * Lq + q0) * ldq + h * AD, ldq, Lq - q0);
  __syncthreads();
  uint32_t qf[4][4];
  load_a_frags(qf, Qs, warp * 16, lane);
  float oacc[8][4];
  zero_acc(oacc);
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  for (int j0 = 0; j0 < Lk; j0 += AT) {
    __syncthreads();
    load_tile(Ks, k + ((int64_t)b * Lk + j0) * ldk + h * AD, ldk, Lk - j0);
    load_tile(Vs, v + ((int64_t)b * Lk + j0) * ldv + h * AD, ldv, Lk - j0);
    __syncthreads();
    float s[8][4];
    zero_acc(s);
    mma_a_bt(s, qf, Ks, lane);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int col = j0 + nt * 8 + 2 * t + (e & 1);
        if (col >= Lk) s[nt][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float mn[2], alpha[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mn[r] = fmaxf(m[r], mx[r]);
      alpha[r] = exp2f((m[r] - mn[r]) * sl2);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float p = exp2f((s[nt][e] - mn[e >> 1]) * sl2);
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) { l[r] = l[r] * alpha[r] + rs[r]; m[r] = mn[r]; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[nt][e] *= alpha[e >> 1];
    mma_p_b(oacc, s, Vs, lane);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int row = q0 + warp * 16 + g + r * 8;
    if (row < Lq) {
      float inv = 1.f / l[r];
      bf16* op = o + ((int64_t)b * Lq + row) * ldo + h * AD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<__nv_bfloat162*>(op + nt * 8 + 2 * t) =
            __floats2bfloat162_rn(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
      if (t == 0) lse[((int64_t)b * H + h) * Lq + row] = m[r] * scale + logf(l[r]);
    }
  }
}

// ------------------------------------------------------------------ backward: dQ
__global__ void __launch_bounds__(128)
attn_bwd_dq_mma_kernel(int H, int Lq, int Lk, const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k,
                       int64_t ldk, const bf16* __restrict__ v, int64_t ldv, const bf16* __restrict__ dout,
                       int64_t lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                       bf16* __restrict__ dq, int64_t lddq, float scale) {
  __shared__ __align__(16) bf16 Qs[AT][ALD];
  __shared__ __align__(16) bf16 Os[AT][ALD];
  __shared__ __align__(16) bf16 Ks[AT][ALD];
  __shared__ __align__(16) bf16 Vs[AT][ALD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * AT;
  const float sl2 = scale * kLog2e;
  load_tile(Qs, q + ((int64_t)b * Lq + q0) * ldq + h * AD, ldq, Lq - q0);
  load_tile(Os, dout + ((int64_t)b * Lq + q0) * lddo + h * AD, lddo, Lq - q0);
  __syncthreads();
  uint32_t qf[4][4], of[4][4];
  load_a_frags(qf, Qs, warp * 16, lane);
  load_a_frags(of, Os, warp * 16, lane);
  float lse_r[2], del_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int row = q0 + warp * 16 + g + r * 8;
    bool ok = row < Lq;
    lse_r[r] = ok ? lse[((int64_t)b * H + h) * Lq + row] * kLog2e : 0.f;
    del_r[r] = ok ? delta[((int64_t)b * H + h) * Lq + row] : 0.f;
  }
  float dqa[8][4];
  zero_acc(dqa);
  for (int j0 = 0; j0 < Lk; j0 += AT) {
    __syncthreads();
    load_tile(Ks, k + ((int64_t)b * Lk + j0) * ldk + h * AD, ldk, Lk - j0);
    load_tile(Vs, v + ((int64_t)b * Lk + j0) * ldv + h * AD, ldv, Lk - j0);
    __syncthreads();
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_bt(s, qf, Ks, lane);
    mma_a_bt(dp, of, Vs, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int col = j0 + nt * 8 + 2 * t + (e & 1);
        float p = (col < Lk) ? exp2f(s[nt][e] * sl2 - lse_r[e >> 1]) : 0.f;
        s[nt][e] = p * (dp[nt][e] - del_r[e >> 1]);
      }
    mma_p_b(dqa, s, Ks, lane);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int row = q0 + warp * 16 + g + r * 8;
    if (row < Lq) {
      bf16* op = dq + ((int64_t)b * Lq + row) * lddq + h * AD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<__nv_bfloat162*>(op + nt * 8 + 2 * t) =
            __floats2bfloat162_rn(dqa[nt][2 * r] * scale, dqa[nt][2 * r + 1] * scale);
    }
  }
}

// ------------------------------------------------------------------ backward: dK, dV
__global__ void __launch_bounds__(128)
attn_bwd_dkv_mma_kernel(int H, int Lq, int Lk, const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k,
                        int64_t ldk, const bf16* __restrict__ v, int64_t ldv, const bf16* __restrict__ dout,
                        int64_t lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                        bf16* __restrict__ dk, int64_t lddk, bf16* __restrict__ dv, int64_t lddv, float scale) {
  __shared__ __align__(16) bf16 Ks[AT][ALD];
  __shared__ __align__(16) bf16 Vs[AT][ALD];
  __shared__ __align__(16) bf16 Qs[AT][ALD];
  __shared__ __align__(16) bf16 Os[AT][ALD];
  __shared__ float Ls[AT], Ds[AT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int j0 = blockIdx.x * AT;
  const float sl2 = scale * kLog2e;
  load_tile(Ks, k + ((int64_t)b * Lk + j0) * ldk + h * AD, ldk, Lk - j0);
  load_tile(Vs, v + ((int64_t)b * Lk + j0) * ldv + h * AD, ldv, Lk - j0);
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(kf, Ks, warp * 16, lane);
  load_a_frags(vf, Vs, warp * 16, lane);
  float dka[8][4], dva[8][4];
  zero_acc(dka);
  zero_acc(dva);
  for (int i0 = 0; i0 < Lq; i0 += AT) {
    __syncthreads();
    load_tile(Qs, q + ((int64_t)b * Lq + i0) * ldq + h * AD, ldq, Lq - i0);
    load_tile(Os, dout + ((int64_t)b * Lq + i0) * lddo + h * AD, lddo, Lq - i0);
    if (threadIdx.x < AT) {
      int i = i0 + threadIdx.x;
      Ls[threadIdx.x] = i < Lq ? lse[((int64_t)b * H + h) * Lq + i] * kLog2e : 0.f;
      Ds[threadIdx.x] = i < Lq ? delta[((int64_t)b * H + h) * Lq + i] : 0.f;
    }
    __syncthreads();
    float st[8][4], dpt[8][4];      // S^T and dP^T: rows = keys (this warp's 16), cols = queries
    zero_acc(st);
    zero_acc(dpt);
    mma_a_bt(st, kf, Qs, lane);
    mma_a_bt(dpt, vf, Os, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int ci = nt * 8 + 2 * t + (e & 1);
        float p = (i0 + ci < Lq) ? exp2f(st[nt][e] * sl2 - Ls[ci]) : 0.f;
        st[nt][e] = p;
        dpt[nt][e] = p * (dpt[nt][e] - Ds[ci]);
      }
    mma_p_b(dva, st, Os, lane);     // dV += P^T dO
    mma_p_b(dka, dpt, Qs, lane);    // dK += dS^T Q
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int row = j0 + warp * 16 + g + r * 8;
    if (row < Lk) {
      bf16* kp = dk + ((int64_t)b * Lk + row) * lddk + h * AD;
      bf16* vp = dv + ((int64_t)b * Lk + row) * lddv + h * AD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<__nv_bfloat162*>(kp + nt * 8 + 2 * t) =
            __floats2bfloat162_rn(dka[nt][2 * r] * scale, dka[nt][2 * r + 1] * scale);
        *reinterpret_cast<__nv_bfloat162*>(vp + nt * 8 + 2 * t) =
            __floats2bfloat162_rn(dva[nt][2 * r], dva[nt][2 * r + 1]);
      }
    }
  }
}

template <typename T, int D>
int attn_delta(int64_t B, int64_t H, int64_t Lq, const void* o, int64_t ldo, const void* dout, int64_t lddo, float* delta,
               cudaStream_t s);

bool attn_mma_supported(int64_t D, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k,
                        const void* v, const void* o) {
  auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  return D == 64 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && al(q) && al(k) && al(v) && al(o);
}

int attn_fwd_mma_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                      const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(Lq, AT), (unsigned)(B * H));
  ProfScope prof("attn_fwd_mma", s, 2.0 * 64 * (double)B * H * (2.0 * Lq + 2.0 * Lk));
  attn_fwd_mma_kernel<<<grid, 128, 0, s>>>((int)H, (int)Lq, (int)Lk, (const bf16*)q, ldq, (const bf16*)k, ldk,
                                          (const bf16*)v, ldv, (bf16*)o, ldo, lse, 0.125f);
  MMSA_LAUNCH_CHECK("attn_fwd_mma_kernel");
  return MMSA_OK;
}

int attn_bwd_mma_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                      const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo,
                      const float* lse, float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                      int64_t lddv, cudaStream_t s) {
  int rc = attn_delta<bf16, 64>(B, H, Lq, o, ldo, dout, lddo, delta, s);
  if (rc) return rc;
  dim3 g1((unsigned)ceil_div(Lq, AT), (unsigned)(B * H));
  {
  ProfScope prof("attn_bwd_dq_mma", s, 2.0 * 64 * (double)B * H * (3.0 * Lq + 2.0 * Lk));
  attn_bwd_dq_mma_kernel<<<g1, 128, 0, s>>>((int)H, (int)Lq, (int)Lk, (const bf16*)q, ldq, (const bf16*)k, ldk,
                                           (const bf16*)v, ldv, (const bf16*)dout, lddo, lse, delta, (bf16*)dq, lddq,
                                           0.125f);
  }
  MMSA_LAUNCH_CHECK("attn_bwd_dq_mma_kernel");
  dim3 g2((unsigned)ceil_div(Lk, AT), (unsigned)(B * H));
  ProfScope prof("attn_bwd_dkv_mma", s, 2.0 * 64 * (double)B * H * (2.0 * Lq + 4.0 * Lk));
  attn_bwd_dkv_mma_kernel<<<g2, 128, 0, s>>>((int)H, (int)Lq, (int)Lk, (const bf16*)q, ldq, (const bf16*)k, ldk,
                                            (const bf16*)v, ldv, (const bf16*)dout, lddo, lse, delta, (bf16*)dk, lddk,
                                            (bf16*)dv, lddv, 0.125f);
  MMSA_LAUNCH_CHECK("attn_bwd_dkv_mma_kernel");
  return MMSA_OK;
}

}  // namespace mmsa
