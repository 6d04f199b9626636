// attention.cu -- C ABI of the multi-head attention core; picks the engine.
//   bf16, D == 64  -> tensor-core kernels (attention_mma.cu)
//   otherwise      -> exact-fp32 CUDA-core kernels (attention_simt.cu)
#include "common.cuh"

namespace mmsa {

template <typename T, int D>
int attn_fwd_simt(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                  const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s);
template <typename T, int D>
int attn_bwd_simt(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                  const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                  float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, cudaStream_t s);

int attn_fwd_mma_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                      const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s);
int attn_bwd_mma_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                      const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo,
                      const float* lse, float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                      int64_t lddv, cudaStream_t s);
int attn_fwd_tc_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                     const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, cudaStream_t s);
bool attn_tc_supported(int64_t D, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k,
                       const void* v, const void* o);
int attn_bwd_tc_bf16(int64_t B, int64_t H, int64_t Lq, int64_t Lk, const void* q, int64_t ldq, const void* k, int64_t ldk,
                     const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo,
                     const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     cudaStream_t s);
bool attn_bwd_tc_supported(int64_t Lq, int64_t Lk);
bool attn_mma_supported(int64_t D, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k,
                        const void* v, const void* o);

static int g_force_simt_attn = 0;
static int g_attn_engine = 0;      // 0: tcgen05 forward (default), 1: mma.sync forward

}  // namespace mmsa

using namespace mmsa;

extern "C" {

// test hook: route bf16 attention through the CUDA-core engine (engine cross-check)
void mmsa_debug_force_simt_attention(int on) { g_force_simt_attn = on; }
// test hook: 0 = tcgen05/TMA forward (default), 1 = mma.sync forward (engine cross-check)
void mmsa_debug_attention_engine(int engine) { g_attn_engine = engine; }

int mmsa_attn_fwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const void* q, int64_t ldq,
                  const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, float* lse, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(D == 32 || D == 64, "mmsa_attn_fwd: head dim %lld not in {32,64}", (long long)D);
  MMSA_REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0, "mmsa_attn_fwd: bad shape");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == MMSA_BF16 && !g_force_simt_attn && g_attn_engine != 1 && attn_tc_supported(D, ldq, ldk, ldv, ldo, q, k, v, o))
    return attn_fwd_tc_bf16(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, s);      // tcgen05 + TMA
  if (dtype == MMSA_BF16 && !g_force_simt_attn && attn_mma_supported(D, ldq, ldk, ldv, ldo, q, k, v, o))
    return attn_fwd_mma_bf16(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, s);     // mma.sync (engine cross-check)
  if (dtype == MMSA_F32) {
    if (D == 64) return attn_fwd_simt<float, 64>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, s);
    return attn_fwd_simt<float, 32>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, s);
  }
  if (dtype == MMSA_BF16) {
    if (D == 64) return attn_fwd_simt<bf16, 64>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, s);
    return attn_fwd_simt<bf16, 32>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, lse, s);
  }
  set_error("mmsa_attn_fwd: bad dtype %d", dtype);
  return MMSA_ERR_ARG;
}

int mmsa_attn_bwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const void* q, int64_t ldq,
                  const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout,
                  int64_t lddo, const float* lse, float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                  int64_t lddv, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(D == 32 || D == 64, "mmsa_attn_bwd: head dim %lld not in {32,64}", (long long)D);
  MMSA_REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0, "mmsa_attn_bwd: bad shape");
  if (B == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == MMSA_BF16 && !g_force_simt_attn && g_attn_engine != 1 && attn_bwd_tc_supported(Lq, Lk) &&
      attn_tc_supported(D, ldq, ldk, ldv, ldo, q, k, v, o) && attn_tc_supported(D, lddq, lddk, lddv, lddo, dq, dk, dv, dout))
    return attn_bwd_tc_bf16(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, dq, lddq, dk, lddk, dv, lddv, s);
  if (dtype == MMSA_BF16 && !g_force_simt_attn && attn_mma_supported(D, ldq, ldk, ldv, ldo, q, k, v, o) &&
      attn_mma_supported(D, lddq, lddk, lddv, lddo, dq, dk, dv, dout))
    return attn_bwd_mma_bf16(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk,
                             dv, lddv, s);
  if (dtype == MMSA_F32) {
    if (D == 64) return attn_bwd_simt<float, 64>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, s);
    return attn_bwd_simt<float, 32>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, s);
  }
  if (dtype == MMSA_BF16) {
    if (D == 64) return attn_bwd_simt<bf16, 64>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, s);
    return attn_bwd_simt<bf16, 32>(B, H, Lq, Lk, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, s);
  }
  set_error("mmsa_attn_bwd: bad dtype %d", dtype);
  return MMSA_ERR_ARG;
}

}  // extern "C"
