/* mmsa.h -- C ABI of the B200 (sm_100a) fusion / ME-MHACL / contrastive hot path.
 *
 * The reference (zhouyuchenzyccccc/Multimodal-Sentiment-Aanalysis, /root/reference/MML_ZYC) is pure
 * PyTorch and has no FFI of its own: its "operator API" for this path is the nn.Module call
 * `model(eeg, eye, pps, labels)` made by Trainer.py:60 / Tester.py:53 /
 * dataLoader/MultiTaskTrainer.py:199.  Each entry point below replaces the ATen op group that one
 * reference line dispatches; the citation after each declaration names that line.  The Python host
 * (multimodal-sentiment-aanalysis_b200/mmsa) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the caller allocates every
 *    output and workspace, the callee allocates nothing and keeps no reference after returning;
 *  - `dtype` is the storage type of activations: MMSA_F32 (parity mode) or MMSA_BF16 (performance
 *    mode, fp32 accumulation).  Parameters, statistics, losses and parameter gradients are fp32;
 *  - `stream` is a cudaStream_t; all work is enqueued on it, there are no hidden synchronisations;
 *  - return 0 on success, non-zero on error (message through mmsa_last_error()); sm_100 only:
 *    on any other device every compute entry point fails with MMSA_ERR_DEVICE (no fallback).
 */
#ifndef MMSA_H_
#define MMSA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMSA_F32 0
#define MMSA_BF16 1

#define MMSA_OK 0
#define MMSA_ERR_ARG 1
#define MMSA_ERR_CUDA 2
#define MMSA_ERR_DEVICE 3

#define MMSA_ACT_NONE 0
#define MMSA_ACT_SIGMOID 1
#define MMSA_ACT_GELU 2
#define MMSA_ACT_RELU 3

/* BatchNorm block orders */
#define MMSA_BN_THEN_GELU 0 /* Linear-BN-GELU-Dropout, MultimodalModel.py:179-199 */
#define MMSA_RELU_THEN_BN 1 /* Linear-ReLU-BN-Dropout, ME-MHACL/model.py:82-97 */
#define MMSA_BN_ONLY 2

/* contrastive loss kinds */
#define MMSA_LOSS_INFONCE 0 /* MultimodalModel.py:232-260 */
#define MMSA_LOSS_SUPCON 1  /* train.py:16-40 */
#define MMSA_LOSS_NTXENT 2  /* ME-MHACL/train.py:47-66 */

const char* mmsa_version(void);
const char* mmsa_last_error(void);
/* 0 when the current CUDA device is sm_100; MMSA_ERR_DEVICE otherwise. */
int mmsa_check_device(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t mmsa_launch_count(void);
/* Per-launch device timing for bench.py's roofline leg.  mmsa_prof_enable(1) clears and starts
 * recording: every kernel launch is bracketed by CUDA events on its own stream and booked, with its
 * algorithmic work (FLOPs for contraction kernels, bytes for memory-bound ones), under the kernel's
 * name; mmsa_prof_enable(0) stops.  mmsa_prof_collect synchronises the device and returns the
 * number of entries written: names[i*name_stride], counts[i] launches, ms[i] summed device time,
 * work[i] summed algorithmic work.  Not usable while a CUDA graph is being captured. */
void mmsa_prof_enable(int on);
int mmsa_prof_collect(char* names_host, int name_stride, int64_t* counts_host, double* ms_host,
                      double* work_host, int max_entries);

/* ---- dtype plumbing (host `.float()` boundary, Trainer.py:53-54) ---------------------------- */
int mmsa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* `count` independent casts in one launch (fp32 master weights -> bf16 operand copies, once per step):
 * src_host/dst_host/numel_host are HOST arrays of device pointers / element counts. */
int mmsa_cast_multi(int count, const void* const* src_host, void* const* dst_host, const int64_t* numel_host,
                    int src_dtype, int dst_dtype, void* stream);

/* fp32 -> split-bf16 operand for an fp32-accurate tensor-core product (InfoNCE cosine block and its two backward
 * GEMMs in bf16 mode, MultimodalModel.py:237): hi = bf16(x), lo = bf16(x - hi); three copies are concatenated along
 * the reduction axis, (hi,hi,lo) for the A side (b_side = 0) or (hi,lo,hi) for the B side (b_side = 1), so that a
 * plain bf16 GEMM over 3K yields hi.hi + hi.lo + lo.hi.  src:[R,C] fp32 (row stride ld); dst_col:[R,3C] bf16
 * (K-major operand) and/or dst_row:[3R,C] bf16 (MN-major operand); either may be NULL. */
int mmsa_split3(const float* src, int64_t R, int64_t C, int64_t ld, void* dst_col, int b_side_col, void* dst_row,
                int b_side_row, void* stream);

/* test hook: bf16 products below 0.13 GFLOP (the [B,*] tail) run on the latency-optimised mma.sync cluster kernel;
 * 1 keeps them on the tcgen05 engine (engine cross-check), 0 restores the default. */
void mmsa_debug_gemm_engine(int engine);
/* ---- Linear: y = x W^T + b   (nn.Linear: MultimodalModel.py:86,112-121,172-198) -------------
 * x:[M,K] (row stride ldx), optional second operand x2:[M,K2] concatenated on the feature axis
 * (the gate's cat[q, attn], MultimodalModel.py:147), W:[N,K+K2] fp32 master (row stride ldw) or
 * bf16 copy when dtype==BF16 (w_dtype), bias:[N] fp32 or NULL, residual:[M,N] (dtype, stride ldr)
 * or NULL, act applied last.  y:[M,N] (out_dtype, row stride ldy). */
int mmsa_linear_fwd(int dtype, int64_t M, int64_t N, int64_t K, int64_t K2,
                    const void* x, int64_t ldx, const void* x2, int64_t ldx2,
                    const void* w, int64_t ldw, const float* bias,
                    const void* residual, int64_t ldr, int act,
                    void* y, int64_t ldy, int out_dtype, void* stream);
/* tuning probe: raw bf16 tensor-core GEMM, explicit operand majors (scripts/gemm_majors.py). */
int mmsa_debug_gemm(int a_mn, int b_mn, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                    const void* B, int64_t ldb, float* C, int64_t ldc, int splits, int bn, void* stream);
/* dgrad: dx[M,K] = dy[M,N] W[N,K] (+ residual).  W row stride ldw lets a column block of a wider
 * weight be addressed (gate.0.weight[:, :E] / [:, E:]). */
int mmsa_linear_dgrad(int dtype, int64_t M, int64_t N, int64_t K,
                      const void* dy, int64_t lddy, const void* w, int64_t ldw,
                      const void* residual, int64_t ldr,
                      void* dx, int64_t lddx, int out_dtype, void* stream);
/* wgrad: dw[N,K] (fp32, row stride lddw) = dy[M,N]^T x[M,K]; db[N] = column sums of dy (or NULL).
 * workspace: fp32, mmsa_linear_wgrad_workspace(...) bytes (split-K partials). */
int64_t mmsa_linear_wgrad_workspace(int dtype, int64_t M, int64_t N, int64_t K);
int mmsa_linear_wgrad(int dtype, int64_t M, int64_t N, int64_t K,
                      const void* dy, int64_t lddy, const void* x, int64_t ldx,
                      float* dw, int64_t lddw, float* db, void* workspace, void* stream);

/* dw[N, K+K2] = dy^T [x | x2] (+ db): weight gradient of a Linear over the feature-axis concat of two tensors (the gate's
 * cat[q, attn], MultimodalModel.py:147) without the concat.  workspace: mmsa_linear_wgrad_workspace(dtype, M, N, K + K2). */
int mmsa_linear_wgrad2(int dtype, int64_t M, int64_t N, int64_t K, int64_t K2,
                       const void* dy, int64_t lddy, const void* x, int64_t ldx, const void* x2, int64_t ldx2,
                       float* dw, int64_t lddw, float* db, void* workspace, void* stream);

/* ---- multi-head attention core (torch F.multi_head_attention_forward need_weights branch,
 *      reached from MultimodalModel.py:139-143, ME-MHACL/model.py:71) ------------------------
 * q:[B,Lq,H*D] row stride ldq, k/v:[B,Lk,H*D] row strides ldk/ldv (K and V may alias one packed
 * [B,Lk,2E] buffer), o:[B,Lq,H*D] row stride ldo, lse:[B,H,Lq] fp32 (log-sum-exp of the scaled
 * scores; saved for backward instead of the B*H*Lq*Lk probability matrix).  scale = 1/sqrt(D)
 * is applied to q before the product, as torch does.  D in {32, 64}. */
/* test hook: route bf16 attention through the CUDA-core engine (engine cross-check). */
void mmsa_debug_force_simt_attention(int on);
/* test hook: bf16 forward engine, 0 = tcgen05 + TMA (default), 1 = mma.sync (engine cross-check). */
void mmsa_debug_attention_engine(int engine);
int mmsa_attn_fwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D,
                  const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  void* o, int64_t ldo, float* lse, void* stream);
/* delta:[B,H,Lq] fp32 scratch. */
int mmsa_attn_bwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D,
                  const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                  float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                  void* stream);

/* Same core with dropout on the attention PROBABILITIES -- nn.MultiheadAttention(dropout=p) in training mode, as
 * nn.TransformerEncoderLayer(dropout=0.3) builds it (MultimodalModel.py:89-95): O = (softmax(S) o M / (1-p)) V.
 * keep_mask:[B,H,Lq,Lk] uint8 (explicit mask, parity tests) or NULL: the keep decision of element (b,h,i,j) is drawn
 * from Philox at counter offset + ((b*H+h)*Lq+i)*Lk+j and RE-DRAWN by the backward (no B*H*Lq*Lk mask is stored), so fwd
 * and bwd must be given the same seed / offset / rng_state contents.  rng_state: device {seed, position} or NULL (see
 * mmsa_bn_act_fwd).  Runs on the CUDA-core engine (the encoder-tail shapes: L <= 100, 4 heads). */
int mmsa_attn_dropout_fwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D,
                          const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          void* o, int64_t ldo, float* lse, float dropout_p, const uint8_t* keep_mask,
                          uint64_t seed, uint64_t offset, const uint64_t* rng_state, void* stream);
int mmsa_attn_dropout_bwd(int dtype, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D,
                          const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          const void* o, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                          float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                          float dropout_p, const uint8_t* keep_mask, uint64_t seed, uint64_t offset,
                          const uint64_t* rng_state, void* stream);

/* ---- sigmoid gate + blend + LayerNorm (MultimodalModel.py:147-149) --------------------------
 * gate_pre:[M,E] = Linear(2E,E)(cat[q,attn]) before the sigmoid; g_out = sigmoid(gate_pre);
 * u = g*q + (1-g)*attn; y = LayerNorm(u; gamma, beta, eps).  mean/rstd:[M] fp32 saved. */
int mmsa_gate_ln_fwd(int dtype, int64_t M, int64_t E, const void* gate_pre, const void* q, const void* attn,
                     const float* gamma, const float* beta, float eps,
                     void* g_out, void* y, float* mean, float* rstd, void* stream);
/* dy:[M,E]; when dy_rows_per_sample > 0, dy is [M/dy_rows_per_sample, E] and is broadcast over the
 * tokens of a sample and scaled by 1/dy_rows_per_sample (backward of a token mean-pool).
 * dq_bcast (or NULL): [M/bcast_rows, E] gradient of a token mean-pool of q, added to dq_part as
 * dq_bcast[row / bcast_rows] / bcast_rows (saves materialising the broadcast).  dq_add (or NULL):
 * [M,E] further gradient w.r.t. q (e.g. from the mirrored block, where q served as key/value).
 * outputs: dq_part = du*g (+ bcast), dattn_part = du*(1-g), dgate_pre = du*(q-attn)*g*(1-g);
 * dgamma/dbeta:[E] fp32 (partials: [nblk,2,E] fp32 workspace, nblk from mmsa_gate_ln_bwd_blocks). */
int64_t mmsa_gate_ln_bwd_blocks(int64_t M);
int mmsa_gate_ln_bwd(int dtype, int64_t M, int64_t E, const void* dy, int64_t dy_rows_per_sample,
                     const void* g, const void* q, const void* attn,
                     const float* gamma, const float* mean, const float* rstd,
                     const void* dq_bcast, int64_t bcast_rows, const void* dq_add,
                     void* dq_part, void* dattn_part, int64_t ld_parts, void* dgate_pre,
                     float* dgamma, float* dbeta, float* partials, void* stream);
/* ld_parts: row stride (elements; 0 = E) of dq_part and dattn_part -- they may be the two column halves of ONE [M, 2E]
 * buffer, so that the gate's two input gradients dgate W[:, :E] + dq_part and dgate W[:, E:] + dattn_part come out of a
 * single N = 2E GEMM with that buffer as its residual. */

/* Fused form for a block whose output only feeds a token mean-pool (the text+image path pools t' and
 * v' straight away): y is never written; pooled_y[b,:] = mean over the L rows of sample b of LN(u)
 * and pooled_q[b,:] = mean of q are accumulated in fp32 (pooled_q_lp: optional copy in `dtype`, the
 * operand of the modality-weight GEMM).  g_out, mean, rstd are saved for the backward. */
int mmsa_gate_ln_pool_fwd(int dtype, int64_t B, int64_t L, int64_t E, const void* gate_pre, const void* q,
                          const void* attn, const float* gamma, const float* beta, float eps, void* g_out,
                          float* mean, float* rstd, float* pooled_y, float* pooled_q, void* pooled_q_lp,
                          void* stream);
/* dpooled_y/dpooled_q: [B,E] fp32 gradients of the two pooled outputs (dpooled_q may be NULL); dq_add
 * (or NULL): [B*L,E] extra gradient w.r.t. q.  Outputs as mmsa_gate_ln_bwd. */
int mmsa_gate_ln_pool_bwd(int dtype, int64_t B, int64_t L, int64_t E, const float* dpooled_y,
                          const float* dpooled_q, const void* dq_add, const void* g, const void* q,
                          const void* attn, const float* gamma, const float* mean, const float* rstd,
                          void* dq_part, void* dattn_part, int64_t ld_parts, void* dgate_pre, float* dgamma, float* dbeta,
                          float* partials, void* stream);

/* ---- residual add + LayerNorm, positional table (the encoder tail in front of the path: Subnetwork,
 *      MultimodalModel.py:83-105 = proj -> +PositionalEncoding (:19-20) -> 2 x post-norm nn.TransformerEncoderLayer -> LayerNorm;
 *      SURVEY.md section 8(f) rank 2) -----------------------------------------------------------
 * y = LayerNorm(x + r; gamma, beta, eps), r may be NULL (plain LayerNorm); mean/rstd:[M] fp32 saved.
 * bwd: du = dL/d(x + r) (gradient of both summands); dgamma/dbeta:[E]; partials: [mmsa_gate_ln_bwd_blocks(M), 2, E] fp32. */
int mmsa_add_ln_fwd(int dtype, int64_t M, int64_t E, const void* x, const void* r, const float* gamma, const float* beta,
                    float eps, void* y, float* mean, float* rstd, void* stream);
int mmsa_add_ln_bwd(int dtype, int64_t M, int64_t E, const void* dy, const void* x, const void* r, const float* gamma,
                    const float* mean, const float* rstd, void* du, float* dgamma, float* dbeta, float* partials,
                    void* stream);
/* y[m,:] = x[m,:] + pe[m % L, :]  (pe:[>=L, E] fp32, the sinusoidal buffer); backward = identity. */
int mmsa_add_rows(int dtype, int64_t M, int64_t E, int64_t L, const void* x, const float* pe, void* y, void* stream);

/* ---- token pooling (mean: MultimodalModel.py:76 / ME-MHACL/model.py:73; max: MultimodalModel.py:401)
 * x:[B,L,E] -> y:[B,E]; argmax:[B,E] int32 only for max. */
int mmsa_pool_fwd(int dtype, int64_t B, int64_t L, int64_t E, const void* x, int is_max,
                  void* y, int32_t* argmax, void* stream);
int mmsa_pool_bwd(int dtype, int64_t B, int64_t L, int64_t E, const void* dy, int is_max,
                  const int32_t* argmax, void* dx, void* stream);

/* ---- modality weights softmax + weighted concat (MultimodalModel.py:171-176 tail, :299-306) --
 * Mixed-precision rule of the [B,*] tail: every GEMM OUTPUT is fp32, every GEMM OPERAND is `dtype`;
 * the elementwise kernels between two GEMMs read fp32 and write `dtype` (identity in fp32 mode), so
 * reductions over the batch (BatchNorm statistics, bias gradients) never see bf16-rounded values.
 * logits:[B,S] fp32, slots: S x [B,E] fp32 -> w = softmax(logits) (fp32 [B,S]);
 * fused[B,S*E] (dtype) = cat_s(slot_s * w[:,s]).  bwd: dfused fp32 -> dslots fp32, dlogits (dtype). */
int mmsa_modal_concat_fwd(int dtype, int64_t B, int64_t E, int S, const void* logits,
                          const void* const* slots_host, float* w, void* fused, void* stream);
int mmsa_modal_concat_bwd(int dtype, int64_t B, int64_t E, int S, const void* dfused, const float* w,
                          const void* const* slots_host, void* const* dslots_host, void* dlogits, void* stream);

/* ---- modality head: the rest of `attention_weights` behind its first Linear + the weighted concat, one kernel
 *      (MultimodalModel.py:173-175 GELU, Linear(64,3), Softmax(dim=1); :299-306 weights x features, cat) --------------
 * h_pre:[B,Hd] fp32 (output of attention_weights.0), w2:[S,Hd] in `dtype` (operand copy of attention_weights.2.weight),
 * b2:[S] fp32; slots as in mmsa_modal_concat_fwd.  Outputs: hg:[B,Hd] `dtype` = GELU(h_pre) (operand of the W2 weight
 * gradient), w:[B,S] fp32 softmax weights, fused:[B,S*E] fp32, fused_lp:[B,S*E] `dtype` or NULL (bf16 operand copy for the
 * next Linear; ignored when dtype is fp32).  Backward: dfused:[B,S*E] fp32 -> dslots (fp32, NULL entries skipped),
 * dlogits:[B,S] and dh_pre:[B,Hd] in `dtype` (GEMM operands of the two Linear backward products). */
int mmsa_modal_head_fwd(int dtype, int64_t B, int64_t E, int S, int64_t Hd, const float* h_pre, const void* w2, const float* b2,
                        const void* const* slots_host, void* hg, float* w, float* fused, void* fused_lp, void* stream);
int mmsa_modal_head_bwd(int dtype, int64_t B, int64_t E, int S, int64_t Hd, const float* dfused, const float* w,
                        const void* const* slots_host, void* const* dslots_host, const float* h_pre, const void* w2,
                        void* dlogits, void* dh_pre, void* stream);

/* ---- activation (nn.GELU exact-erf, MultimodalModel.py:173; ReLU, ME-MHACL/model.py:108) ----
 * x (and dy) fp32, y / dx in `dtype`. */
int mmsa_act_fwd(int dtype, int64_t n, const void* x, int act, void* y, void* stream);
int mmsa_act_bwd(int dtype, int64_t n, const void* x, const void* dy, int act, void* dx, void* stream);

/* ---- BatchNorm1d + activation + dropout on [B,N] (MultimodalModel.py:180-183, 193-196;
 *      ME-MHACL/model.py:84-87) ---------------------------------------------------------------
 * training!=0: batch statistics (biased var), running stats updated in place with momentum and
 * unbiased var; else running stats.  keep_mask:[B,N] uint8 or NULL: if dropout_p>0 and
 * mask_given==0 the kernel draws it (Philox, seed/offset) and writes it; if mask_given it reads it.
 * rng_state (device, {seed, position}) or NULL: when given, the kernel reads the seed from it and adds its
 * position to `offset` -- the stream position then lives in device memory and moves with mmsa_rng_advance, so a
 * captured CUDA graph draws a new mask on every replay (by-value seed/offset are frozen at capture).
 * num_batches_tracked (device int64 scalar or NULL): incremented by one in training mode (nn.BatchNorm1d's counter, kept
 * by the kernel so that the step has no separate add launch).
 * save_mean/save_rstd:[N] fp32.  x (and dy in bwd) are fp32 GEMM outputs; y / dx are written in `dtype`; y_lp (or NULL):
 * an additional bf16 copy of y when `dtype` is fp32 (the autograd-boundary output that is also the next Linear's operand). */
int mmsa_bn_act_fwd(int dtype, int64_t B, int64_t N, int order, const void* x,
                    const float* gamma, const float* beta, float* running_mean, float* running_var,
                    int64_t* num_batches_tracked, float momentum, float eps, int training,
                    float dropout_p, uint8_t* keep_mask, int mask_given, uint64_t seed, uint64_t offset,
                    const uint64_t* rng_state, void* y, void* y_lp, float* save_mean, float* save_rstd, void* stream);
/* Linear + BatchNorm1d block in ONE launch (bf16 operands, fp32 accumulation): z = x W^T + bias is written in fp32 (what
 * mmsa_bn_act_bwd and the Linear's backward need), then y = dropout(act(BN(z))) exactly as mmsa_bn_act_fwd would compute
 * it from z (same statistics, running-stat update, num_batches_tracked, Philox indexing, keep_mask, y / y_lp outputs).
 * x:[M,K] bf16 (row stride ldx), w:[N,K] bf16 (row stride ldw), bias:[N] fp32 or NULL, z:[M,N] fp32, y:[M,N] out_dtype.
 * One CTA owns all M rows of 16 columns, so M <= 256; also N % 8 == 0, K % 8 == 0, 16-byte aligned operands:
 * mmsa_linear_bn_act_supported(...) != 0 says whether a shape qualifies (host-only check); other shapes take
 * mmsa_linear_fwd + mmsa_bn_act_fwd.  (MultimodalModel.py:179-199, ME-MHACL/model.py:82-97.) */
int mmsa_linear_bn_act_supported(int64_t M, int64_t N, int64_t K, int64_t ldx, int64_t ldw);
int mmsa_linear_bn_act_fwd(int64_t M, int64_t N, int64_t K, const void* x, int64_t ldx, const void* w, int64_t ldw,
                           const float* bias, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float momentum, float eps, int training,
                           int order, float dropout_p, uint8_t* keep_mask, int mask_given, uint64_t seed, uint64_t offset,
                           const uint64_t* rng_state, float* z, int out_dtype, void* y, void* y_lp, float* save_mean,
                           float* save_rstd, void* stream);
int mmsa_bn_act_bwd(int dtype, int64_t B, int64_t N, int order, const void* x, const void* dy,
                    const float* gamma, const float* beta, const float* save_mean, const float* save_rstd, int training,
                    float dropout_p, const uint8_t* keep_mask,
                    void* dx, float* dgamma, float* dbeta, float* dbias_prev, void* stream);
/* dbias_prev (or NULL): [N] fp32 column sums of dx BEFORE rounding to `dtype` -- the bias gradient of the
 * Linear in front of the BatchNorm (exactly zero in exact arithmetic when training). */

/* ---- stand-alone dropout (nn.Dropout after ReLU, ME-MHACL/model.py:105-109); y = keep ? x/(1-p) : 0.
 * x fp32, y in `dtype`.
 * keep_mask:[n] uint8 is drawn (Philox) and written unless mask_given; backward = same call on dy
 * with mask_given=1. */
int mmsa_dropout(int dtype, int64_t n, const void* x, float p, uint8_t* keep_mask, int mask_given,
                 uint64_t seed, uint64_t offset, const uint64_t* rng_state, void* y, void* stream);
/* rng_state[1] += n (one-thread kernel on `stream`): moves the device-resident Philox position past the n
 * draws a step consumed; call once per step after the last dropout kernel of that step. */
int mmsa_rng_advance(uint64_t* rng_state, uint64_t n, void* stream);

/* ---- softmax cross-entropy, mean reduction (nn.CrossEntropyLoss, Trainer.py:17,68) ----------
 * logits:[B,C] fp32, labels:[B] int64; loss:[1] fp32, pred:[B] int64 (argmax, Trainer.py:87).
 * addend:[n_add] fp32 device values or NULL (n_add <= 64): loss = CE + sum(addend) -- the trainer's
 * `CE + contrastive_weight * contrastive_loss` (Trainer.py:68-71) folded into the same launch. */
int mmsa_ce_fwd(int64_t B, int64_t C, const float* logits, const int64_t* labels, const float* addend, int64_t n_add,
                float* loss, int64_t* pred, float* row_loss, void* stream);
/* dlogits:[B,C] in `dtype` (the operand type of the head's dgrad / wgrad GEMMs). */
int mmsa_ce_bwd(int dtype, int64_t B, int64_t C, const float* logits, const int64_t* labels,
                const float* dloss, void* dlogits, void* stream);

/* ---- L2 row normalisation (F.normalize, MultimodalModel.py:234-235) -------------------------- */
int mmsa_l2norm_fwd(int dtype, int64_t B, int64_t E, const void* x, void* y, float* norm, void* stream);
/* dx = (dy - y <y,dy>) / max(norm, 1e-12), dy = dy1 (+ dy2 when not NULL), both fp32 [B,E]. */
int mmsa_l2norm_bwd(int dtype, int64_t B, int64_t E, const void* y, const float* norm, const float* dy1,
                    const float* dy2, void* dx, void* stream);

/* ---- contrastive losses on a similarity block -------------------------------------------------
 * sim:[B,Bg] fp32 = f1n f2n^T (un-scaled cosine block, from mmsa_linear_fwd with out_dtype F32),
 * rows are local samples with global index row_offset+i, columns the gathered global batch.
 * kind INFONCE: MultimodalModel.py:237-260 (inv_temp = 1/temperature, device scalar temperature,
 *   grad flows through the row max; dtemp accumulates dL/dtemperature);
 * kind SUPCON: train.py:24-40; kind NTXENT: ME-MHACL/train.py:55-65 (partner = (i+Bg/2) mod Bg).
 * fwd writes row_stats:[B,4] fp32 (max, all, pos, argmax-as-float bits) and loss:[1] = sum_i l_i/denominator.
 * bwd writes G:[B,Bg] (g_dtype) = dloss * dL/dsim (w.r.t. the UN-scaled cosine) and dtemp:[1].
 * weight (device scalar or NULL): the learnable loss weight of MultimodalModel.py:315-317 folded into the same launches --
 *   fwd: loss = weight * mean, loss_raw:[1] (or NULL) = the unweighted mean; bwd: the upstream gradient becomes
 *   dloss * weight, and dweight:[1] (or NULL) = dloss * loss_raw. */
int mmsa_contrastive_fwd(int kind, int64_t B, int64_t Bg, int64_t row_offset, const float* sim,
                         const int64_t* labels_rows, const int64_t* labels_cols,
                         const float* temperature, float temperature_const, int64_t denom, const float* weight,
                         float* row_stats, float* row_loss, float* loss, float* loss_raw, void* stream);
int mmsa_contrastive_bwd(int kind, int64_t B, int64_t Bg, int64_t row_offset, const float* sim,
                         const int64_t* labels_rows, const int64_t* labels_cols,
                         const float* temperature, float temperature_const, int64_t denom,
                         const float* row_stats, const float* dloss, const float* weight, const float* loss_raw,
                         void* G, int g_dtype, float* dtemp_rows, float* dtemp, float* dweight, void* stream);

/* ---- fused global-norm clip + AdamW over a flat fp32 parameter arena (Trainer.py:19-21,80-81;
 *      SURVEY.md section 8(f) rank 1) ----------------------------------------------------------- */
int mmsa_sumsq(const float* x, int64_t n, float* partials, int64_t nblk, float* out, void* stream);
int mmsa_clip_adamw(float* p, const float* g, float* m, float* v, int64_t n, const float* gradsq,
                    float max_norm, double lr, double beta1, double beta2, double eps, double weight_decay,
                    int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMSA_H_ */
