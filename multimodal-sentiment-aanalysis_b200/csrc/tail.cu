// tail.cu -- small [B, N] kernels behind the pooled features:
//   BatchNorm1d(+GELU/ReLU)+dropout fwd/bwd, softmax cross-entropy, contrastive row reductions
//   (InfoNCE / SupCon / NT-Xent) and their gradients, global-norm clip + AdamW.
// Reference arithmetic: MultimodalModel.py:179-199 (fusion / heads), Trainer.py:17,68 (CE),
// MultimodalModel.py:232-260 (InfoNCE), train.py:16-40 (SupCon), ME-MHACL/train.py:47-66 (NT-Xent),
// Trainer.py:19-21,80-81 (clip + AdamW).
#include "common.cuh"

namespace mmsa {

// ------------------------------------------------------------------ BatchNorm1d + act + dropout
// block (kBnCols columns x kBnRows row lanes = 1024 threads); one block per kBnCols columns; three passes over the
// (tiny, L1-resident) [B,N] input.  The [B,*] tail is latency-bound: the wide block keeps each thread's serial row walk to
// B / 64 rows, and 16-column blocks spread N = 128..256 over 8..16 SMs.
constexpr int kBnCols = 16, kBnRows = 64;
template <typename T>
__global__ void __launch_bounds__(kBnCols * kBnRows)
bn_act_fwd_kernel(int64_t B, int N, int order, const float* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ running_mean,
                  float* __restrict__ running_var, int64_t* __restrict__ num_batches_tracked, float momentum, float eps,
                  int training,
                  float dropout_p, uint8_t* __restrict__ keep_mask, int mask_given, uint64_t seed,
                  uint64_t offset, const uint64_t* __restrict__ rng_state, T* __restrict__ y, bf16* __restrict__ y_lp,
                  float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float red[kBnRows][kBnCols + 1];
  if (rng_state != nullptr) { seed = rng_state[0]; offset += rng_state[1]; }   // device-resident stream position
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * kBnCols + tx;
  const bool ok = col < N;
  const bool pre_relu = (order == MMSA_RELU_THEN_BN);
  float mean = 0.f, var = 1.f;
  if (training) {
    float s = 0.f;
    if (ok)
      for (int64_t r = ty; r < B; r += kBnRows) {
        float v = to_f(x[r * N + col]);
        if (pre_relu) v = fmaxf(v, 0.f);
        s += v;
      }
    red[ty][tx] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRows; ++k) s += red[k][tx];
    mean = s / (float)B;
    __syncthreads();
    float q = 0.f;
    if (ok)
      for (int64_t r = ty; r < B; r += kBnRows) {
        float v = to_f(x[r * N + col]);
        if (pre_relu) v = fmaxf(v, 0.f);
        float d = v - mean;
        q += d * d;
      }
    red[ty][tx] = q;
    __syncthreads();
    q = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRows; ++k) q += red[k][tx];
    var = q / (float)B;
    if (num_batches_tracked != nullptr && blockIdx.x == 0 && tx == 0 && ty == 0) *num_batches_tracked += 1;
    if (ok && ty == 0 && running_mean != nullptr) {
      float unbiased = B > 1 ? q / (float)(B - 1) : var;
      running_mean[col] = (1.f - momentum) * running_mean[col] + momentum * mean;
      running_var[col] = (1.f - momentum) * running_var[col] + momentum * unbiased;
    }
  } else if (ok) {
    mean = running_mean[col];
    var = running_var[col];
  }
  const float rstd = 1.f / sqrtf(var + eps);
  if (ok && ty == 0) { save_mean[col] = mean; save_rstd[col] = rstd; }
  if (!ok) return;
  const float gm = gamma[col], bt = beta[col];
  const bool drop = training && dropout_p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - dropout_p) : 1.f;
  for (int64_t r = ty; r < B; r += kBnRows) {
    float v = to_f(x[r * N + col]);
    if (pre_relu) v = fmaxf(v, 0.f);
    float z = (v - mean) * rstd * gm + bt;
    if (order == MMSA_BN_THEN_GELU) z = gelu_erf(z);
    if (drop) {
      uint8_t keep;
      if (mask_given) keep = keep_mask[r * N + col];
      else {
        uint32_t rnd = philox_first(seed, offset + (uint64_t)(r * N + col));
        keep = ((float)(rnd >> 8) * (1.f / 16777216.f)) >= dropout_p ? 1 : 0;
        keep_mask[r * N + col] = keep;
      }
      z = keep ? z * keep_scale : 0.f;
    }
    y[r * N + col] = from_f<T>(z);
    if (y_lp != nullptr) y_lp[r * N + col] = __float2bfloat16_rn(z);
  }
}

// Register-resident forms for B <= kBnRegRows * kBnRows (= 256, every batch of the tail on the benchmarked path): each
// thread loads its (at most) four rows ONCE, all loads in flight together, and the later passes run from registers.  The
// loops of the general kernels above / below re-read x per pass, four dependent L2 round trips each: at these sizes the
// kernels are nothing but that latency (8-9 us; ~3 us in this form).  Same arithmetic in the same order: bit-identical.
constexpr int kBnRegRows = 4;
template <typename T>
__global__ void __launch_bounds__(kBnCols * kBnRows)
bn_act_fwd_small_kernel(int B, int N, int order, const float* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ running_mean,
                        float* __restrict__ running_var, int64_t* __restrict__ num_batches_tracked, float momentum,
                        float eps, int training, float dropout_p, uint8_t* __restrict__ keep_mask, int mask_given,
                        uint64_t seed, uint64_t offset, const uint64_t* __restrict__ rng_state, T* __restrict__ y,
                        bf16* __restrict__ y_lp, float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float red[kBnRows][kBnCols + 1];
  if (rng_state != nullptr) { seed = rng_state[0]; offset += rng_state[1]; }
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * kBnCols + tx;
  const bool ok = col < N;
  const bool pre_relu = (order == MMSA_RELU_THEN_BN);
  const bool drop = training && dropout_p > 0.f;
  float v[kBnRegRows];
  uint8_t keep[kBnRegRows];
  bool live[kBnRegRows];
#pragma unroll
  for (int i = 0; i < kBnRegRows; ++i) {
    const int r = ty + i * kBnRows;
    live[i] = ok && r < B;
    v[i] = live[i] ? x[(int64_t)r * N + col] : 0.f;
    keep[i] = (live[i] && drop && mask_given) ? keep_mask[(int64_t)r * N + col] : (uint8_t)1;
    if (pre_relu) v[i] = fmaxf(v[i], 0.f);
  }
  float mean = 0.f, var = 1.f;
  if (training) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kBnRegRows; ++i) if (live[i]) s += v[i];
    red[ty][tx] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRows; ++k) s += red[k][tx];
    mean = s / (float)B;
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kBnRegRows; ++i) if (live[i]) { const float d = v[i] - mean; q += d * d; }
    red[ty][tx] = q;
    __syncthreads();
    q = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRows; ++k) q += red[k][tx];
    var = q / (float)B;
    if (num_batches_tracked != nullptr && blockIdx.x == 0 && tx == 0 && ty == 0) *num_batches_tracked += 1;
    if (ok && ty == 0 && running_mean != nullptr) {
      float unbiased = B > 1 ? q / (float)(B - 1) : var;
      running_mean[col] = (1.f - momentum) * running_mean[col] + momentum * mean;
      running_var[col] = (1.f - momentum) * running_var[col] + momentum * unbiased;
    }
  } else if (ok) {
    mean = running_mean[col];
    var = running_var[col];
  }
  const float rstd = 1.f / sqrtf(var + eps);
  if (ok && ty == 0) { save_mean[col] = mean; save_rstd[col] = rstd; }
  if (!ok) return;
  const float gm = gamma[col], bt = beta[col];
  const float keep_scale = drop ? 1.f / (1.f - dropout_p) : 1.f;
#pragma unroll
  for (int i = 0; i < kBnRegRows; ++i) {
    if (!live[i]) continue;
    const int64_t idx = (int64_t)(ty + i * kBnRows) * N + col;
    float z = (v[i] - mean) * rstd * gm + bt;
    if (order == MMSA_BN_THEN_GELU) z = gelu_erf(z);
    if (drop) {
      uint8_t kp = keep[i];
      if (!mask_given) {
        uint32_t rnd = philox_first(seed, offset + (uint64_t)idx);
        kp = ((float)(rnd >> 8) * (1.f / 16777216.f)) >= dropout_p ? 1 : 0;
        keep_mask[idx] = kp;
      }
      z = kp ? z * keep_scale : 0.f;
    }
    y[idx] = from_f<T>(z);
    if (y_lp != nullptr) y_lp[idx] = __float2bfloat16_rn(z);
  }
}

template <typename T>
__global__ void __launch_bounds__(kBnCols * kBnRows)
bn_act_bwd_small_kernel(int B, int N, int order, const float* __restrict__ x, const float* __restrict__ dy,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ save_mean, const float* __restrict__ save_rstd, int training,
                        float dropout_p, const uint8_t* __restrict__ keep_mask, T* __restrict__ dx,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias_prev) {
  __shared__ float red[2][kBnRows][kBnCols + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * kBnCols + tx;
  const bool ok = col < N;
  const bool pre_relu = (order == MMSA_RELU_THEN_BN);
  const bool drop = training && dropout_p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - dropout_p) : 1.f;
  float raw[kBnRegRows], d[kBnRegRows];
  uint8_t keep[kBnRegRows];
  bool live[kBnRegRows];
#pragma unroll
  for (int i = 0; i < kBnRegRows; ++i) {                       // every load of the kernel, in flight together
    const int r = ty + i * kBnRows;
    live[i] = ok && r < B;
    const int64_t idx = (int64_t)r * N + col;
    raw[i] = live[i] ? x[idx] : 0.f;
    d[i] = live[i] ? dy[idx] : 0.f;
    keep[i] = (live[i] && drop) ? keep_mask[idx] : (uint8_t)1;
  }
  const float mean = ok ? save_mean[col] : 0.f, rstd = ok ? save_rstd[col] : 0.f;
  const float gm = ok ? gamma[col] : 0.f, bt = ok ? beta[col] : 0.f;
  float xh[kBnRegRows];
  float sdz = 0.f, sdzx = 0.f;
#pragma unroll
  for (int i = 0; i < kBnRegRows; ++i) {
    if (!live[i]) { xh[i] = 0.f; continue; }
    const float v = pre_relu ? fmaxf(raw[i], 0.f) : raw[i];
    xh[i] = (v - mean) * rstd;
    if (drop) d[i] = keep[i] ? d[i] * keep_scale : 0.f;
    if (order == MMSA_BN_THEN_GELU) d[i] *= gelu_erf_grad(xh[i] * gm + bt);
    sdz += d[i];
    sdzx += d[i] * xh[i];
  }
  red[0][ty][tx] = sdz;
  red[1][ty][tx] = sdzx;
  __syncthreads();
  sdz = 0.f; sdzx = 0.f;
#pragma unroll
  for (int k = 0; k < kBnRows; ++k) { sdz += red[0][k][tx]; sdzx += red[1][k][tx]; }
  if (ok && ty == 0) { dgamma[col] = sdzx; dbeta[col] = sdz; }
  const float invB = 1.f / (float)B;
  float sdx = 0.f;
#pragma unroll
  for (int i = 0; i < kBnRegRows; ++i) {
    if (!live[i]) continue;
    float dv = training ? gm * rstd * (d[i] - sdz * invB - xh[i] * sdzx * invB) : gm * rstd * d[i];
    if (pre_relu && raw[i] <= 0.f) dv = 0.f;
    sdx += dv;
    dx[(int64_t)(ty + i * kBnRows) * N + col] = from_f<T>(dv);
  }
  __syncthreads();
  red[0][ty][tx] = sdx;
  __syncthreads();
  if (ok && ty == 0 && dbias_prev != nullptr) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRows; ++k) t += red[0][k][tx];
    dbias_prev[col] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(kBnCols * kBnRows)
bn_act_bwd_kernel(int64_t B, int N, int order, const float* __restrict__ x, const float* __restrict__ dy,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ save_mean, const float* __restrict__ save_rstd, int training,
                  float dropout_p, const uint8_t* __restrict__ keep_mask, T* __restrict__ dx,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias_prev) {
  __shared__ float red[2][kBnRows][kBnCols + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * kBnCols + tx;
  const bool ok = col < N;
  const bool pre_relu = (order == MMSA_RELU_THEN_BN);
  const bool drop = training && dropout_p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - dropout_p) : 1.f;
  const float mean = ok ? save_mean[col] : 0.f, rstd = ok ? save_rstd[col] : 0.f;
  const float gm = ok ? gamma[col] : 0.f, bt = ok ? beta[col] : 0.f;
  float sdz = 0.f, sdzx = 0.f;
  if (ok)
    for (int64_t r = ty; r < B; r += kBnRows) {
      float v = to_f(x[r * N + col]);
      if (pre_relu) v = fmaxf(v, 0.f);
      float xh = (v - mean) * rstd;
      float d = to_f(dy[r * N + col]);
      if (drop) d = keep_mask[r * N + col] ? d * keep_scale : 0.f;
      if (order == MMSA_BN_THEN_GELU) d *= gelu_erf_grad(xh * gm + bt);
      sdz += d;
      sdzx += d * xh;
    }
  red[0][ty][tx] = sdz;
  red[1][ty][tx] = sdzx;
  __syncthreads();
  sdz = 0.f; sdzx = 0.f;
#pragma unroll
  for (int k = 0; k < kBnRows; ++k) { sdz += red[0][k][tx]; sdzx += red[1][k][tx]; }
  if (ok && ty == 0) { dgamma[col] = sdzx; dbeta[col] = sdz; }
  const float invB = 1.f / (float)B;
  float sdx = 0.f;
  if (ok)
    for (int64_t r = ty; r < B; r += kBnRows) {
      float raw = to_f(x[r * N + col]);
      float v = pre_relu ? fmaxf(raw, 0.f) : raw;
      float xh = (v - mean) * rstd;
      float d = to_f(dy[r * N + col]);
      if (drop) d = keep_mask[r * N + col] ? d * keep_scale : 0.f;
      if (order == MMSA_BN_THEN_GELU) d *= gelu_erf_grad(xh * gm + bt);
      float dv = training ? gm * rstd * (d - sdz * invB - xh * sdzx * invB) : gm * rstd * d;
      if (pre_relu && raw <= 0.f) dv = 0.f;
      sdx += dv;
      dx[r * N + col] = from_f<T>(dv);
    }
  // column sums of dx (no early exits above: every thread of the block reaches these barriers)
  __syncthreads();
  red[0][ty][tx] = sdx;
  __syncthreads();
  if (ok && ty == 0 && dbias_prev != nullptr) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRows; ++k) t += red[0][k][tx];
    dbias_prev[col] = t;
  }
}

// ------------------------------------------------------------------ softmax cross-entropy
__global__ void ce_fwd_kernel(int64_t B, int C, const float* __restrict__ logits, const int64_t* __restrict__ labels,
                              int64_t* __restrict__ pred, float* __restrict__ row_loss) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const float* z = logits + r * C;
  float mx = z[0]; int am = 0;
  for (int c = 1; c < C; ++c) if (z[c] > mx) { mx = z[c]; am = c; }
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(z[c] - mx);
  int64_t y = labels[r];
  row_loss[r] = (logf(s) + mx) - z[y];
  if (pred) pred[r] = am;
}

// One-launch form for the batch sizes of the path (B <= kCeOneMax): row losses, argmax, the ordered mean and the optional
// addend (loss = CE + sum(addend): the trainer's `CE + w * contrastive`, Trainer.py:68-71, without an add kernel).
constexpr int64_t kCeOneMax = 16384;
__global__ void __launch_bounds__(1024)
ce_fwd_one_kernel(int64_t B, int C, const float* __restrict__ logits, const int64_t* __restrict__ labels,
                  int64_t* __restrict__ pred, float* __restrict__ row_loss, const float* __restrict__ addend, int n_add,
                  float* __restrict__ loss) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int64_t r = threadIdx.x; r < B; r += blockDim.x) {
    const float* z = logits + r * C;
    float mx = z[0]; int am = 0;
    for (int c = 1; c < C; ++c) if (z[c] > mx) { mx = z[c]; am = c; }
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(z[c] - mx);
    const float l = (logf(s) + mx) - z[labels[r]];
    row_loss[r] = l;
    if (pred) pred[r] = am;
    acc += l;
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    float extra = 0.f;
    for (int i = 0; i < n_add; ++i) extra += addend[i];
    loss[0] = acc / (float)B + extra;
  }
}

// deterministic single-block sum: out[0] = scale * sum(v[0..n)) (+ sum(addend[0..n_add)))
__global__ void sum_scale_add_kernel(const float* __restrict__ v, int64_t n, float scale, const float* __restrict__ addend,
                                     int n_add, float* __restrict__ out) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, sm);
  if (threadIdx.x == 0) {
    float extra = 0.f;
    for (int i = 0; i < n_add; ++i) extra += addend[i];
    out[0] = s * scale + extra;
  }
}

// tail launches of the contrastive losses: the ordered mean of the row losses, times the (device-resident) loss weight
// when there is one (`contrastive_weight * loss`, MultimodalModel.py:315-317), and in the backward the temperature
// gradient's ordered sum plus d weight = d loss * unweighted loss
__global__ void contrastive_finish_kernel(const float* __restrict__ row_loss, int64_t n, float scale,
                                          const float* __restrict__ weight, float* __restrict__ loss,
                                          float* __restrict__ loss_raw) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += row_loss[i];
  s = block_sum(s, sm);
  if (threadIdx.x == 0) {
    const float raw = s * scale;
    if (loss_raw != nullptr) loss_raw[0] = raw;
    loss[0] = weight != nullptr ? weight[0] * raw : raw;
  }
}
__global__ void contrastive_bwd_finish_kernel(const float* __restrict__ dtemp_rows, int64_t n, float* __restrict__ dtemp,
                                              const float* __restrict__ dloss, const float* __restrict__ loss_raw,
                                              float* __restrict__ dweight) {
  __shared__ float sm[32];
  float s = 0.f;
  if (dtemp_rows != nullptr && dtemp != nullptr) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += dtemp_rows[i];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) dtemp[0] = s;
  }
  if (threadIdx.x == 0 && dweight != nullptr) dweight[0] = dloss[0] * loss_raw[0];
}

// deterministic single-block sum: out[0] = scale * sum(v[0..n))
__global__ void sum_scale_kernel(const float* __restrict__ v, int64_t n, float scale, float* __restrict__ out) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, sm);
  if (threadIdx.x == 0) out[0] = s * scale;
}

template <typename T>
__global__ void ce_bwd_kernel(int64_t B, int C, const float* __restrict__ logits, const int64_t* __restrict__ labels,
                              const float* __restrict__ dloss, T* __restrict__ dlogits) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const float* z = logits + r * C;
  float mx = z[0];
  for (int c = 1; c < C; ++c) mx = fmaxf(mx, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(z[c] - mx);
  float g = dloss[0] / (float)B;
  int64_t y = labels[r];
  for (int c = 0; c < C; ++c) dlogits[r * C + c] = from_f<T>(g * (expf(z[c] - mx) / s - (c == y ? 1.f : 0.f)));
}

// ------------------------------------------------------------------ contrastive row kernels
// one warp per row of the [B, Bg] cosine block.
__device__ __forceinline__ float load_temp(const float* tptr, float tconst) { return tptr ? tptr[0] : tconst; }

__global__ void __launch_bounds__(256)
contrastive_fwd_kernel(int kind, int64_t B, int64_t Bg, int64_t row_offset, const float* __restrict__ sim,
                       const int64_t* __restrict__ lab_r, const int64_t* __restrict__ lab_c,
                       const float* __restrict__ tptr, float tconst, float* __restrict__ row_stats,
                       float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= B) return;
  const float T = load_temp(tptr, tconst);
  const float* row = sim + i * Bg;
  const int64_t gi = row_offset + i;
  if (kind == MMSA_LOSS_INFONCE) {
    const int64_t yi = lab_r[i];
    float mx = -INFINITY; int64_t am = 0;
    for (int64_t j = lane; j < Bg; j += 32) {
      float s = row[j] / T;
      if (s > mx) { mx = s; am = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // max with first-index tie break (torch.max CPU semantics)
      float om = __shfl_xor_sync(0xffffffffu, mx, o);
      long long oa = __shfl_xor_sync(0xffffffffu, (long long)am, o);
      if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    // The arg-max column contributes exp(0) = 1 exactly; summing the REST separately keeps
    // all - 1 and pos - [argmax is a positive] exact, which the backward needs when the other
    // terms underflow against 1 (the T = 0.01 regime the reference starts in).
    float all_rest = 0.f, pos_rest = 0.f;
    for (int64_t j = lane; j < Bg; j += 32) {
      if (j == am) continue;
      float e = expf(row[j] / T - mx);
      all_rest += e;
      if (j != gi && lab_c[j] == yi) pos_rest += e;
    }
    all_rest = warp_sum(all_rest); pos_rest = warp_sum(pos_rest);
    if (lane == 0) {
      const float am_pos = (am != gi && lab_c[am] == yi) ? 1.f : 0.f;
      row_stats[i * 4 + 0] = mx; row_stats[i * 4 + 1] = all_rest; row_stats[i * 4 + 2] = pos_rest;
      row_stats[i * 4 + 3] = __int_as_float((int)am);
      row_loss[i] = log1pf(all_rest + 1e-12f) - logf(pos_rest + am_pos + 1e-12f);
    }
  } else if (kind == MMSA_LOSS_SUPCON) {
    const int64_t yi = lab_r[i];
    float se = 0.f, cnt = 0.f, ms = 0.f;
    for (int64_t j = lane; j < Bg; j += 32) {
      if (j == gi) continue;
      float s = row[j] / T;
      se += expf(s);
      if (lab_c[j] == yi) { cnt += 1.f; ms += s; }
    }
    se = warp_sum(se); cnt = warp_sum(cnt); ms = warp_sum(ms);
    if (lane == 0) {
      float lse = logf(se + 1e-8f);
      row_stats[i * 4 + 0] = se; row_stats[i * 4 + 1] = cnt; row_stats[i * 4 + 2] = ms; row_stats[i * 4 + 3] = lse;
      row_loss[i] = -(ms - cnt * lse) / (cnt + 1e-8f);
    }
  } else {  // NT-Xent: CE over the row with the diagonal masked, target = partner view
    const int64_t partner = (gi + Bg / 2) % Bg;
    float mx = -INFINITY;
    for (int64_t j = lane; j < Bg; j += 32) if (j != gi) mx = fmaxf(mx, row[j] / T);
    mx = warp_max(mx);
    float se = 0.f;
    for (int64_t j = lane; j < Bg; j += 32) if (j != gi) se += expf(row[j] / T - mx);
    se = warp_sum(se);
    if (lane == 0) {
      row_stats[i * 4 + 0] = mx; row_stats[i * 4 + 1] = se; row_stats[i * 4 + 2] = 0.f; row_stats[i * 4 + 3] = 0.f;
      row_loss[i] = (logf(se) + mx) - row[partner] / T;
    }
  }
}

template <typename TG>
__global__ void __launch_bounds__(256)
contrastive_bwd_kernel(int kind, int64_t B, int64_t Bg, int64_t row_offset, const float* __restrict__ sim,
                       const int64_t* __restrict__ lab_r, const int64_t* __restrict__ lab_c,
                       const float* __restrict__ tptr, float tconst, float inv_denom,
                       const float* __restrict__ row_stats, const float* __restrict__ dloss,
                       const float* __restrict__ weight, TG* __restrict__ G, float* __restrict__ dtemp_rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= B) return;
  const float T = load_temp(tptr, tconst);
  const float up = (weight != nullptr ? dloss[0] * weight[0] : dloss[0]) * inv_denom;
  const float* row = sim + i * Bg;
  TG* grow = G + i * Bg;
  const int64_t gi = row_offset + i;
  float dts = 0.f;   // sum_j gs_ij * s_ij
  if (kind == MMSA_LOSS_INFONCE) {
    const int64_t yi = lab_r[i];
    const float mx = row_stats[i * 4 + 0], all_rest = row_stats[i * 4 + 1], pos_rest = row_stats[i * 4 + 2];
    const int64_t am = (int64_t)__float_as_int(row_stats[i * 4 + 3]);
    const float am_pos = (am != gi && lab_c[am] == yi) ? 1.f : 0.f;
    const float ia = 1.f / (1.f + all_rest + 1e-12f), ip = 1.f / (pos_rest + am_pos + 1e-12f);
    // d/ds_ij of -log((pos+eps)/(all+eps)) with s_ij = sim_ij/T - max_i: e_ij/all' - [pos_ij] e_ij/pos'.
    // Autograd also flows through the subtracted row max (MultimodalModel.py:245): the arg-max column
    // (first index on ties) receives pos/pos' - all/all' on top; with e_am = 1 the two combine to
    // -all_rest/all' + pos_rest/pos', free of cancellation.
    const float g_am = pos_rest * ip - all_rest * ia;
    // dL/dT = -(1/T) sum_j g_ij s_ij.  The g_ij of a row sum to zero exactly (see above), so s_ij may
    // be replaced by the shifted s_ij - max_i: the arg-max term drops out and the remaining terms are
    // all of the size of the e_ij -- no cancellation between logits of magnitude 1/T.
    for (int64_t j = lane; j < Bg; j += 32) {
      float s = row[j] / T;
      float gs;
      if (j == am) gs = g_am;
      else {
        float e = expf(s - mx);
        gs = e * ia;
        if (j != gi && lab_c[j] == yi) gs -= e * ip;
        dts += gs * up * (s - mx);
      }
      gs *= up;
      grow[j] = from_f<TG>(gs / T);
    }
  } else if (kind == MMSA_LOSS_SUPCON) {
    const int64_t yi = lab_r[i];
    const float se = row_stats[i * 4 + 0], cnt = row_stats[i * 4 + 1];
    const float ic = 1.f / (cnt + 1e-8f), cfrac = cnt * ic, ise = 1.f / (se + 1e-8f);
    for (int64_t j = lane; j < Bg; j += 32) {
      float gs = 0.f, s = row[j] / T;
      if (j != gi) {
        gs = cfrac * expf(s) * ise;
        if (lab_c[j] == yi) gs -= ic;
      }
      gs *= up;
      dts += gs * s;
      grow[j] = from_f<TG>(gs / T);
    }
  } else {
    const int64_t partner = (gi + Bg / 2) % Bg;
    const float mx = row_stats[i * 4 + 0], se = row_stats[i * 4 + 1];
    for (int64_t j = lane; j < Bg; j += 32) {
      float gs = 0.f, s = row[j] / T;
      if (j != gi) gs = expf(s - mx) / se;
      if (j == partner) gs -= 1.f;
      gs *= up;
      dts += gs * s;
      grow[j] = from_f<TG>(gs / T);
    }
  }
  dts = warp_sum(dts);
  if (lane == 0 && dtemp_rows) dtemp_rows[i] = -dts / T;   // ds/dT = -s/T
}

// ------------------------------------------------------------------ clip + AdamW
__global__ void sumsq_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partials) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    s += v * v;
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// torch.nn.utils.clip_grad_norm_(params, max_norm) then AdamW (decoupled weight decay), Trainer.py:19-21,80-81
__global__ void clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, int64_t n, const float* __restrict__ gradsq,
                                  float max_norm, float decay, float beta1, float omb1, float beta2, float omb2,
                                  float eps, float step_size, float bc2_sqrt) {
  // every derived scalar (1 - lr*wd, 1 - beta, lr / bias_correction1, sqrt(bias_correction2)) is formed in DOUBLE on
  // the host and rounded once, as torch.optim.AdamW does with its Python floats: 1.f - 0.999f is 4.7e-5 off 0.001
  float total = sqrtf(gradsq[0]);
  float coef = max_norm / (total + 1e-6f);
  coef = coef < 1.f ? coef : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * coef;
    float pi = p[i] * decay;
    float mi = m[i] + (gi - m[i]) * omb1;                 // exp_avg.lerp_(grad, 1 - beta1)
    float vi = beta2 * v[i] + omb2 * (gi * gi);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace mmsa

using namespace mmsa;

extern "C" {

int mmsa_bn_act_fwd(int dtype, int64_t B, int64_t N, int order, const void* x, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    float momentum, float eps,
                    int training, float dropout_p, uint8_t* keep_mask, int mask_given, uint64_t seed,
                    uint64_t offset, const uint64_t* rng_state, void* y, void* y_lp, float* save_mean, float* save_rstd,
                    void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(B > 0 && N > 0, "mmsa_bn_act_fwd: empty input");
  MMSA_REQUIRE(training || (running_mean && running_var), "mmsa_bn_act_fwd: eval mode needs running stats");
  MMSA_REQUIRE(!(training && dropout_p > 0.f) || keep_mask != nullptr, "mmsa_bn_act_fwd: dropout needs keep_mask storage");
  MMSA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "mmsa_bn_act_fwd: dropout_p out of [0,1)");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("bn_act_fwd", s, (double)B * N * 2.0 * (dtype == MMSA_F32 ? 4 : 2));
  dim3 block(kBnCols, kBnRows);
  if (B <= kBnRegRows * kBnRows) {
    MMSA_DISPATCH_DTYPE(dtype, T, (bn_act_fwd_small_kernel<T><<<(unsigned)ceil_div(N, kBnCols), block, 0, s>>>(
        (int)B, (int)N, order, (const float*)x, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
        training, dropout_p, keep_mask, mask_given, seed, offset, rng_state, (T*)y,
        (bf16*)(dtype == MMSA_F32 ? y_lp : nullptr), save_mean, save_rstd)));
    MMSA_LAUNCH_CHECK("bn_act_fwd_small_kernel");
    return MMSA_OK;
  }
  MMSA_DISPATCH_DTYPE(dtype, T, (bn_act_fwd_kernel<T><<<(unsigned)ceil_div(N, kBnCols), block, 0, s>>>(
      B, (int)N, order, (const float*)x, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
      training, dropout_p,
      keep_mask, mask_given, seed, offset, rng_state, (T*)y, (bf16*)(dtype == MMSA_F32 ? y_lp : nullptr), save_mean, save_rstd)));
  MMSA_LAUNCH_CHECK("bn_act_fwd_kernel");
  return MMSA_OK;
}

int mmsa_bn_act_bwd(int dtype, int64_t B, int64_t N, int order, const void* x, const void* dy,
                    const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
                    int training, float dropout_p, const uint8_t* keep_mask, void* dx, float* dgamma,
                    float* dbeta, float* dbias_prev, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(B > 0 && N > 0, "mmsa_bn_act_bwd: empty input");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("bn_act_bwd", s, (double)B * N * 3.0 * (dtype == MMSA_F32 ? 4 : 2));
  dim3 block(kBnCols, kBnRows);
  if (B <= kBnRegRows * kBnRows) {
    MMSA_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_small_kernel<T><<<(unsigned)ceil_div(N, kBnCols), block, 0, s>>>(
        (int)B, (int)N, order, (const float*)x, (const float*)dy, gamma, beta, save_mean, save_rstd, training, dropout_p,
        keep_mask, (T*)dx, dgamma, dbeta, dbias_prev)));
    MMSA_LAUNCH_CHECK("bn_act_bwd_small_kernel");
    return MMSA_OK;
  }
  MMSA_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_kernel<T><<<(unsigned)ceil_div(N, kBnCols), block, 0, s>>>(
      B, (int)N, order, (const float*)x, (const float*)dy, gamma, beta, save_mean, save_rstd, training, dropout_p, keep_mask,
      (T*)dx, dgamma, dbeta, dbias_prev)));
  MMSA_LAUNCH_CHECK("bn_act_bwd_kernel");
  return MMSA_OK;
}

int mmsa_ce_fwd(int64_t B, int64_t C, const float* logits, const int64_t* labels, const float* addend, int64_t n_add,
                float* loss, int64_t* pred, float* row_loss, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(B > 0 && C > 0 && C <= 64, "mmsa_ce_fwd: bad shape B=%lld C=%lld", (long long)B, (long long)C);
  MMSA_REQUIRE(n_add >= 0 && n_add <= 64 && (n_add == 0 || addend != nullptr), "mmsa_ce_fwd: bad addend (n_add=%lld)",
               (long long)n_add);
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("ce_fwd", s, (double)B * (C * 4.0 + 20.0));
  if (B <= kCeOneMax) {
    ce_fwd_one_kernel<<<1, 1024, 0, s>>>(B, (int)C, logits, labels, pred, row_loss, addend, (int)n_add, loss);
    MMSA_LAUNCH_CHECK("ce_fwd_one_kernel");
    return MMSA_OK;
  }
  ce_fwd_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(B, (int)C, logits, labels, pred, row_loss);
  MMSA_LAUNCH_CHECK("ce_fwd_kernel");
  sum_scale_add_kernel<<<1, 256, 0, s>>>(row_loss, B, 1.f / (float)B, addend, (int)n_add, loss);
  MMSA_LAUNCH_CHECK("sum_scale_add_kernel");
  return MMSA_OK;
}

int mmsa_ce_bwd(int dtype, int64_t B, int64_t C, const float* logits, const int64_t* labels, const float* dloss,
                void* dlogits, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(B > 0 && C > 0 && C <= 64, "mmsa_ce_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("ce_bwd", s, (double)B * (C * 8.0 + 8.0));
  MMSA_DISPATCH_DTYPE(dtype, T, (ce_bwd_kernel<T><<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(B, (int)C, logits, labels, dloss, (T*)dlogits)));
  MMSA_LAUNCH_CHECK("ce_bwd_kernel");
  return MMSA_OK;
}

int mmsa_contrastive_fwd(int kind, int64_t B, int64_t Bg, int64_t row_offset, const float* sim,
                         const int64_t* labels_rows, const int64_t* labels_cols, const float* temperature,
                         float temperature_const, int64_t denom, const float* weight, float* row_stats, float* row_loss,
                         float* loss, float* loss_raw, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(kind >= 0 && kind <= 2, "mmsa_contrastive_fwd: bad kind %d", kind);
  MMSA_REQUIRE(B > 0 && Bg > 0 && denom > 0, "mmsa_contrastive_fwd: empty input");
  MMSA_REQUIRE(kind == MMSA_LOSS_NTXENT || (labels_rows && labels_cols), "mmsa_contrastive_fwd: labels required");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("contrastive_fwd", s, (double)B * Bg * 4.0);
  contrastive_fwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, s>>>(kind, B, Bg, row_offset, sim, labels_rows,
                                                                  labels_cols, temperature, temperature_const,
                                                                  row_stats, row_loss);
  MMSA_LAUNCH_CHECK("contrastive_fwd_kernel");
  contrastive_finish_kernel<<<1, 256, 0, s>>>(row_loss, B, 1.f / (float)denom, weight, loss, loss_raw);
  MMSA_LAUNCH_CHECK("contrastive_finish_kernel");
  return MMSA_OK;
}

int mmsa_contrastive_bwd(int kind, int64_t B, int64_t Bg, int64_t row_offset, const float* sim,
                         const int64_t* labels_rows, const int64_t* labels_cols, const float* temperature,
                         float temperature_const, int64_t denom, const float* row_stats, const float* dloss,
                         const float* weight, const float* loss_raw, void* G, int g_dtype, float* dtemp_rows, float* dtemp,
                         float* dweight, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(kind >= 0 && kind <= 2, "mmsa_contrastive_bwd: bad kind %d", kind);
  MMSA_REQUIRE(B > 0 && Bg > 0 && denom > 0, "mmsa_contrastive_bwd: empty input");
  MMSA_REQUIRE(dweight == nullptr || loss_raw != nullptr, "mmsa_contrastive_bwd: dweight needs the unweighted loss");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("contrastive_bwd", s, (double)B * Bg * (4.0 + (g_dtype == MMSA_F32 ? 4 : 2)));
  float inv_denom = 1.f / (float)denom;
  if (g_dtype == MMSA_F32)
    contrastive_bwd_kernel<float><<<(unsigned)ceil_div(B, 8), 256, 0, s>>>(
        kind, B, Bg, row_offset, sim, labels_rows, labels_cols, temperature, temperature_const, inv_denom,
        row_stats, dloss, weight, (float*)G, dtemp_rows);
  else
    contrastive_bwd_kernel<bf16><<<(unsigned)ceil_div(B, 8), 256, 0, s>>>(
        kind, B, Bg, row_offset, sim, labels_rows, labels_cols, temperature, temperature_const, inv_denom,
        row_stats, dloss, weight, (bf16*)G, dtemp_rows);
  MMSA_LAUNCH_CHECK("contrastive_bwd_kernel");
  if ((dtemp_rows && dtemp) || dweight) {
    contrastive_bwd_finish_kernel<<<1, 256, 0, s>>>(dtemp_rows, B, dtemp, dloss, loss_raw, dweight);
    MMSA_LAUNCH_CHECK("contrastive_bwd_finish_kernel");
  }
  return MMSA_OK;
}

int mmsa_sumsq(const float* x, int64_t n, float* partials, int64_t nblk, float* out, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(nblk > 0 && nblk <= 4096, "mmsa_sumsq: nblk out of range");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("sumsq", s, (double)n * 4.0);
  sumsq_partial_kernel<<<(unsigned)nblk, 256, 0, s>>>(x, n, partials);
  MMSA_LAUNCH_CHECK("sumsq_partial_kernel");
  sum_scale_kernel<<<1, 256, 0, s>>>(partials, nblk, 1.f, out);
  MMSA_LAUNCH_CHECK("sum_scale_kernel");
  return MMSA_OK;
}

int mmsa_clip_adamw(float* p, const float* g, float* m, float* v, int64_t n, const float* gradsq,
                    float max_norm, double lr, double beta1, double beta2, double eps, double weight_decay,
                    int64_t step, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(step >= 1, "mmsa_clip_adamw: step starts at 1");
  if (n == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("clip_adamw", s, (double)n * 28.0);
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  clip_adamw_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, g, m, v, n, gradsq, max_norm, (float)(1.0 - lr * weight_decay),
                                                    (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                    (float)eps, (float)(lr / bc1), (float)sqrt(bc2));
  MMSA_LAUNCH_CHECK("clip_adamw_kernel");
  return MMSA_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ stand-alone dropout
// nn.Dropout(0.5) after a ReLU with no BatchNorm in between (Classifier.shared, ME-MHACL/model.py:105-109).
// Backward is the same kernel applied to dy with the saved mask.
namespace mmsa {
template <typename T>
__global__ void dropout_kernel(int64_t n, const float* __restrict__ x, float p, uint8_t* __restrict__ keep_mask,
                               int mask_given, uint64_t seed, uint64_t offset,
                               const uint64_t* __restrict__ rng_state, T* __restrict__ y) {
  if (rng_state != nullptr) { seed = rng_state[0]; offset += rng_state[1]; }
  const float scale = 1.f / (1.f - p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t keep;
    if (mask_given) keep = keep_mask[i];
    else {
      uint32_t rnd = philox_first(seed, offset + (uint64_t)i);
      keep = ((float)(rnd >> 8) * (1.f / 16777216.f)) >= p ? 1 : 0;
      keep_mask[i] = keep;
    }
    y[i] = from_f<T>(keep ? to_f(x[i]) * scale : 0.f);
  }
}
}  // namespace mmsa

extern "C" int mmsa_dropout(int dtype, int64_t n, const void* x, float p, uint8_t* keep_mask, int mask_given,
                            uint64_t seed, uint64_t offset, const uint64_t* rng_state, void* y, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(p >= 0.f && p < 1.f && keep_mask != nullptr, "mmsa_dropout: bad arguments");
  if (n == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof("dropout", s, (double)n * (2.0 * (dtype == MMSA_F32 ? 4 : 2) + 1.0));
  int64_t blocks = mmsa::ceil_div(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  MMSA_DISPATCH_DTYPE(dtype, T, (mmsa::dropout_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(
      n, (const float*)x, p, keep_mask, mask_given, seed, offset, rng_state, (T*)y)));
  MMSA_LAUNCH_CHECK("dropout_kernel");
  return MMSA_OK;
}

// ------------------------------------------------------------------ device-resident Philox stream position
// state = {seed, position}: the dropout kernels above add `position` to their by-value offset when given the state
// pointer, and this one-thread kernel moves the position on by what a step consumed.  Both are ordinary kernel nodes,
// so a captured CUDA graph draws a fresh mask on every replay (the by-value seed/offset alone would be frozen at capture).
namespace mmsa {
__global__ void rng_advance_kernel(uint64_t* __restrict__ state, uint64_t n) { state[1] += n; }
}  // namespace mmsa

extern "C" int mmsa_rng_advance(uint64_t* rng_state, uint64_t n, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(rng_state != nullptr, "mmsa_rng_advance: null state");
  mmsa::rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng_state, n);
  MMSA_LAUNCH_CHECK("rng_advance_kernel");
  return MMSA_OK;
}
