// api_core.cu -- error reporting, device gate, launch accounting, dtype casts.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <string.h>
#include "common.cuh"

namespace mmsa {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-launch timing (bench.py roofline leg) ----
int g_prof_on = 0;
struct ProfRec { int entry; cudaEvent_t a, b; };
struct ProfEntry { char name[48]; int64_t count; double ms; double work; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<ProfEntry> g_prof_entries;
static std::mutex g_prof_mu;
static thread_local int g_prof_open = -1;

void prof_begin(const char* name, cudaStream_t s, double work) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int e = -1;
  for (size_t i = 0; i < g_prof_entries.size(); ++i)
    if (strncmp(g_prof_entries[i].name, name, sizeof(g_prof_entries[i].name) - 1) == 0) { e = (int)i; break; }
  if (e < 0) {
    ProfEntry pe{};
    strncpy(pe.name, name, sizeof(pe.name) - 1);
    g_prof_entries.push_back(pe);
    e = (int)g_prof_entries.size() - 1;
  }
  g_prof_entries[e].count += 1;
  g_prof_entries[e].work += work;
  ProfRec r{e, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  g_prof_recs.push_back(r);
  g_prof_open = (int)g_prof_recs.size() - 1;
}
void prof_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_open >= 0 && g_prof_open < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[g_prof_open].b, s);
  g_prof_open = -1;
}

bool device_ok() {
  static unsigned char state[kMaxDevices] = {};      // per device: 0 unknown, 1 sm_100, 2 something else
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev < 0 || dev >= kMaxDevices) return false;
  if (state[dev] == 0) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    state[dev] = (major == 10) ? 1 : 2;
  }
  return state[dev] == 1;
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t nvec = n / 8;
  for (int64_t v = i; v < nvec; v += stride) {
    float f[8];
    if (sizeof(S) == 4) {
      load_vec<float>(reinterpret_cast<const float*>(src) + v * 8, f);
      load_vec<float>(reinterpret_cast<const float*>(src) + v * 8 + 4, f + 4);
    } else {
      load_vec<bf16>(reinterpret_cast<const bf16*>(src) + v * 8, f);
    }
    if (sizeof(D) == 4) {
      store_vec<float>(reinterpret_cast<float*>(dst) + v * 8, f);
      store_vec<float>(reinterpret_cast<float*>(dst) + v * 8 + 4, f + 4);
    } else {
      store_vec<bf16>(reinterpret_cast<bf16*>(dst) + v * 8, f);
    }
  }
  for (int64_t k = nvec * 8 + i; k < n; k += stride) dst[k] = from_f<D>(to_f(src[k]));
}

// one launch casting many tensors (the fp32 master weights -> bf16 operand copies of a training step)
constexpr int kCastMaxTensors = 64;
constexpr int kCastChunk = 8192;          // elements per block
struct CastTable {
  const void* src[kCastMaxTensors];
  void* dst[kCastMaxTensors];
  long long n[kCastMaxTensors];
  int blk_start[kCastMaxTensors + 1];
  int count;
};
template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_multi_kernel(const CastTable tb) {
  int t = 0;
  while (t + 1 < tb.count && (int)blockIdx.x >= tb.blk_start[t + 1]) ++t;
  const long long off = (long long)((int)blockIdx.x - tb.blk_start[t]) * kCastChunk;
  const long long n = tb.n[t];
  const S* src = reinterpret_cast<const S*>(tb.src[t]);
  D* dst = reinterpret_cast<D*>(tb.dst[t]);
  const long long end = off + kCastChunk < n ? off + kCastChunk : n;
  if ((n & 7) == 0) {
    for (long long i = off + threadIdx.x * 8; i < end; i += 256 * 8) {
      float f[8];
      if (sizeof(S) == 4) {
        load_vec<float>(reinterpret_cast<const float*>(src) + i, f);
        load_vec<float>(reinterpret_cast<const float*>(src) + i + 4, f + 4);
      } else load_vec<bf16>(reinterpret_cast<const bf16*>(src) + i, f);
      if (sizeof(D) == 4) {
        store_vec<float>(reinterpret_cast<float*>(dst) + i, f);
        store_vec<float>(reinterpret_cast<float*>(dst) + i + 4, f + 4);
      } else store_vec<bf16>(reinterpret_cast<bf16*>(dst) + i, f);
    }
  } else {
    for (long long i = off + threadIdx.x; i < end; i += 256) dst[i] = from_f<D>(to_f(src[i]));
  }
}

// fp32 -> two-term bf16 split (hi = bf16(x), lo = bf16(x - hi)), laid out so that ONE bf16 tensor-core GEMM over a
// three-times longer reduction reproduces the fp32 product to ~2^-16:  sum_k a_k b_k ~= hi_a.hi_b + hi_a.lo_b + lo_a.hi_b.
// The A-side operand repeats (hi, hi, lo), the B-side (hi, lo, hi); the three copies are concatenated along the
// reduction axis: along columns for a K-major operand (dst_col [R, 3C]) and along rows for an MN-major one
// (dst_row [3R, C]).  Used by the InfoNCE similarity GEMM and its two backward GEMMs in bf16 mode.
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ src, int64_t R, int64_t C, int64_t ld, bf16* __restrict__ dst_col, int b_side_col,
              bf16* __restrict__ dst_row, int b_side_row) {
  const int64_t total = R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C, c = i % C;
    const float x = src[r * ld + c];
    const bf16 hi = __float2bfloat16_rn(x);
    const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    if (dst_col) {
      bf16* d = dst_col + r * 3 * C + c;
      d[0] = hi; d[C] = b_side_col ? lo : hi; d[2 * C] = b_side_col ? hi : lo;
    }
    if (dst_row) {
      bf16* d = dst_row + r * C + c;
      d[0] = hi; d[R * C] = b_side_row ? lo : hi; d[2 * R * C] = b_side_row ? hi : lo;
    }
  }
}

}  // namespace mmsa

using namespace mmsa;

extern "C" {

const char* mmsa_version(void) { return "mmsa-b200 0.1 (sm_100a)"; }
const char* mmsa_last_error(void) { return g_err; }
int mmsa_check_device(void) {
  if (!device_ok()) {
    set_error("mmsa: current CUDA device is not sm_100 (B200); no fallback");
    return MMSA_ERR_DEVICE;
  }
  return MMSA_OK;
}
int64_t mmsa_launch_count(void) { return g_launches.load(); }

void mmsa_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (on) {
    for (auto& r : g_prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof_recs.clear();
    g_prof_entries.clear();
  }
  g_prof_on = on ? 1 : 0;
}

int mmsa_prof_collect(char* names, int name_stride, int64_t* counts, double* ms, double* work, int max_entries) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& e : g_prof_entries) e.ms = 0.0;
  for (auto& r : g_prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) g_prof_entries[r.entry].ms += (double)t;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  (void)cudaGetLastError();
  g_prof_recs.clear();
  int n = (int)g_prof_entries.size() < max_entries ? (int)g_prof_entries.size() : max_entries;
  for (int i = 0; i < n; ++i) {
    if (names && name_stride > 0) {
      strncpy(names + (size_t)i * name_stride, g_prof_entries[i].name, (size_t)name_stride - 1);
      names[(size_t)i * name_stride + name_stride - 1] = 0;
    }
    counts[i] = g_prof_entries[i].count;
    ms[i] = g_prof_entries[i].ms;
    work[i] = g_prof_entries[i].work;
  }
  return n;
}

int mmsa_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (n == 0) return MMSA_OK;
  MMSA_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "mmsa_cast: pointers must be 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int threads = 256;
  int64_t blocks = ceil_div(ceil_div(n, 8), threads);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  ProfScope prof("cast", s, (double)n * ((sdt == MMSA_F32 ? 4 : 2) + (ddt == MMSA_F32 ? 4 : 2)));
  if (sdt == MMSA_F32 && ddt == MMSA_BF16)
    cast_kernel<float, bf16><<<(unsigned)blocks, threads, 0, s>>>((const float*)src, (bf16*)dst, n);
  else if (sdt == MMSA_BF16 && ddt == MMSA_F32)
    cast_kernel<bf16, float><<<(unsigned)blocks, threads, 0, s>>>((const bf16*)src, (float*)dst, n);
  else if (sdt == MMSA_F32 && ddt == MMSA_F32)
    cast_kernel<float, float><<<(unsigned)blocks, threads, 0, s>>>((const float*)src, (float*)dst, n);
  else if (sdt == MMSA_BF16 && ddt == MMSA_BF16)
    cast_kernel<bf16, bf16><<<(unsigned)blocks, threads, 0, s>>>((const bf16*)src, (bf16*)dst, n);
  else {
    set_error("mmsa_cast: bad dtypes %d -> %d", sdt, ddt);
    return MMSA_ERR_ARG;
  }
  MMSA_LAUNCH_CHECK("cast_kernel");
  return MMSA_OK;
}

int mmsa_split3(const float* src, int64_t R, int64_t C, int64_t ld, void* dst_col, int b_side_col, void* dst_row,
                int b_side_row, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(R >= 0 && C > 0 && ld >= C && src != nullptr && (dst_col != nullptr || dst_row != nullptr), "mmsa_split3: bad arguments");
  if (R == 0) return MMSA_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = ceil_div(R * C, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope prof("split3", s, (double)R * C * (4.0 + 6.0 * ((dst_col ? 1 : 0) + (dst_row ? 1 : 0))));
  split3_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, R, C, ld, (bf16*)dst_col, b_side_col, (bf16*)dst_row, b_side_row);
  MMSA_LAUNCH_CHECK("split3_kernel");
  return MMSA_OK;
}

int mmsa_cast_multi(int count, const void* const* src_host, void* const* dst_host, const int64_t* numel_host,
                    int sdt, int ddt, void* stream) {
  MMSA_REQUIRE_DEVICE();
  MMSA_REQUIRE(count >= 0, "mmsa_cast_multi: bad count");
  cudaStream_t s = (cudaStream_t)stream;
  for (int base = 0; base < count; base += kCastMaxTensors) {
    CastTable tb{};
    int c = count - base < kCastMaxTensors ? count - base : kCastMaxTensors;
    int blocks = 0;
    double bytes = 0;
    for (int i = 0; i < c; ++i) {
      MMSA_REQUIRE(((uintptr_t)src_host[base + i] % 16 == 0) && ((uintptr_t)dst_host[base + i] % 16 == 0),
                   "mmsa_cast_multi: pointers must be 16B aligned");
      tb.src[i] = src_host[base + i]; tb.dst[i] = dst_host[base + i]; tb.n[i] = numel_host[base + i];
      tb.blk_start[i] = blocks;
      blocks += (int)ceil_div(numel_host[base + i], kCastChunk);
      bytes += (double)numel_host[base + i] * ((sdt == MMSA_F32 ? 4 : 2) + (ddt == MMSA_F32 ? 4 : 2));
    }
    tb.blk_start[c] = blocks;
    tb.count = c;
    if (blocks == 0) continue;
    ProfScope prof("cast_multi", s, bytes);
    if (sdt == MMSA_F32 && ddt == MMSA_BF16) cast_multi_kernel<float, bf16><<<blocks, 256, 0, s>>>(tb);
    else if (sdt == MMSA_BF16 && ddt == MMSA_F32) cast_multi_kernel<bf16, float><<<blocks, 256, 0, s>>>(tb);
    else { set_error("mmsa_cast_multi: bad dtypes %d -> %d", sdt, ddt); return MMSA_ERR_ARG; }
    MMSA_LAUNCH_CHECK("cast_multi_kernel");
  }
  return MMSA_OK;
}

}  // extern "C"
