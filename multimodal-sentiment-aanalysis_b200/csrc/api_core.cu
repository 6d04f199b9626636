// api_core.cu -- error reporting, device gate, launch accounting, dtype casts.
#include <stdarg.h>
#include <atomic>
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

bool device_ok() {
  static int cached_dev = -1;
  static bool cached_ok = false;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev == cached_dev) return cached_ok;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  cached_dev = dev;
  cached_ok = (major == 10);
  return cached_ok;
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

int mmsa_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream) {
  MMSA_REQUIRE_DEVICE();
  if (n == 0) return MMSA_OK;
  MMSA_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "mmsa_cast: pointers must be 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int threads = 256;
  int64_t blocks = ceil_div(ceil_div(n, 8), threads);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
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

}  // extern "C"
