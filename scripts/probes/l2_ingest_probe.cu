// l2_ingest_probe.cu -- how many operand bytes per second can the SMs of a B200 pull through L2 with TMA, and does TMA
// MULTICAST across a thread-block cluster raise that ceiling?  (Decides whether sharing the weight tile between two CTA
// pairs is worth building into the GEMM.)  No tensor-core work: a producer thread streams tiles into a shared-memory
// ring, a consumer thread frees each stage as soon as it has landed.
//   mode 0  unicast : every CTA loads A (16 KB, its own rows) + B (16 KB, from a small L2-resident matrix all CTAs share)
//   mode 1  mcast   : cluster of CS CTAs; B is loaded in CS slices of 16/CS KB, each slice multicast to the whole cluster
//   mode 2  A only  : 16 KB per stage (floor for the activation stream alone)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_ingest_probe l2_ingest_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static constexpr int STAGES = 6;
static constexpr int A_BYTES = 16384, B_BYTES = 16384, STAGE_BYTES = A_BYTES + B_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar) { asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory"); }
__global__ void fill_kernel(uint32_t* p, size_t n, uint32_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = (x & 0x3FFF3FFFu) | 0x3C003C00u;     // two finite bf16 values with random mantissas
  }
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask) : "memory");
}

// A: [rows_a, 64*kblocks] bf16 (box 64 x 128 rows = 16 KB); B: [128, 64*kblocks] (box 64 x (128/CS) rows)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int mode, int cs, int kblocks, int tiles,
             int rows_a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  uint32_t crank = 0;
  if (cs > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), mode == 1 ? (uint32_t)cs : 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cs > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int bslice = 128 / (mode == 1 ? cs : 1);               // rows of B this CTA issues
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < tiles; ++t) {
      const int row0 = (int)(((long long)blockIdx.x * tiles + t) * 128 % rows_a);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty(stage), phase ^ 1u);
        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
        const uint32_t bytes = mode == 2 ? A_BYTES : STAGE_BYTES;
        mbar_expect_tx(full(stage), bytes);
        tma_load_2d(sa, &tmA, kb * 64, row0, full(stage));
        if (mode == 0) tma_load_2d(sb, &tmB, kb * 64, 0, full(stage));
        else if (mode == 1) tma_load_2d_mc(sb + crank * (bslice * 128), &tmB, kb * 64, (int)crank * bslice, full(stage), (uint16_t)((1u << cs) - 1u));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < tiles; ++t)
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full(stage), phase);
        if (mode == 1) { for (int r = 0; r < cs; ++r) mbar_arrive_remote(mapa_u32(empty(stage), (uint32_t)r)); }   // every producer of the cluster writes into this stage
        else mbar_arrive(empty(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
  }
  __syncthreads();
  if (cs > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeTiledFn fn, CUtensorMap* m, void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn fn = (EncodeTiledFn)fnp;
  const int kblocks = 12, rows_a = 262144;                   // A = 403 MB: far beyond the 126 MB L2, streams from HBM                    // K = 768: the E x E GEMM family
  const uint64_t K = 64ull * kblocks;
  __nv_bfloat16 *A, *B;
  CK(cudaMalloc(&A, (size_t)rows_a * K * 2)); CK(cudaMalloc(&B, (size_t)128 * K * 2));
  fill_kernel<<<1024, 256>>>((uint32_t*)A, (size_t)rows_a * K / 2, 1u);       // incompressible operands (zeros would be)
  fill_kernel<<<64, 256>>>((uint32_t*)B, (size_t)128 * K / 2, 2u);
  CK(cudaDeviceSynchronize());
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int tiles = 40;
  struct Cfg { int mode, cs; const char* name; } cfgs[] = {
      {2, 1, "A only 16 KB/stage"}, {0, 1, "unicast A+B 32 KB/stage"}, {0, 2, "unicast, cluster 2"}, {1, 2, "B multicast, cluster 2"},
      {1, 4, "B multicast, cluster 4"}, {1, 8, "B multicast, cluster 8"}};
  for (auto& c : cfgs) {
    CUtensorMap tmA, tmB;
    make_map(fn, &tmA, A, K, rows_a, 64, 128);
    make_map(fn, &tmB, B, K, 128, 64, (uint32_t)(128 / (c.mode == 1 ? c.cs : 1)));
    int grid = 148 / c.cs * c.cs;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, probe_kernel, tmA, tmB, c.mode, c.cs, kblocks, tiles, rows_a));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int w = 0; w < reps; ++w) CK(cudaLaunchKernelEx(&cfg, probe_kernel, tmA, tmB, c.mode, c.cs, kblocks, tiles, rows_a));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double stages = (double)tiles * kblocks;
    const double recv = (c.mode == 2 ? A_BYTES : STAGE_BYTES) * stages * grid;
    printf("%-28s grid %3d: %8.1f us  %6.0f ns/stage  received %5.2f TB/s chip, %5.1f B/clk/SM @1.965GHz\n", c.name, grid, ms * 1e3,
           ms * 1e6 / stages, recv / (ms * 1e-3) / 1e12, recv / grid / (ms * 1e-3) / 1.965e9);
  }
  return 0;
}
