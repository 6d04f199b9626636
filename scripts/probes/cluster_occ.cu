// probe: how many clusters of size S (192 threads, ~205 KB smem per CTA) can be co-resident on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char sm[]; if (p) p[0] = sm[0]; }
int main() {
  int smem = 205 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int S = 1; S <= 16; ++S) {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(S * 64);
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d (%d CTAs) %s\n", S, n, n * S, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
