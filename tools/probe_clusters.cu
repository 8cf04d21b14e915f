// How many thread-block clusters of a given size are co-resident on this GPU (one CTA per SM: 200 KB of shared memory each)?
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a --cudart shared -o build/probe_clusters tools/probe_clusters.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) p[0] = 1; }
int main() {
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c : {1, 2, 4, 6, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c * 64);
    cfg.blockDim = dim3(288);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
