// Where does the block scheduler put 200 CTAs (2 resident per SM at most) on 148 SMs?  Prints, per SM, the block ids it received.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/probe_placement tools/probe_placement.cu
#include <cstdio>
#include <vector>
#include <map>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) probe(int* smid, long long* t0, int spin_us) {
  extern __shared__ unsigned char sm[];
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) { smid[lin] = (int)id; t0[lin] = t; sm[0] = 1; }
  long long now = t;
  while (now - t < (long long)spin_us * 1000) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
}
int main() {
  for (int variant = 0; variant < 2; ++variant) {
    dim3 grid = variant == 0 ? dim3(20, 5, 2) : dim3(200, 1, 1);
    const int n = grid.x * grid.y * grid.z;
    int* d_smid; long long* d_t0;
    cudaMalloc(&d_smid, n * 4); cudaMalloc(&d_t0, n * 8);
    const size_t smem = 100 * 1024;  // two CTAs per SM fit, three do not
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<grid, 256, smem>>>(d_smid, d_t0, 200);
    cudaDeviceSynchronize();
    std::vector<int> smid(n); std::vector<long long> t0(n);
    cudaMemcpy(smid.data(), d_smid, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(t0.data(), d_t0, n * 8, cudaMemcpyDeviceToHost);
    std::map<int, std::vector<int>> per;
    long long tmin = t0[0];
    for (int i = 0; i < n; ++i) { per[smid[i]].push_back(i); if (t0[i] < tmin) tmin = t0[i]; }
    int doubles = 0, first148_distinct = 0, late_second = 0;
    std::map<int, int> seen;
    for (int i = 0; i < 148 && i < n; ++i) if (!seen[smid[i]]++) ++first148_distinct;
    for (auto& kv : per) if (kv.second.size() > 1) { ++doubles; for (size_t k = 1; k < kv.second.size(); ++k) if (kv.second[k] >= 148) ++late_second; }
    long long tmax = 0; for (int i = 0; i < n; ++i) if (t0[i] - tmin > tmax) tmax = t0[i] - tmin;
    printf("grid (%d,%d,%d): %d blocks on %zu SMs, %d SMs hold two; first 148 linear ids on %d distinct SMs; second residents with id >= 148: %d; start-time spread %lld ns\n",
           grid.x, grid.y, grid.z, n, per.size(), doubles, first148_distinct, late_second, tmax);
    printf("  doubled SMs hold:");
    int shown = 0;
    for (auto& kv : per) if (kv.second.size() > 1 && shown++ < 12) printf(" sm%d{%d,%d}", kv.first, kv.second[0], kv.second[1]);
    printf("\n");
    cudaFree(d_smid); cudaFree(d_t0);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
