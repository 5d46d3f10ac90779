// Largest dynamic shared memory a cluster-of-2 kernel (576 threads, 96 registers, 1 KB static alignment slack) launches with.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_smem_limit.bin tools/cluster_smem_limit.cu  (then run it under gpurun)
// Measured on B200: a trivial kernel launches up to the 232 448 B opt-in limit with or without a cluster; the fused front
// (96 registers, tcgen05, mbarriers) stops at ~224 000 B — see DESIGN.md §3.1(a).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) k(int* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0) smem[0] = 1;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = smem[0];
}
int main() {
  int* d; cudaMalloc(&d, 4);
  int optin = 0; cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0);
  int persm = 0; cudaDeviceGetAttribute(&persm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, 0);
  int reserved = 0; cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, 0);
  printf("optin %d per-SM %d reserved/block %d\n", optin, persm, reserved);
  for (int cl = 1; cl <= 2; ++cl)
    for (int s = 200 * 1024; s <= 233 * 1024; s += 512) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, s);
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaError_t le = e == cudaSuccess ? cudaLaunchKernelEx(&cfg, k, d) : e;
      cudaError_t se = cudaDeviceSynchronize();
      if (le != cudaSuccess || se != cudaSuccess) { printf("cluster %d: first failure at %d B (%s / %s)\n", cl, s, cudaGetErrorString(le), cudaGetErrorString(se)); cudaGetLastError(); break; }
    }
  return 0;
}
