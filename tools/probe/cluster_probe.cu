// How many thread-block clusters of a given size are co-resident on this GPU for a kernel shaped like
// the fused sweep (192 threads, ~220 KB dynamic smem, 1 CTA per SM)?  Decides whether 4-CTA clusters
// (two CTA pairs sharing a multicast database tile) can cover all 148 SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_probe tools/probe/cluster_probe.cu && /tmp/cluster_probe
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(192, 1) probe_kernel(int* out) {
  extern __shared__ char smem[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 148);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d -> %d SMs busy (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
