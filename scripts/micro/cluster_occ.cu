// How many clusters of a 1-CTA-per-SM kernel can be co-resident on this GPU, by cluster size and dynamic shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_occ scripts/micro/cluster_occ.cu && /tmp/cluster_occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(672, 1) k(int* p) { if (p) p[0] = 1; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smems[] = {0, 100 * 1024, 163840 + 1280, 229376 + 1280};
  for (int smem : smems) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cl : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(16 * cl); cfg.blockDim = dim3(672); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %6d cluster %2d: max active clusters %3d (= %3d CTAs) %s\n", smem, cl, n, n * cl, e ? cudaGetErrorString(e) : "");
    }
  }
  return 0;
}
