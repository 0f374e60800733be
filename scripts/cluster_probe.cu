// How many CTAs of the fused GEMM's footprint (256 threads, ~200 KB dynamic shared memory, 1 CTA per SM) can
// be co-resident on this GPU at cluster sizes 1 / 2 / 4 / 8?  cudaOccupancyMaxActiveClusters answers for the
// GPC layout of the device it runs on.  Evidence for DESIGN.md section 3a (why kernel (a) stays at CTA pairs).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/build/cluster_probe scripts/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void footprint_kernel(int* out) {
  extern __shared__ int smem[];
  if (out && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(footprint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(footprint_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"smem_per_cta\": %d", prop.name, prop.multiProcessorCount, smem);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cs;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int clusters = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&clusters, footprint_kernel, &cfg);
    if (e != cudaSuccess) { clusters = -1; cudaGetLastError(); }
    printf(", \"cluster%d\": {\"max_active_clusters\": %d, \"ctas\": %d}", cs, clusters, clusters * cs);
  }
  printf("}\n");
  return 0;
}
