// How many thread-block clusters of a given size can be co-resident on this GPU with ~220 KB of shared memory per CTA?
// (planning input for cluster-wide kernels: a cluster cannot span GPCs)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s: %d SMs\n", pr.name, pr.multiProcessorCount);
    const int smem = 220 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
