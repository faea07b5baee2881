// Where do the two CTAs of a 2-CTA cluster land when two such CTAs fit on one SM (109 KB of shared memory, 160 threads)?
// Prints how many clusters have both CTAs on the SAME SM.   nvcc -arch=sm_100a -o smid_pairs smid_pairs.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 2) probe(int* smid, long long spin) {
    extern __shared__ unsigned char smem[];
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) smid[blockIdx.x] = static_cast<int>(id);
    smem[threadIdx.x] = 1;
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
}
__global__ void __launch_bounds__(160, 2) probe_nocluster(int* smid, long long spin) {
    extern __shared__ unsigned char smem[];
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) smid[blockIdx.x] = static_cast<int>(id);
    smem[threadIdx.x] = 1;
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
}
int main() {
    const int grid = 296, smem = 109 * 1024;
    int* d;
    cudaMalloc(&d, grid * sizeof(int));
    int h[grid];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_nocluster, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<grid, 160, smem>>>(d, 2000000);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cluster launch: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int same = 0;
    for (int c = 0; c < grid / 2; ++c) same += h[2 * c] == h[2 * c + 1];
    printf("clusters of 2 with both CTAs on one SM: %d of %d\n", same, grid / 2);
    printf("first pairs (smid):");
    for (int c = 0; c < 8; ++c) printf(" (%d,%d)", h[2 * c], h[2 * c + 1]);
    printf("\n");
    probe_nocluster<<<grid, 160, smem>>>(d, 2000000);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int half = 0, adj = 0;
    for (int b = 0; b < grid / 2; ++b) half += h[b] == h[b + grid / 2];
    for (int c = 0; c < grid / 2; ++c) adj += h[2 * c] == h[2 * c + 1];
    printf("plain launch: CTA b and b+148 on one SM: %d of %d; CTA 2c and 2c+1 on one SM: %d of %d\n", half, grid / 2, adj, grid / 2);
    printf("first CTAs (smid):");
    for (int b = 0; b < 8; ++b) printf(" %d", h[b]);
    printf(" ... CTA 148..151:");
    for (int b = 148; b < 152; ++b) printf(" %d", h[b]);
    printf("\n");
    return 0;
}
