// Which hardware warp slot (%warpid) do the warps of co-resident 4-warp blocks get?  (scheduler partition = slot % 4)
// nvcc -gencode arch=compute_100a,code=sm_100a -o warp_slots warp_slots.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int* out, long long spin) {
    extern __shared__ unsigned char sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2] = smid; out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2 + 1] = wid; }
    if (threadIdx.x == 0) sm[0] = 1;
}
int main() {
    const int grid = 444, threads = 128, smem = 69264;
    int* d; cudaMalloc(&d, grid * 4 * 2 * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<grid, threads, smem>>>(d, 2000000);
    cudaDeviceSynchronize();
    static int h[444 * 4 * 2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int b = 0; b < grid; ++b) {
        if (h[b * 8] > 2) continue;                       // print the blocks of SMs 0-2
        printf("block %3d sm %d slots %d %d %d %d\n", b, h[b * 8], h[b * 8 + 1], h[b * 8 + 3], h[b * 8 + 5], h[b * 8 + 7]);
    }
    return 0;
}
