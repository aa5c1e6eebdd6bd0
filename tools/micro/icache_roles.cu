// Microbenchmark: do warps running DIFFERENT loop bodies on one SM slow each other down (instruction cache)?
// 5 warps per block, 3 blocks per SM (70 KB smem each), every warp runs a 4-chain FP64 loop body of ~BODY
// instructions; "same": all warps run variant 0; "diff": warp w runs variant w (distinct code).
#include <cstdio>
#include <cuda_runtime.h>
template <int V, int BODY>
__device__ __forceinline__ void body(double& a, double& b, double& c, double& d, double k) {
#pragma unroll
    for (int i = 0; i < BODY / 4; ++i) {
        a = fma(a, k, 1.0 + V + i * 1e-3); b = fma(b, k, 2.0 + V + i * 1e-3);
        c = fma(c, k, 3.0 + V + i * 1e-3); d = fma(d, k, 4.0 + V + i * 1e-3);
    }
}
template <int BODY, bool DIFF>
__global__ void __launch_bounds__(160) roles(double* out, long long* cyc, int iters, double k) {
    extern __shared__ double sm[];
    const int w = threadIdx.x >> 5;
    double a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int v = DIFF ? w : 0;
        switch (v) {
            case 0: body<0, BODY>(a, b, c, d, k); break;
            case 1: body<1, BODY>(a, b, c, d, k); break;
            case 2: body<2, BODY>(a, b, c, d, k); break;
            case 3: body<3, BODY>(a, b, c, d, k); break;
            default: body<4, BODY>(a, b, c, d, k); break;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + sm[0] * 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int BODY, bool DIFF> void run(const char* name) {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 444 * 160); cudaMalloc(&cyc, 8);
    auto k = roles<BODY, DIFF>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    const int iters = 400;
    k<<<444, 160, 70 * 1024>>>(out, cyc, iters, 0.999999);
    k<<<444, 160, 70 * 1024>>>(out, cyc, iters, 0.999999);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s body=%4d instr: %.2f cycles per instruction per warp (15 warps/SM; ideal 2.0 x 3.75 warps/SMSP = 7.5)\n", name, BODY, (double)h / (iters * (double)BODY));
    cudaFree(out); cudaFree(cyc);
}
__global__ void whereami(int* out) {
    unsigned smid, warpid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    extern __shared__ double sm[];
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 5 + (threadIdx.x >> 5)) * 2] = smid; out[(blockIdx.x * 5 + (threadIdx.x >> 5)) * 2 + 1] = warpid; }
    for (volatile int i = 0; i < 100000; ++i) {}
}
int main() {
    run<128, false>("same code"); run<128, true>("different code per warp");
    run<256, false>("same code"); run<256, true>("different code per warp");
    run<384, false>("same code"); run<384, true>("different code per warp");
    run<512, false>("same code"); run<512, true>("different code per warp");
    run<1024, false>("same code"); run<1024, true>("different code per warp");
    int* o; cudaMalloc(&o, 444 * 5 * 2 * 4); cudaFuncSetAttribute(whereami, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    whereami<<<444, 160, 70 * 1024>>>(o);
    static int h[444 * 10]; cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
    for (int sm = 0; sm < 2; ++sm) { printf("SM %d:", sm); for (int b = 0; b < 444; ++b) if (h[b * 10] == sm) { printf("  block %d warps->hw slots", b); for (int w = 0; w < 5; ++w) printf(" %d", h[(b * 5 + w) * 2 + 1]); } printf("\n"); }
    return 0;
}
