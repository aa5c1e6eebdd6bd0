// Microbenchmark: FP64 dependent-issue latency and per-warp throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = a + c + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], b, a);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH> void run(int blocks, int threads) {
    double* out; long long* cyc; cudaMalloc(&out, 8 * blocks * threads); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    chain<CH><<<blocks, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    chain<CH><<<blocks, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("chains=%d blocks=%d threads=%d : %.2f cycles per DFMA-round (%.2f cycles per DFMA per warp)\n", CH, blocks, threads, (double)h / (iters * 8), (double)h / (iters * 8 * CH));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(1, 32); run<2>(1, 32); run<4>(1, 32); run<8>(1, 32);
    run<1>(1, 128); run<1>(1, 512); run<1>(1, 1024); run<4>(1, 512); run<4>(1, 1024);
    return 0;
}
