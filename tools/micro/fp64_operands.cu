// Microbenchmark: single-warp FP64 issue rate with realistic operand patterns (three distinct register operands per
// DFMA, DMUL/DADD mixes) on sm_100a.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, int MODE>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[CH], y[CH], w[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { x[c] = a + c + threadIdx.x; y[c] = b + 1e-9 * (c + threadIdx.x); w[c] = a - 1e-9 * (c + threadIdx.x); }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (MODE == 0) x[c] = fma(x[c], b, a);                       // two loop-invariant operands
                if (MODE == 1) x[c] = fma(x[c], y[c], w[c]);                 // three distinct register operands
                if (MODE == 2) x[c] = fma(x[c], y[c], x[(c + 1) % CH]);      // three distinct, all changing
                if (MODE == 3) x[c] = x[c] * y[c];                           // DMUL
                if (MODE == 4) x[c] = x[c] + w[c];                           // DADD
                if (MODE == 5) { x[c] = fma(x[c], y[c], w[c]); y[c] = y[c] * b; }   // two instructions per chain step
            }
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; ++c) s += x[c] + y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH, int MODE> void run(int threads) {
    double* out; long long* cyc; cudaMalloc(&out, 8 * threads); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int k = 0; k < 2; ++k) chain<CH, MODE><<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int per = MODE == 5 ? 2 : 1;
    printf("mode=%d chains=%d threads=%d : %.2f cycles per FP64 instruction per warp\n", MODE, CH, threads, (double)h / (iters * 8 * CH * per));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<8, 0>(32); run<8, 1>(32); run<8, 2>(32); run<8, 3>(32); run<8, 4>(32); run<8, 5>(32);
    run<4, 1>(32); run<12, 1>(32);
    run<8, 1>(128); run<8, 1>(256); run<8, 2>(256); run<8, 1>(512);
    return 0;
}
