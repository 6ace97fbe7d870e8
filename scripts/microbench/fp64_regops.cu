// Development microbenchmark: FP64 pipe throughput with REGISTER operands, as real code has them.
// The peak microbenchmark of bench.py (ttvb_fp64_peak) chains c = fma(c, a, b) with a, b kernel
// parameters (constant-bank operands: one register read per DFMA). Here every DFMA reads two or three
// different registers, at 2 warps per scheduler x 8 independent chains (the occupancy of the main
// kernel) and at 4 / 8 warps per scheduler.
//   nvcc -arch=sm_100a -O3 -o fp64_regops fp64_regops.cu && ./fp64_regops
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int iters) {
    double c[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i] = in[i] + threadIdx.x; x[i] = in[8 + i]; y[i] = in[16 + i]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) c[i] = fma(c[i], 0.999999, 1e-9);              // immediates / constants
                if (MODE == 1) c[i] = fma(c[i], x[i], 1e-9);                   // two registers
                if (MODE == 2) c[i] = fma(c[i], x[i], y[i]);                   // three registers, same chain
                if (MODE == 3) c[i] = fma(c[i], x[(i + u) & 7], y[(i + 3 * u) & 7]);   // three registers, rotating
                if (MODE == 4) c[i] = fma(c[(i + 1) & 7], x[(i + u) & 7], c[i]);       // products of other chains (ILP 8 kept)
                if (MODE == 5) c[i] = fma(c[(i + 3) & 7], x[i], c[i]);                 // another chain's value, fixed x
                if (MODE == 6) c[i] = fma(c[i], x[i], c[(i + 4) & 7]);                 // addend from another chain
                if (MODE == 7) c[i] = fma(c[(i + 5) & 7], c[(i + 2) & 7], c[i]);       // all three operands are chain values
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* label, int warps_per_sm) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int tpb = 128, blocks = sms * warps_per_sm / 4, iters = 20000;
    double *out, *in;
    cudaMalloc(&out, sizeof(double) * blocks * tpb);
    cudaMalloc(&in, sizeof(double) * 24);
    double h[24];
    for (int i = 0; i < 24; i++) h[i] = i < 8 ? 1.0 + i : (i < 16 ? 0.999999 - 1e-9 * i : 1e-9 * i);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, tpb>>>(out, in, 100);
    cudaEventRecord(e0);
    k<MODE><<<blocks, tpb>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 64.0 * iters * (double)blocks * tpb;
    printf("%-44s %2d warps/SM: %7.2f TFLOP/s (%.1f %% of %d SMs x 64 FMA x %.3f GHz)\n", label, warps_per_sm,
           fl / (ms * 1e-3) / 1e12, 100.0 * fl / (ms * 1e-3) / (sms * 128.0 * khz * 1e3), sms, khz * 1e-6);
    cudaFree(out); cudaFree(in);
}

int main() {
    for (int w : {8, 16}) {
        run<0>("fma(c, const, const)", w);
        run<1>("fma(c, x, const)", w);
        run<2>("fma(c, x, y)  same registers per chain", w);
        run<3>("fma(c, x', y') rotating registers", w);
        run<4>("fma(c', x', c) operands of other chains", w);
        run<5>("fma(c[i+3], x, c)", w);
        run<6>("fma(c, x, c[i+4])", w);
        run<7>("fma(c[i+5], c[i+2], c)", w);
    }
    return 0;
}
