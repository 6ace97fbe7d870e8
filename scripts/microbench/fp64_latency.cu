// Development microbenchmark: dependent-issue latency and throughput of the FP64 pipe on the
// device at hand (single warp / many warps, dependent chain / independent chains).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void mufu_chain(double* out, long long* cyc, int iters) {
    double c = 1.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c));
            c = y + 2.0;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps_per_block, int blocks, const char* label) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * blocks * warps_per_block * 32);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    const int iters = 2000;
    chain<ILP><<<blocks, warps_per_block * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    chain<ILP><<<blocks, warps_per_block * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    long long h;
    cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 16.0 * ILP);
    printf("%-34s ILP=%d warps/block=%2d : %.2f cycles per DFMA per warp-chain, %.2f cycles per DFMA per SM-quadrant-issue\n",
           label, ILP, warps_per_block, per * ILP, per * ILP / ILP / (warps_per_block >= 4 ? warps_per_block / 4.0 : 1.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1>(1, 1, "1 warp, dependent chain");
    run<2>(1, 1, "1 warp, 2 independent chains");
    run<4>(1, 1, "1 warp, 4 independent chains");
    run<8>(1, 1, "1 warp, 8 independent chains");
    run<1>(4, 1, "4 warps (1/SMSP), dependent");
    run<1>(8, 1, "8 warps (2/SMSP), dependent");
    run<1>(16, 1, "16 warps (4/SMSP), dependent");
    run<1>(32, 1, "32 warps (8/SMSP), dependent");
    run<2>(8, 1, "8 warps (2/SMSP), ILP 2");
    run<4>(8, 1, "8 warps (2/SMSP), ILP 4");
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 32); cudaMalloc(&cyc, 8);
    mufu_chain<<<1, 32>>>(out, cyc, 1000);
    mufu_chain<<<1, 32>>>(out, cyc, 1000);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RSQ64H + DADD dependent pair: %.2f cycles\n", (double)h / 16000.0);
    return 0;
}
