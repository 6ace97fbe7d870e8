// Development microbenchmark: does a warp with only 16 (or 8) active lanes issue FP64
// instructions faster than a full warp? (The FP64 pipe of a scheduler is 16 lanes wide, a full
// warp instruction takes two passes.) One warp per block, ILP independent FMA chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/halfwarp_fp64 scripts/microbench/halfwarp_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chains(double* out, long long* cyc, int iters, int active, double a, double b) {
    if ((int)threadIdx.x >= active) return;
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
    out[blockIdx.x * 32 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int active, int lane_stride_note) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 32 * 4); cudaMalloc(&cyc, 8 * 4);
    const int iters = 4000;
    for (int rep = 0; rep < 2; rep++) chains<ILP><<<1, 32>>>(out, cyc, iters, active, 0.999999, 1e-9);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP=%d active lanes=%2d : %.2f cycles per warp-level DFMA\n", ILP, active, (double)h / (iters * 8.0 * ILP));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int act : {32, 16, 8}) { run<1>(act, 0); run<4>(act, 0); run<8>(act, 0); }
    return 0;
}
