// mufu_dump.cu -- dumps the FP64 seed functions of the sm_100a special-function unit.
//
// `rcp.approx.ftz.f64` (MUFU.RCP64H) and `rsqrt.approx.ftz.f64` (MUFU.RSQ64H) read only the
// HIGH 32-bit word of their operand and produce a high word (low word zero). The canonical
// arithmetic of libttvb ("v3", nauyaca_b200/csrc/ttvb_math.cuh) starts its reciprocals and
// reciprocal square roots from these seeds WITHOUT a correctly-rounding final step, so the CPU
// oracle has to reproduce the seeds bit for bit. This program writes them out exhaustively:
//   rcp: operands 1.m  (hi = 0x3FF00000 | m, m = 0..2^20-1)            -> 2^20 high words
//   rsq: operands 1.m and 2*(1.m) (hi = 0x3FF00000 | m, 0x40000000 | m) -> 2^21 high words
// and checks (a) that the low operand word is ignored and the low result word is zero,
// (b) that other exponents only rescale the result by an exact power of two (the rule the
// oracle uses: rcp(2^E x) = 2^-E rcp(x), rsq(2^(2k) x) = 2^-k rsq(x)), over 2^26 random operands
// with exponents in [-300, 300].
//
// Build/run (GPU box):  nvcc -arch=sm_100a -O2 -o mufu_dump scripts/mufu_dump.cu && ./mufu_dump out_dir
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ double rcp_seed(double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    return y;
}
__device__ __forceinline__ double rsq_seed(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}

__global__ void dump(uint32_t* rcp, uint32_t* rsq, unsigned long long* bad) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 20)) return;
    const uint32_t los[3] = {0u, 0xFFFFFFFFu, 0x9E3779B9u * (m + 1u)};
    for (int e = 0; e < 2; e++) {
        const uint32_t hi = (e == 0 ? 0x3FF00000u : 0x40000000u) | m;
        uint32_t r0 = 0, s0 = 0;
        for (int l = 0; l < 3; l++) {
            const double x = __hiloint2double((int)hi, (int)los[l]);
            const double r = rcp_seed(x), s = rsq_seed(x);
            if (__double2loint(r) != 0 || __double2loint(s) != 0) atomicAdd(bad, 1ull);
            if (l == 0) { r0 = (uint32_t)__double2hiint(r); s0 = (uint32_t)__double2hiint(s); }
            else if ((uint32_t)__double2hiint(r) != r0 || (uint32_t)__double2hiint(s) != s0) atomicAdd(bad + 1, 1ull);
        }
        if (e == 0) rcp[m] = r0;
        rsq[(size_t)e * (1u << 20) + m] = s0;
    }
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// exponent rule: for E in [-300,300]: rcp(2^E * 1.m) == 2^-E * rcp(1.m); rsq(2^(2k+p) * 1.m) == 2^-k rsq(2^p 1.m)
__global__ void check_scale(const uint32_t* rcp, const uint32_t* rsq, unsigned long long* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t h = mix(i * 2654435761u + 12345u);
    const uint32_t m = h & 0xFFFFFu;
    const int E = (int)(mix(h) % 601u) - 300;
    const uint32_t hi = ((uint32_t)(1023 + E) << 20) | m;
    const double x = __hiloint2double((int)hi, (int)mix(h + 7u));
    // rcp
    {
        const uint32_t t = rcp[m];
        const int te = (int)(t >> 20) - 1023;            // exponent of rcp(1.m): 0 or -1
        const uint32_t want = ((uint32_t)(1023 + te - E) << 20) | (t & 0xFFFFFu);
        const double r = rcp_seed(x);
        if ((uint32_t)__double2hiint(r) != want || __double2loint(r) != 0) atomicAdd(bad + 2, 1ull);
    }
    // rsq
    {
        const int p = E & 1;                             // E = 2k + p, p in {0,1} (also for negative E)
        const int k = (E - p) / 2;
        const uint32_t t = rsq[(size_t)p * (1u << 20) + m];
        const int te = (int)(t >> 20) - 1023;
        const uint32_t want = ((uint32_t)(1023 + te - k) << 20) | (t & 0xFFFFFu);
        const double s = rsq_seed(x);
        if ((uint32_t)__double2hiint(s) != want || __double2loint(s) != 0) atomicAdd(bad + 3, 1ull);
    }
}

int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : ".";
    uint32_t *d_rcp, *d_rsq;
    unsigned long long* d_bad;
    cudaMalloc(&d_rcp, 4u << 20);
    cudaMalloc(&d_rsq, 8u << 20);
    cudaMalloc(&d_bad, 4 * 8);
    cudaMemset(d_bad, 0, 4 * 8);
    dump<<<(1 << 20) / 256, 256>>>(d_rcp, d_rsq, d_bad);
    check_scale<<<(1 << 26) / 256, 256>>>(d_rcp, d_rsq, d_bad);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
    uint32_t* rcp = (uint32_t*)malloc(4u << 20);
    uint32_t* rsq = (uint32_t*)malloc(8u << 20);
    unsigned long long bad[4];
    cudaMemcpy(rcp, d_rcp, 4u << 20, cudaMemcpyDeviceToHost);
    cudaMemcpy(rsq, d_rsq, 8u << 20, cudaMemcpyDeviceToHost);
    cudaMemcpy(bad, d_bad, 32, cudaMemcpyDeviceToHost);
    printf("low result word nonzero: %llu, low operand word matters: %llu, rcp exponent rule broken: %llu, rsq exponent rule broken: %llu\n",
           bad[0], bad[1], bad[2], bad[3]);
    char path[1024];
    snprintf(path, sizeof(path), "%s/mufu_rcp64h.u32", dir);
    FILE* f = fopen(path, "wb"); fwrite(rcp, 4, 1u << 20, f); fclose(f);
    snprintf(path, sizeof(path), "%s/mufu_rsq64h.u32", dir);
    f = fopen(path, "wb"); fwrite(rsq, 4, 2u << 20, f); fclose(f);
    // worst relative error of the seeds over the table (informative)
    double wr = 0, ws = 0;
    for (uint32_t m = 0; m < (1u << 20); m++) {
        union { double d; uint64_t u; } x, y;
        x.u = ((uint64_t)(0x3FF00000u | m)) << 32;
        y.u = ((uint64_t)rcp[m]) << 32;
        double e = y.d * x.d - 1.0; if (e < 0) e = -e; if (e > wr) wr = e;
        for (int p = 0; p < 2; p++) {
            x.u = ((uint64_t)((p ? 0x40000000u : 0x3FF00000u) | m)) << 32;
            y.u = ((uint64_t)rsq[(size_t)p * (1u << 20) + m]) << 32;
            e = y.d * y.d * x.d - 1.0; if (e < 0) e = -e; if (e > ws) ws = e;
        }
    }
    printf("max |x*rcp-1| = %.3e (lo word 0), max |x*rsq^2-1| = %.3e\n", wr, ws);
    return (bad[0] | bad[1] | bad[2] | bad[3]) ? 2 : 0;
}
