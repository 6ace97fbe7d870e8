// Development microbenchmark: the bare Wisdom-Holman step loop (kick + drift, no transit
// search) with one thread per system (ttvb_math.cuh map_B / map_A_after_B) against one lane
// per planet (ttvb_lane.cuh). Checks that both end in the same bits and times them at full
// occupancy and for a lone warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -lineinfo \
//        -o scripts/microbench/lane_step scripts/microbench/lane_step.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "ttvb_lane.cuh"

using namespace ttvb;

#ifndef LANE_TPB
#define LANE_TPB 128
#endif
#ifndef LANE_MINB
#define LANE_MINB 4
#endif

template <int N>
__global__ void init_kernel(const double* el, double gm0, double* st, double* cst, int nsys) {
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= nsys) return;
    double eta_prev = gm0, R[3] = {0, 0, 0}, V[3] = {0, 0, 0};
    for (int i = 0; i < N; i++) {
        double e[7];
        for (int c = 0; c < 7; c++) e[c] = el[((size_t)sys * N + i) * 7 + c];
        PlanetInit pi;
        planet_init(e, gm0, eta_prev, pi);
        double* c5 = cst + ((size_t)sys * N + i) * 5;
        c5[0] = pi.gm; c5[1] = pi.kc; c5[2] = pi.q; c5[3] = pi.ieta; c5[4] = pi.s;
        double* s6 = st + ((size_t)sys * N + i) * 6;
        for (int k = 0; k < 3; k++) {
            double px = pi.x[k] - R[k], pv = pi.v[k] - V[k];
            s6[k] = px; s6[3 + k] = pv;
            R[k] = fma(pi.q, px, R[k]);
            V[k] = fma(pi.q, pv, V[k]);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(128, 1)
sys_steps(const double* st, const double* cst, double* out, int nsys, int nsteps, double dt) {
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= nsys) return;
    Consts<N> C;
    State<N> p;
    for (int i = 0; i < N; i++) {
        const double* c5 = cst + ((size_t)sys * N + i) * 5;
        C.gm_[i] = c5[0]; C.kc_[i] = c5[1]; C.q_[i] = c5[2]; C.ieta_[i] = c5[3]; C.s_[i] = c5[4];
        const double* s6 = st + ((size_t)sys * N + i) * 6;
        for (int k = 0; k < 3; k++) { p.x[i][k] = s6[k]; p.v[i][k] = s6[3 + k]; }
    }
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < nsteps; k++) {
        double rj[N], irj[N];
        map_B<N>(C, p, dt, rj, irj);
        ok = map_A_after_B<N>(C, p, dt, rj, irj) && ok;
    }
    for (int i = 0; i < N; i++) {
        double* o6 = out + ((size_t)sys * N + i) * 6;
        for (int k = 0; k < 3; k++) { o6[k] = p.x[i][k]; o6[3 + k] = ok ? p.v[i][k] : -1.0; }
    }
}

template <int N>
__global__ void __launch_bounds__(LANE_TPB, LANE_MINB)
lane_steps(const double* st, const double* cst, double* out, int nsys, int nsteps, double dt) {
    const LaneMap<N> L;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int g = L.real ? L.g : 0;
    int sys = warp_global * LaneMap<N>::SPW + g;
    const bool live = sys < nsys;
    if (!live) sys = nsys - 1;
    LaneConsts<N> C;
    double x[3], v[3];
    {
        const double* c5 = cst + ((size_t)sys * N + L.i) * 5;
        C.gm = c5[0]; C.kc = c5[1]; C.q = c5[2]; C.ieta = c5[3]; C.s = c5[4];
        const double* s6 = st + ((size_t)sys * N + L.i) * 6;
        for (int k = 0; k < 3; k++) { x[k] = s6[k]; v[k] = s6[3 + k]; }
    }
    C.finish(L);
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < nsteps; k++) {
        double rj, irj;
        kick_lane<N>(L, C, x, v, dt, rj, irj);
        ok = drift_lane(C.kc, dt, rj, irj, x, v) && ok;
    }
    if (L.real && live) {
        double* o6 = out + ((size_t)sys * N + L.i) * 6;
        for (int k = 0; k < 3; k++) { o6[k] = x[k]; o6[3 + k] = ok ? v[k] : -1.0; }
    }
}

static double urand() { return rand() / (RAND_MAX + 1.0); }

template <int N>
void run(int nsys, int nsteps) {
    const double periods[6] = {10.47, 14.11, 23.57, 37.9, 61.3, 99.1};
    std::vector<double> el((size_t)nsys * N * 7);
    for (int s = 0; s < nsys; s++)
        for (int i = 0; i < N; i++) {
            double* e = &el[((size_t)s * N + i) * 7];
            e[0] = 3.0 + 30.0 * urand(); e[1] = periods[i] * (1.0 + 0.01 * (urand() - 0.5));
            e[2] = 0.005 + 0.1 * urand(); e[3] = 88.5 + 2.0 * urand(); e[4] = 360.0 * urand();
            e[5] = 360.0 * urand(); e[6] = 180.0 + 2.0 * (urand() - 0.5);
        }
    double *d_el, *d_st, *d_cst, *d_o1, *d_o2;
    cudaMalloc(&d_el, el.size() * 8);
    cudaMalloc(&d_st, (size_t)nsys * N * 6 * 8);
    cudaMalloc(&d_cst, (size_t)nsys * N * 5 * 8);
    cudaMalloc(&d_o1, (size_t)nsys * N * 6 * 8);
    cudaMalloc(&d_o2, (size_t)nsys * N * 6 * 8);
    cudaMemcpy(d_el, el.data(), el.size() * 8, cudaMemcpyHostToDevice);
    const double gm0 = TTVB_G * 0.522, dt = 0.1;
    init_kernel<N><<<(nsys + 127) / 128, 128>>>(d_el, gm0, d_st, d_cst, nsys);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms_sys = 0, ms_lane = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        sys_steps<N><<<(nsys + 127) / 128, 128>>>(d_st, d_cst, d_o1, nsys, nsteps, dt);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms_sys, a, b);
        const int spw = 32 / N, wpb = LANE_TPB / 32;
        const int nwarps = (nsys + spw - 1) / spw, nblk = (nwarps + wpb - 1) / wpb;
        cudaEventRecord(a);
        lane_steps<N><<<nblk, LANE_TPB>>>(d_st, d_cst, d_o2, nsys, nsteps, dt);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms_lane, a, b);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<double> o1((size_t)nsys * N * 6), o2(o1.size());
    cudaMemcpy(o1.data(), d_o1, o1.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2.data(), d_o2, o2.size() * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < o1.size(); i++) if (memcmp(&o1[i], &o2[i], 8) != 0) bad++;
    printf("N=%d nsys=%7d nsteps=%d : per-system %.3f ms (%.1f ps/sys-step), per-lane %.3f ms (%.1f ps/sys-step), "
           "ratio %.3f, mismatching doubles %zu / %zu  sample %.17g\n",
           N, nsys, nsteps, ms_sys, ms_sys * 1e9 / ((double)nsys * nsteps), ms_lane,
           ms_lane * 1e9 / ((double)nsys * nsteps), ms_sys / ms_lane, bad, o1.size(), o1[0]);
    cudaFree(d_el); cudaFree(d_st); cudaFree(d_cst); cudaFree(d_o1); cudaFree(d_o2);
}

int main(int argc, char** argv) {
    const int nsteps = argc > 1 ? atoi(argv[1]) : 2000;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lane_steps<3>, LANE_TPB, 0);
    printf("lane_steps<3>: %d CTAs of %d threads per SM\n", nb, LANE_TPB);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lane_steps<2>, LANE_TPB, 0);
    printf("lane_steps<2>: %d CTAs of %d threads per SM\n", nb, LANE_TPB);
    // lone warp (latency), one wave, several waves
    run<3>(10, nsteps); run<3>(32, nsteps); run<3>(148 * 256, nsteps); run<3>(148 * 256 * 4, nsteps);
    run<2>(16, nsteps); run<2>(32, nsteps); run<2>(10000, nsteps); run<2>(148 * 256 * 4, nsteps);
    run<4>(148 * 256 * 2, nsteps);
    return 0;
}
