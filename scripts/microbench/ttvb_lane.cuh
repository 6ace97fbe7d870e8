// ttvb_lane.cuh -- the Wisdom-Holman step with one LANE per planet.
//
// A planetary system of N planets occupies N adjacent lanes of a warp (group g = lanes
// g*N .. g*N+N-1, 32/N groups per warp). Each lane keeps the Jacobi state of ITS planet
// (6 doubles) in registers; the Kepler drift, which is most of the arithmetic, needs nothing
// else. The interaction kick needs the other planets: their positions and partial sums travel
// through warp shuffles. Every sum is accumulated in the same order as the per-system form in
// ttvb_math.cuh (accel<N>), so the bits are the same as the oracle's.
//
// Why: the per-system form keeps 6N state doubles plus N lockstep Newton chains in one thread
// and needs the whole 255-register budget (8 warps per SM, FP64 pipe 62 % busy, bound by
// dependent-issue latency). One planet per lane needs about 100 registers, so 16-20 warps are
// resident and the scheduler always finds an independent FP64 instruction.
#pragma once
#include "../../nauyaca_b200/csrc/ttvb_math.cuh"

namespace ttvb {

// Lanes beyond the last full group of a warp mirror group 0 (they compute on valid numbers,
// never on garbage, and never write results).
template <int N>
struct LaneMap {
    static constexpr int SPW = 32 / N;           // systems per warp
    int i;        // planet index of this lane
    int base;     // first lane of the group
    int g;        // group (system) index within the warp; >= SPW for mirror lanes
    bool real;
    __device__ __forceinline__ LaneMap() {
        const int lane = threadIdx.x & 31;
        g = lane / N;
        i = lane - g * N;
        real = g < SPW;
        base = real ? g * N : 0;
    }
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Per-lane constants of the step loop.
template <int N>
struct LaneConsts {
    double gm, kc, q, ieta, s;     // of this lane's planet
    double gmv[N];                 // G*m_j of every planet of the system
    double qv[N];                  // G*m_j/eta_j
    // pair schedule: for offset d = 1..N/2 this lane computes the pair with planet (i+d)%N;
    // for d = 1..(N-1)/2 it fetches the pair with planet (i-d)%N from that lane.
    double coef_own[N / 2 + 1];    // +G*m_j if i<j else -G*m_j, j = (i+d)%N  (index d)
    double coef_fet[N / 2 + 1];    // same for j = (i-d)%N
    __device__ __forceinline__ void finish(const LaneMap<N>& L) {
#pragma unroll
        for (int j = 0; j < N; j++) { gmv[j] = shfl_d(gm, L.base + j); qv[j] = shfl_d(q, L.base + j); }
#pragma unroll
        for (int d = 1; d <= N / 2; d++) {
            const int j = (L.i + d) % N;
            const double g = shfl_d(gm, L.base + j);
            coef_own[d] = (L.i < j) ? g : -g;
        }
#pragma unroll
        for (int d = 1; d <= (N - 1) / 2; d++) {
            const int j = (L.i - d + N) % N;
            const double g = shfl_d(gm, L.base + j);
            coef_fet[d] = (L.i < j) ? g : -g;
        }
    }
    // accessor concept of map_A_joint<1,...>
    __device__ __forceinline__ double kc_(int) const { return kc; }
};

template <int N>
struct LaneKc {
    double v;
    __device__ __forceinline__ double kc(int) const { return v; }
};

// Interaction acceleration of this lane's Jacobi velocity (same sums, same order as accel<N>);
// also |x_i| and 1/|x_i| for the drift that follows.
template <int N>
__device__ __forceinline__ void accel_lane(const LaneMap<N>& L, const LaneConsts<N>& C, const double (&x)[3],
                                           double (&a)[3], double& rj, double& irj) {
    const int i = L.i;
    // ---- astrocentric position r_i = x_i + sum_{j<i} q_j x_j
    double r[3] = {0.0, 0.0, 0.0}, R[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (i == j) {
#pragma unroll
            for (int k = 0; k < 3; k++) r[k] = x[k] + R[k];
        }
        if (j < N - 1) {
#pragma unroll
            for (int k = 0; k < 3; k++) R[k] = fma(C.qv[j], shfl_d(x[k], L.base + j), R[k]);
        }
    }
    double w[3], rr, irr;
    over_r3(r[0], r[1], r[2], w[0], w[1], w[2], rr, irr);
    // ---- direct terms D_i = sum_{j!=i} (+-)G m_j (r_hi - r_lo)/|r_hi - r_lo|^3, j ascending
    double D[3] = {0.0, 0.0, 0.0};
    if constexpr (N >= 2) {
        double own[N / 2 + 1][3], fet[N / 2 + 1][3];
#pragma unroll
        for (int d = 1; d <= N / 2; d++) {
            const int j = (i + d) % N;
            double rp[3];
#pragma unroll
            for (int k = 0; k < 3; k++) rp[k] = shfl_d(r[k], L.base + j);
            const bool lo = i < j;
            double df[3];
#pragma unroll
            for (int k = 0; k < 3; k++) df[k] = (lo ? rp[k] : r[k]) - (lo ? r[k] : rp[k]);
            over_r3(df[0], df[1], df[2], own[d][0], own[d][1], own[d][2]);
        }
#pragma unroll
        for (int d = 1; d <= (N - 1) / 2; d++) {
            const int j = (i - d + N) % N;
#pragma unroll
            for (int k = 0; k < 3; k++) fet[d][k] = shfl_d(own[d][k], L.base + j);
        }
        if constexpr (N == 2) {
#pragma unroll
            for (int k = 0; k < 3; k++) D[k] = fma(C.coef_own[1], own[1][k], D[k]);
        } else if constexpr (N == 3) {
            // partners: own[1] -> (i+1)%3, fet[1] -> (i-1)%3. Ascending j: lane 0: 1,2 (own,fet);
            // lane 1: 0,2 (fet,own); lane 2: 0,1 (own,fet).
            const bool sw = (i == 1);
            const double c0 = sw ? C.coef_fet[1] : C.coef_own[1], c1 = sw ? C.coef_own[1] : C.coef_fet[1];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double p0 = sw ? fet[1][k] : own[1][k], p1 = sw ? own[1][k] : fet[1][k];
                D[k] = fma(c0, p0, D[k]);
                D[k] = fma(c1, p1, D[k]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; j++) {
                double c = 0.0, pw[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int d = 1; d <= N / 2; d++) {
                    const bool m = ((i + d) % N) == j;
                    c = m ? C.coef_own[d] : c;
#pragma unroll
                    for (int k = 0; k < 3; k++) pw[k] = m ? own[d][k] : pw[k];
                }
#pragma unroll
                for (int d = 1; d <= (N - 1) / 2; d++) {
                    const bool m = ((i - d + N) % N) == j;
                    c = m ? C.coef_fet[d] : c;
#pragma unroll
                    for (int k = 0; k < 3; k++) pw[k] = m ? fet[d][k] : pw[k];
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double t = fma(c, pw[k], D[k]);
                    D[k] = (i == j) ? D[k] : t;
                }
            }
        }
    }
    // ---- T_i = sum_{k>i} G m_k w_k, accumulated from the outermost planet inwards
    double T[3] = {0.0, 0.0, 0.0}, acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k2 = N - 1; k2 >= 0; k2--) {
        if (i == k2) {
#pragma unroll
            for (int k = 0; k < 3; k++) T[k] = acc[k];
        }
        if (k2 > 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) acc[k] = fma(C.gmv[k2], shfl_d(w[k], L.base + k2), acc[k]);
        }
    }
    // ---- U_i = sum_{j<i} G m_j D_j
    double U[3] = {0.0, 0.0, 0.0}, us[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (i == j) {
#pragma unroll
            for (int k = 0; k < 3; k++) U[k] = us[k];
        }
        if (j < N - 1) {
#pragma unroll
            for (int k = 0; k < 3; k++) us[k] = fma(C.gmv[j], shfl_d(D[k], L.base + j), us[k]);
        }
    }
    // ---- Jacobi Kepler term and assembly
    double wj[3], rx, irx;
    over_r3(x[0], x[1], x[2], wj[0], wj[1], wj[2], rx, irx);
    const bool first = (i == 0);
    rj = first ? rr : rx;
    irj = first ? irr : irx;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double t0 = fma(C.kc, wj[k] - w[k], D[k]);
        const double t1 = fma(-C.s, T[k], first ? D[k] : t0);
        const double t2 = fma(-C.ieta, U[k], t1);
        a[k] = first ? t1 : t2;
    }
}

// One planet's drift with known |x| and 1/|x| (bits of map_A_joint<N,...,true> per planet).
__device__ __forceinline__ bool drift_lane(double kc, double tau, double r0, double ir0, double (&x)[3],
                                           double (&v)[3]) {
    State<1> p;
#pragma unroll
    for (int k = 0; k < 3; k++) { p.x[0][k] = x[k]; p.v[0][k] = v[k]; }
    LaneKc<1> C{kc};
    const bool ok = map_A_joint<1, LaneKc<1>, true>(C, p, tau, &r0, &ir0);
#pragma unroll
    for (int k = 0; k < 3; k++) { x[k] = p.x[0][k]; v[k] = p.v[0][k]; }
    return ok;
}

__device__ __forceinline__ bool drift_lane(double kc, double tau, double (&x)[3], double (&v)[3]) {
    State<1> p;
#pragma unroll
    for (int k = 0; k < 3; k++) { p.x[0][k] = x[k]; p.v[0][k] = v[k]; }
    LaneKc<1> C{kc};
    const bool ok = map_A_joint<1, LaneKc<1>, false>(C, p, tau, nullptr, nullptr);
#pragma unroll
    for (int k = 0; k < 3; k++) { x[k] = p.x[0][k]; v[k] = p.v[0][k]; }
    return ok;
}

// B(tau) for this lane's planet; returns the Jacobi radius for the drift.
template <int N>
__device__ __forceinline__ void kick_lane(const LaneMap<N>& L, const LaneConsts<N>& C, const double (&x)[3],
                                          double (&v)[3], double tau, double& rj, double& irj) {
    double a[3];
    accel_lane<N>(L, C, x, a, rj, irj);
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = fma(tau, a[k], v[k]);
}

}  // namespace ttvb
