// TEST-ONLY host build of the kernel's per-system arithmetic (nauyaca_b200/csrc/ttvb_math.cuh).
//
// Purpose: let the CPU-only test suite (`-m "not gpu"`) check that the arithmetic the CUDA
// kernel executes is bit-identical to the independent oracle (oracle/ttv_oracle.c) without a
// GPU. This file is NOT part of libttvb.so, is not importable from the nauyaca_b200 package
// and is not a fallback: the product has no CPU path. It mirrors the control flow of
// ttvb_main_kernel for ONE system (prologue, corrector loop, two-deep state ring, trigger,
// refinement of node k then k-1 / k+1), minus the warp queue.
//
// Build: g++ -O2 -ffp-contract=off -mfma -fPIC -shared -o libhh.so harness.cpp
#include "../../nauyaca_b200/csrc/ttvb_math.cuh"

using namespace ttvb;

template <int N>
static int run(const double* flat, double mstar, double t0, double dt, double total, int cap, int* n_events,
               int* planet, int* epoch, double* time, double* rsky, double* vsky) {
    int status = 0;
    *n_events = 0;
    long long nsteps = 0;
    { double T = t0; while (T < total) { T += dt; nsteps++; } }
    Consts<N> C;
    State<N> p;
    C.gm0 = TTVB_G * mstar;
    double eta_prev = C.gm0, R[3] = {0, 0, 0}, V[3] = {0, 0, 0};
    bool valid = true;
    for (int i = 0; i < N; i++) {
        double e[7];
        for (int c = 0; c < 7; c++) e[c] = flat[7 * i + c];
        PlanetInit pi;
        valid = planet_init(e, C.gm0, eta_prev, pi) && valid;
        C.gm_[i] = pi.gm; C.kc_[i] = pi.kc; C.q_[i] = pi.q; C.ieta_[i] = pi.ieta; C.s_[i] = pi.s;
        for (int k = 0; k < 3; k++) {
            p.x[i][k] = pi.x[k] - R[k];
            p.v[i][k] = pi.v[k] - V[k];
            R[k] = fma(pi.q, p.x[i][k], R[k]);
            V[k] = fma(pi.q, p.v[i][k], V[k]);
        }
    }
    if (!valid) return 16;
    bool ok = corrector<N>(C, p, dt);
    double prev_dot[N];
    int count[N], trig[N];
    for (int i = 0; i < N; i++) { prev_dot[i] = sky_g(p.x[i][0], p.x[i][1], p.v[i][0], p.v[i][1]); count[i] = 0; trig[i] = 0; }
    if (ok) ok = map_A<N>(C, p, 0.5 * dt);
    if (!ok) return 2;
    State<N> ring[2];
    double Time = t0, Tm1 = t0, Tm2 = t0;
    for (long long k = 1; k <= nsteps + 1; k++) {
        ring[(k - 1) & 1] = p;
        double rj[N], irj[N];
        map_B<N>(C, p, dt, rj, irj);
        if (!map_A_after_B<N>(C, p, dt, rj, irj)) { status |= 2; break; }
        Tm2 = Tm1; Tm1 = Time; Time += dt;
        for (int i = 0; i < N; i++) {
            if (!trig[i]) continue;
            trig[i] = 0;
            const State<N>* st3[3] = {&ring[k & 1], &ring[(k - 1) & 1], &p};
            const double T3[3] = {Tm2, Tm1, Time};
            double xs[2][3], vs2[2][3], gs[2];
            int which[2] = {1, 0};
            bool okk = true;
            for (int pass = 0; pass < 2; pass++) {
                const State<N>& s = *st3[which[pass]];
                double R0 = 0, R1 = 0, R2 = 0, V0 = 0, V1 = 0, V2 = 0;
                for (int j = 0; j <= i; j++) {
                    double x0 = s.x[j][0], x1 = s.x[j][1], x2 = s.x[j][2], v0 = s.v[j][0], v1 = s.v[j][1], v2 = s.v[j][2];
                    okk = kepler_drift(C.kc(j), -(0.5 * dt), x0, x1, x2, v0, v1, v2) && okk;
                    if (j < i) {
                        R0 = fma(C.q(j), x0, R0); R1 = fma(C.q(j), x1, R1); R2 = fma(C.q(j), x2, R2);
                        V0 = fma(C.q(j), v0, V0); V1 = fma(C.q(j), v1, V1); V2 = fma(C.q(j), v2, V2);
                    } else {
                        xs[pass][0] = x0 + R0; xs[pass][1] = x1 + R1; xs[pass][2] = x2 + R2;
                        vs2[pass][0] = v0 + V0; vs2[pass][1] = v1 + V1; vs2[pass][2] = v2 + V2;
                    }
                }
                gs[pass] = sky_g(xs[pass][0], xs[pass][1], vs2[pass][0], vs2[pass][1]);
                if (pass == 0) which[1] = (gs[0] > 0.0) ? 0 : 2;
            }
            double tt, rs, vv;
            if (okk) {
                const double mu = C.gm0 + C.gm(i);
                if (which[1] == 0) okk = locate_bracket(mu, dt, T3[0], T3[1], xs[1], vs2[1], gs[1], xs[0], vs2[0], gs[0], tt, rs, vv);
                else okk = locate_bracket(mu, dt, T3[1], T3[2], xs[0], vs2[0], gs[0], xs[1], vs2[1], gs[1], tt, rs, vv);
            }
            if (!okk) { tt = rs = vv = TTVB_BAD_TRANSIT; status |= 4; }
            if (*n_events < cap) {
                int e = *n_events;
                planet[e] = i; epoch[e] = count[i]; time[e] = tt; rsky[e] = rs; vsky[e] = vv;
                *n_events = e + 1;
            } else status |= 8;
            count[i]++;
        }
        if (k <= nsteps)
            for (int i = 0; i < N; i++) {
                const double cd = sky_g(p.x[i][0], p.x[i][1], p.v[i][0], p.v[i][1]);
                if (prev_dot[i] < 0.0 && cd > 0.0 && p.x[i][2] > 0.0) trig[i] = 1;
                prev_dot[i] = cd;
            }
    }
    return status;
}

// the seed tables of oracle/mufu_sm100a.npz (see ttvb_math.cuh, "canonical arithmetic v3")
extern "C" void hh_set_seed_tables(const uint32_t* rcp, const uint32_t* rsq) {
    host_seed_tables().rcp = rcp;
    host_seed_tables().rsq = rsq;
}

extern "C" int hh_ephemeris(const double* flat, int npla, double mstar, double t0, double dt, double total, int cap,
                            int* n_events, int* planet, int* epoch, double* time, double* rsky, double* vsky) {
    switch (npla) {
        case 1: return run<1>(flat, mstar, t0, dt, total, cap, n_events, planet, epoch, time, rsky, vsky);
        case 2: return run<2>(flat, mstar, t0, dt, total, cap, n_events, planet, epoch, time, rsky, vsky);
        case 3: return run<3>(flat, mstar, t0, dt, total, cap, n_events, planet, epoch, time, rsky, vsky);
        case 4: return run<4>(flat, mstar, t0, dt, total, cap, n_events, planet, epoch, time, rsky, vsky);
    }
    return -1;
}
