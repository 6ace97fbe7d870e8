// ttvb_kernels.cuh -- the sm_100a kernels of libttvb.so.
//
// Mapping: one THREAD per planetary system (parameter vector); its Jacobi state lives in
// registers for the whole integration. A CTA is persistent and walks over tiles of TPB
// systems. Per CTA, shared memory holds
//   - the observed transit table (epoch, time, sigma) against which chi2 is reduced,
//   - a two-deep ring of every thread's previous Jacobi states (needed only when a
//     transit triggers),
//   - the Kepler constants of every thread (read by other lanes, see below),
//   - per-thread chi2 accumulators and observed-table cursors.
// Transit refinement is rare (a few % of steps) and expensive (about three steps' worth
// of arithmetic). Doing it inline would make every warp pay for it on almost every step
// with one or two active lanes, so each warp instead owns an EVENT QUEUE (global memory,
// L2 resident): a lane that triggers pushes its three bracketing states; when 32 events
// are queued the whole warp refines them together, one event per lane, and the results
// are handed back to the owning lanes in queue order, which keeps the chi2 summation
// order of the reference (utils.py:161-216).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ttvb_math.cuh"

#define TTVB_MAXFLAT 63   // 7 * TTVB_MAX_PLANETS
#ifndef TTVB_TPB
#define TTVB_TPB 128
#endif
#ifndef TTVB_MINB
#define TTVB_MINB 1
#endif
#define TTVB_MAX_PLANETS_P1 10
#define TTVB_ST_BAD_TRANSIT_BIT 4

namespace ttvb {

struct DevSys {
    int npla, ndim, nconst, nobs;
    int rowlen;              // doubles per input row for the current mode
    int mode;
    long long B;
    long long nsteps;
    double gm0;              // G*Mstar
    double t0, dt, rstar_au, second_term;
    // flat slot k (0..7*npla-1) takes x[src[k]] if src[k] >= 0, else cval[-1-src[k]]
    int src[TTVB_MAXFLAT];
    double cval[TTVB_MAXFLAT];
    double bi[TTVB_MAXFLAT], dbf[TTVB_MAXFLAT];   // cube -> physical: bi + x*dbf
    double lo[TTVB_MAXFLAT], hi[TTVB_MAXFLAT];    // bounds of the current mode
    int obs_start[TTVB_MAX_PLANETS_P1];
    // device arrays
    const int* obs_epoch;
    const double* obs_time;
    const double* obs_sigma;
    const double* x;
    double* chi2_out;
    double* logl_out;
    double* chi2_planet_out;
    int* status_out;
    // ephemeris table outputs (nullable)
    int cap;
    int* n_events;
    int* ev_planet;
    int* ev_epoch;
    double* ev_time;
    double* ev_rsky;
    double* ev_vsky;
    // per-warp event queues
    unsigned char* queue;
    long long queue_stride;   // bytes per warp
    // work scheduling: units = (time segment, tile), claimed through sched[0]; sched[1+tile] is
    // the number of finished segments of a tile. With nseg > 1 a tile's integration is cut into
    // nseg pieces that different CTAs may run, the per-system context travelling through ctx.
    int nseg;
    long long seg_len;        // steps per segment
    int* sched;
    double* ctx_d;            // [25N+3][ctx_stride]
    int* ctx_i;               // [2N+4][ctx_stride]
    long long ctx_stride;
};

// ---- queue layout of one warp: 4x32 ints, then (18N+3)x32 doubles
template <int N>
struct QLayout {
    static constexpr int NI = 4;                    // owner, planet, epoch, seq
    static constexpr int ND = 18 * N + 3;           // 3 states x N planets x 6, then T3
    static constexpr long long BYTES = 32LL * (NI * 4 + ND * 8);
};

// per-thread Kepler constants in shared memory: field f of planet i at [(f*N+i)*TPB + tid]
template <int N>
struct SmemConsts {
    double* b;   // base + tid
    __device__ __forceinline__ double gm(int i) const { return b[(0 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double kc(int i) const { return b[(1 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double q(int i) const { return b[(2 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double ieta(int i) const { return b[(3 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double s(int i) const { return b[(4 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ void set(int i, double gm, double kc, double q, double ieta, double s) {
        b[(0 * N + i) * TTVB_TPB] = gm; b[(1 * N + i) * TTVB_TPB] = kc; b[(2 * N + i) * TTVB_TPB] = q;
        b[(3 * N + i) * TTVB_TPB] = ieta; b[(4 * N + i) * TTVB_TPB] = s;
    }
};

struct WarpQ {
    int* qi;
    double* qd;
};

// shared-memory carve-up (dynamic); the input staging area aliases the history ring
__host__ __device__ inline size_t smem_bytes(int N, int nobs, int rowlen) {
    size_t b = 0;
    b += (size_t)nobs * (8 + 8 + 4);
    b = (b + 7) & ~(size_t)7;
    size_t hist = (size_t)2 * 6 * N * TTVB_TPB * 8, stage = (size_t)TTVB_TPB * rowlen * 8;
    b += hist > stage ? hist : stage;                       // history ring / input staging
    b += (size_t)5 * N * TTVB_TPB * 8;                      // gm, kc, q, ieta, s
    b += (size_t)N * TTVB_TPB * 8;                          // chi2 accumulators
    b += (size_t)(TTVB_TPB / 32) * 32 * 2 * 8;              // result area (doubles)
    b += (size_t)N * TTVB_TPB * 4;                          // observed-table cursors
    b += (size_t)(TTVB_TPB / 32) * 32 * 4 * 4;              // result area (ints)
    return b;
}

// Refine up to 32 queued events, one per lane, and hand the results to their owners.
template <int N>
__device__ __noinline__ void process_batch(const DevSys& S, int count, WarpQ Q, const double* kc_s,
                                           const double* q_s, const double* gm_s, int* res_i,
                                           double* res_d, int warp_base_tid, long long warp_base_sys,
                                           const int* obs_epoch_s, const double* obs_time_s,
                                           const double* obs_sigma_s) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const bool act = lane < count;
    int owner = 0, planet = 0, epoch = 0, seq = 0;
    double tt = TTVB_BAD_TRANSIT, rs = TTVB_BAD_TRANSIT, vs = TTVB_BAD_TRANSIT;
    int flags = 0;
    if (act) {
        owner = __ldcg(Q.qi + 0 * 32 + lane);
        planet = __ldcg(Q.qi + 1 * 32 + lane);
        epoch = __ldcg(Q.qi + 2 * 32 + lane);
        seq = __ldcg(Q.qi + 3 * 32 + lane);
        const int otid = warp_base_tid + owner;
        const double hdt = 0.5 * S.dt;
        double T3[3];
        T3[0] = __ldcg(Q.qd + (18 * N + 0) * 32 + lane);
        T3[1] = __ldcg(Q.qd + (18 * N + 1) * 32 + lane);
        T3[2] = __ldcg(Q.qd + (18 * N + 2) * 32 + lane);
        double xs[2][3], vsn[2][3], gs[2];
        int which[2];
        which[0] = 1;
        bool ok = true;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int st = which[pass];
            const double* base = Q.qd + (long long)st * 6 * N * 32 + lane;
            double R0 = 0.0, R1 = 0.0, R2 = 0.0, V0 = 0.0, V1 = 0.0, V2 = 0.0;
#pragma unroll 1
            for (int j = 0; j <= planet; j++) {
                double x0 = __ldcg(base + (j * 6 + 0) * 32), x1 = __ldcg(base + (j * 6 + 1) * 32),
                       x2 = __ldcg(base + (j * 6 + 2) * 32), v0 = __ldcg(base + (j * 6 + 3) * 32),
                       v1 = __ldcg(base + (j * 6 + 4) * 32), v2 = __ldcg(base + (j * 6 + 5) * 32);
                ok = kepler_drift(kc_s[j * TTVB_TPB + otid], -hdt, x0, x1, x2, v0, v1, v2) && ok;
                if (j < planet) {
                    const double qj = q_s[j * TTVB_TPB + otid];
                    R0 = fma(qj, x0, R0); R1 = fma(qj, x1, R1); R2 = fma(qj, x2, R2);
                    V0 = fma(qj, v0, V0); V1 = fma(qj, v1, V1); V2 = fma(qj, v2, V2);
                } else {
                    xs[pass][0] = x0 + R0; xs[pass][1] = x1 + R1; xs[pass][2] = x2 + R2;
                    vsn[pass][0] = v0 + V0; vsn[pass][1] = v1 + V1; vsn[pass][2] = v2 + V2;
                }
            }
            gs[pass] = sky_g(xs[pass][0], xs[pass][1], vsn[pass][0], vsn[pass][1]);
            if (pass == 0) which[1] = (gs[0] > 0.0) ? 0 : 2;
        }
        if (ok) {
            const double mu = S.gm0 + gm_s[planet * TTVB_TPB + otid];
            // pass 0 is node k (mid); pass 1 is node k-1 (which[1]==0) or node k+1
            if (which[1] == 0)
                ok = locate_bracket(mu, S.dt, T3[0], T3[1], xs[1], vsn[1], gs[1], xs[0], vsn[0], gs[0], tt, rs, vs);
            else
                ok = locate_bracket(mu, S.dt, T3[1], T3[2], xs[0], vsn[0], gs[0], xs[1], vsn[1], gs[1], tt, rs, vs);
        }
        if (!ok) {
            tt = TTVB_BAD_TRANSIT; rs = TTVB_BAD_TRANSIT; vs = TTVB_BAD_TRANSIT;
            flags = TTVB_ST_BAD_TRANSIT_BIT;
        }
        if (S.ev_time != nullptr) {
            const long long sys = warp_base_sys + owner;
            if (seq < S.cap) {
                const long long o = sys * (long long)S.cap + seq;
                S.ev_planet[o] = planet; S.ev_epoch[o] = epoch;
                S.ev_time[o] = tt; S.ev_rsky[o] = rs; S.ev_vsky[o] = vs;
            }
        }
    }
    // ---- this event's chi2 term, computed here by all 32 lanes at once: events with
    //      rsky > Rstar are not transits (utils.py:113); pos = first observed epoch of the planet
    //      that is >= the event's epoch (observed epochs before it that the owner has not met
    //      are missing from the simulation, utils.py:199-206)
    int info = planet, pos = 0;
    double term = 0.0;
    if (act && rs <= S.rstar_au) {
        info |= 0x100;
        int a = S.obs_start[planet];
        const int end = S.obs_start[planet + 1];
        int b = end;
        while (a < b) {
            const int m = (a + b) >> 1;
            if (obs_epoch_s[m] < epoch) a = m + 1; else b = m;
        }
        pos = a;
        if (pos < end && obs_epoch_s[pos] == epoch) {
            info |= 0x200;
            const double d = (obs_time_s[pos] - tt) / obs_sigma_s[pos];
            term = d * d;
        }
    }
    res_i[0 * 32 + lane] = info;
    res_i[1 * 32 + lane] = pos;
    res_i[2 * 32 + lane] = flags;
    res_i[3 * 32 + lane] = 0;
    res_d[lane] = term;
    __syncwarp();
    // ---- per owner lane, the set of events it owns (queue order = bit order)
    const unsigned peers = __match_any_sync(0xffffffffu, act ? owner : 32 + lane);
    if (act && (peers & ((1u << lane) - 1u)) == 0u) res_i[3 * 32 + owner] = (int)peers;
    __syncwarp();
}

// Every lane folds the events it owns into its chi2, in queue order (which keeps the summation
// order of the reference, utils.py:161-216): observed epochs skipped by the simulation cost
// 1e20 each (utils.py:199-206), then the event's own term.
__device__ __forceinline__ void consume_results(int lane, int tid, const int* res_i, const double* res_d,
                                                double* acc_s, int* cur_s, int& status) {
    unsigned mine = (unsigned)res_i[3 * 32 + lane];
    while (__any_sync(0xffffffffu, mine != 0u)) {
        if (mine != 0u) {
            const int j = __ffs(mine) - 1;
            mine &= mine - 1u;
            const int info = res_i[0 * 32 + j];
            status |= res_i[2 * 32 + j];
            if (info & 0x100) {
                const int pl = info & 0xff, pos = res_i[1 * 32 + j];
                int cur = cur_s[pl * TTVB_TPB + tid];
                double a = acc_s[pl * TTVB_TPB + tid];
                while (cur < pos) { a += TTVB_PENALTY; cur++; }
                if ((info & 0x200) && cur == pos) { a += res_d[j]; cur++; }
                acc_s[pl * TTVB_TPB + tid] = a;
                cur_s[pl * TTVB_TPB + tid] = cur;
            }
        }
    }
    __syncwarp();
}

template <int N>
__global__ void __launch_bounds__(TTVB_TPB, TTVB_MINB)
ttvb_main_kernel(const __grid_constant__ DevSys S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- carve shared memory (offsets only: no pointer<->integer casts, so that every access
    //      below is compiled as LDS/STS rather than a generic load/store)
    // layout: [ring | staging][consts 5N][acc N][res_d][obs_time][obs_sigma] doubles, then
    //         [cur N][res_i][obs_epoch] ints
    double* const sd = reinterpret_cast<double*>(smem_raw);
    const int ring_d = 2 * 6 * N * TTVB_TPB, stage_d = TTVB_TPB * S.rowlen;
    const int o_cst = ring_d > stage_d ? ring_d : stage_d;
    const int o_acc = o_cst + 5 * N * TTVB_TPB;
    const int o_resd = o_acc + N * TTVB_TPB;
    const int o_obst = o_resd + (TTVB_TPB / 32) * 32 * 2;
    const int o_obss = o_obst + S.nobs;
    const int o_int = o_obss + S.nobs;                 // in doubles
    double* hist = sd;
    double* xstage = sd;               // aliases the ring: only used before the first step
    double* cst_s = sd + o_cst;
    double* gm_s = cst_s;
    double* kc_s = cst_s + N * TTVB_TPB;
    double* q_s = cst_s + 2 * N * TTVB_TPB;
    double* acc_s = sd + o_acc;
    double* res_d_all = sd + o_resd;
    double* obs_time_s = sd + o_obst;
    double* obs_sigma_s = sd + o_obss;
    int* const si = reinterpret_cast<int*>(smem_raw) + 2 * o_int;
    int* cur_s = si;
    int* res_i_all = si + N * TTVB_TPB;
    int* obs_epoch_s = res_i_all + (TTVB_TPB / 32) * 32 * 4;

    for (int i = tid; i < S.nobs; i += TTVB_TPB) {
        obs_time_s[i] = S.obs_time[i];
        obs_sigma_s[i] = S.obs_sigma[i];
        obs_epoch_s[i] = S.obs_epoch[i];
    }
    int* res_i = res_i_all + warp * 32 * 4;
    double* res_d = res_d_all + warp * 32 * 2;
    WarpQ Q;
    {
        const long long gw = (long long)blockIdx.x * (TTVB_TPB / 32) + warp;
        unsigned char* qb = S.queue + gw * S.queue_stride;
        Q.qi = (int*)qb;
        Q.qd = (double*)(qb + 32 * QLayout<N>::NI * 4);
    }
    const long long ntiles = (S.B + TTVB_TPB - 1) / TTVB_TPB;
    const int rowlen = S.rowlen;

    __shared__ int s_unit;
    const int nunits = (int)ntiles * S.nseg;
    for (;;) {
        __syncthreads();   // previous unit done with shared memory; obs table visible
        if (tid == 0) s_unit = atomicAdd(S.sched, 1);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= nunits) break;
        const int seg = unit / (int)ntiles;
        const long long tile = unit - (long long)seg * ntiles;
        const long long tile_base = tile * TTVB_TPB;
        const long long sys = tile_base + tid;
        bool alive = sys < S.B;
        int status = 0;
        SmemConsts<N> C{cst_s + tid};
        State<N> p;
        double prev_dot[N];
        int count[N];
        int trigmask = 0, nev = 0;
        double Time = S.t0, Tm1 = S.t0, Tm2 = S.t0;
        if (seg > 0) {
            // ---- resume: wait until the previous segment of this tile is published, reload the context
            if (tid == 0) {
                volatile int* flag = S.sched + 1 + tile;
                while (*flag < seg) __nanosleep(256);
            }
            __syncthreads();
            __threadfence();
            if (sys < S.B) {
                const double* cd = S.ctx_d + sys;
                const int* ci = S.ctx_i + sys;
                const long long cs = S.ctx_stride;
                int f = 0;
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { p.x[i][c] = __ldcg(cd + (f++) * cs); }
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { p.v[i][c] = __ldcg(cd + (f++) * cs); }
                for (int j = 0; j < 12 * N; j++) hist[(size_t)j * TTVB_TPB + tid] = __ldcg(cd + (f++) * cs);
#pragma unroll
                for (int i = 0; i < N; i++) prev_dot[i] = __ldcg(cd + (f++) * cs);
                for (int i = 0; i < N; i++) acc_s[i * TTVB_TPB + tid] = __ldcg(cd + (f++) * cs);
                for (int j = 0; j < 5 * N; j++) cst_s[(size_t)j * TTVB_TPB + tid] = __ldcg(cd + (f++) * cs);
                Time = __ldcg(cd + (f++) * cs); Tm1 = __ldcg(cd + (f++) * cs); Tm2 = __ldcg(cd + (f++) * cs);
                int g = 0;
#pragma unroll
                for (int i = 0; i < N; i++) count[i] = __ldcg(ci + (g++) * cs);
                for (int i = 0; i < N; i++) cur_s[i * TTVB_TPB + tid] = __ldcg(ci + (g++) * cs);
                trigmask = __ldcg(ci + (g++) * cs);
                nev = __ldcg(ci + (g++) * cs);
                status = __ldcg(ci + (g++) * cs);
                alive = __ldcg(ci + (g++) * cs) != 0;
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) { prev_dot[i] = 0.0; count[i] = 0; }
            }
        } else {
        // ---- first segment: coalesced staging of this tile's input rows, prologue
        {
            const long long nrow = (S.B - tile_base < TTVB_TPB) ? (S.B - tile_base) : TTVB_TPB;
            const long long nel = nrow * rowlen;
            const double* src = S.x + tile_base * rowlen;
            for (long long i = tid; i < nel; i += TTVB_TPB) xstage[i] = src[i];
        }
        __syncthreads();
        // ---- prologue: bounds, cube->physical, constants, elements -> Jacobi state
        if (alive) {
            const double* xr = xstage + (size_t)tid * rowlen;
            if (S.mode != 2) {
                bool inside = true;
                for (int i = 0; i < S.ndim; i++) {
                    const double xv = xr[i];
                    inside = inside && (S.lo[i] <= xv) && (xv <= S.hi[i]);
                }
                if (!inside) { status |= 1; alive = false; }
            }
        }
        if (alive) {
            const double* xr = xstage + (size_t)tid * rowlen;
            double eta_prev = S.gm0;
            double R[3] = {0.0, 0.0, 0.0}, V[3] = {0.0, 0.0, 0.0};
            bool valid = true;
#pragma unroll
            for (int i = 0; i < N; i++) {
                double e[7];
#pragma unroll
                for (int c = 0; c < 7; c++) {
                    const int k = 7 * i + c;
                    const int m = S.src[k];
                    double val;
                    if (S.mode == 2) val = xr[k];
                    else if (m < 0) val = S.cval[-1 - m];
                    else if (S.mode == 0) val = S.bi[m] + xr[m] * S.dbf[m];
                    else val = xr[m];
                    e[c] = val;
                }
                if (S.mode == 0) {
                    const double w = (e[4] + e[5]) / 2.0;
                    const double M = w - e[5];
                    e[4] = wrap360(w);
                    e[5] = wrap360(M);
                }
                PlanetInit pi;
                valid = planet_init(e, S.gm0, eta_prev, pi) && valid;
                C.set(i, pi.gm, pi.kc, pi.q, pi.ieta, pi.s);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    p.x[i][k] = pi.x[k] - R[k];
                    p.v[i][k] = pi.v[k] - V[k];
                    R[k] = fma(pi.q, p.x[i][k], R[k]);
                    V[k] = fma(pi.q, p.v[i][k], V[k]);
                }
            }
            if (!valid) { status |= 16; alive = false; }
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            prev_dot[i] = 0.0; count[i] = 0;
            acc_s[i * TTVB_TPB + tid] = 0.0;
            cur_s[i * TTVB_TPB + tid] = S.obs_start[i];
        }
        if (alive) {
            bool ok = corrector<N>(C, p, S.dt);
#pragma unroll
            for (int i = 0; i < N; i++) prev_dot[i] = sky_g(p.x[i][0], p.x[i][1], p.v[i][0], p.v[i][1]);
            if (ok) ok = map_A<N>(C, p, 0.5 * S.dt);
            if (!ok) { status |= 2; alive = false; }
        }
        }   // seg == 0
        const bool integrate = (sys < S.B) && !(status & 1);   // out of bounds: chi2=+inf
        // ---- integration of steps k0..k1 of this unit
        const long long k0 = (long long)seg * S.seg_len + 1;
        const long long k1 = (seg == S.nseg - 1) ? (S.nsteps + 1) : (long long)(seg + 1) * S.seg_len;
        int qcount = 0;
        const int warp_base_tid = warp * 32;
        const long long warp_base_sys = tile_base + warp_base_tid;
#pragma unroll 1
        for (long long k = k0; k <= k1; k++) {
            if (__ballot_sync(0xffffffffu, alive) == 0u) break;
            if (alive) {
                double* hs = hist + (size_t)((k - 1) & 1) * 6 * N * TTVB_TPB + tid;
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        hs[(i * 6 + c) * TTVB_TPB] = p.x[i][c];
                        hs[(i * 6 + 3 + c) * TTVB_TPB] = p.v[i][c];
                    }
                double rj[N], irj[N];
                map_B<N>(C, p, S.dt, rj, irj);
                if (!map_A_after_B<N>(C, p, S.dt, rj, irj)) { status |= 2; alive = false; trigmask = 0; }
            }
            Tm2 = Tm1; Tm1 = Time; Time += S.dt;
            // ---- push the events triggered at step k-1 (they now have p_{k-2}, p_{k-1}, p_k)
            if (__ballot_sync(0xffffffffu, trigmask != 0) != 0u) {
#pragma unroll 1
                for (int i = 0; i < N; i++) {
                    bool t = (trigmask >> i) & 1;
                    unsigned m = __ballot_sync(0xffffffffu, t);
                    while (m) {
                        const int room = 32 - qcount;
                        const int rank = __popc(m & ((1u << lane) - 1u));
                        if (t && rank < room) {
                            const int slot = qcount + rank;
                            int ci = 0;
#pragma unroll
                            for (int j = 0; j < N; j++) if (j == i) { ci = count[j]; count[j] = ci + 1; }
                            Q.qi[0 * 32 + slot] = lane;
                            Q.qi[1 * 32 + slot] = i;
                            Q.qi[2 * 32 + slot] = ci;
                            Q.qi[3 * 32 + slot] = nev;
                            nev++;
                            const double* h0 = hist + (size_t)(k & 1) * 6 * N * TTVB_TPB + tid;        // p_{k-2}
                            const double* h1 = hist + (size_t)((k - 1) & 1) * 6 * N * TTVB_TPB + tid;  // p_{k-1}
                            double* qd = Q.qd + slot;
#pragma unroll
                            for (int f = 0; f < 6 * N; f++) {
                                qd[(0 * 6 * N + f) * 32] = h0[f * TTVB_TPB];
                                qd[(1 * 6 * N + f) * 32] = h1[f * TTVB_TPB];
                            }
#pragma unroll
                            for (int j = 0; j < N; j++)
#pragma unroll
                                for (int c = 0; c < 3; c++) {
                                    qd[(2 * 6 * N + j * 6 + c) * 32] = p.x[j][c];
                                    qd[(2 * 6 * N + j * 6 + 3 + c) * 32] = p.v[j][c];
                                }
                            qd[(18 * N + 0) * 32] = Tm2;
                            qd[(18 * N + 1) * 32] = Tm1;
                            qd[(18 * N + 2) * 32] = Time;
                            t = false;
                        }
                        const int np = __popc(m);
                        qcount += (np < room) ? np : room;
                        if (qcount == 32) {
                            process_batch<N>(S, 32, Q, kc_s, q_s, gm_s, res_i, res_d, warp_base_tid, warp_base_sys, obs_epoch_s, obs_time_s, obs_sigma_s);
                            consume_results(lane, tid, res_i, res_d, acc_s, cur_s, status);
                            qcount = 0;
                        }
                        m = __ballot_sync(0xffffffffu, t);
                    }
                }
                trigmask = 0;
            }
            // ---- trigger check of step k (unsynchronised Jacobi state, as TTVFast)
            if (alive && k <= S.nsteps) {
#pragma unroll
                for (int i = 0; i < N; i++) {
                    const double cd = sky_g(p.x[i][0], p.x[i][1], p.v[i][0], p.v[i][1]);
                    if (prev_dot[i] < 0.0 && cd > 0.0 && p.x[i][2] > 0.0) trigmask |= (1 << i);
                    prev_dot[i] = cd;
                }
            }
        }
        // ---- drain the queue
        {
            const int cnt = qcount;   // warp-uniform
            if (cnt > 0) {
                process_batch<N>(S, cnt, Q, kc_s, q_s, gm_s, res_i, res_d, warp_base_tid, warp_base_sys, obs_epoch_s, obs_time_s, obs_sigma_s);
                consume_results(lane, tid, res_i, res_d, acc_s, cur_s, status);
            }
        }
        if (seg < S.nseg - 1) {
            // ---- hand the tile over: save the context, publish the segment
            if (sys < S.B) {
                double* cd = S.ctx_d + sys;
                int* ci = S.ctx_i + sys;
                const long long cs = S.ctx_stride;
                int f = 0;
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { cd[(f++) * cs] = p.x[i][c]; }
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { cd[(f++) * cs] = p.v[i][c]; }
                for (int j = 0; j < 12 * N; j++) cd[(f++) * cs] = hist[(size_t)j * TTVB_TPB + tid];
#pragma unroll
                for (int i = 0; i < N; i++) cd[(f++) * cs] = prev_dot[i];
                for (int i = 0; i < N; i++) cd[(f++) * cs] = acc_s[i * TTVB_TPB + tid];
                for (int j = 0; j < 5 * N; j++) cd[(f++) * cs] = cst_s[(size_t)j * TTVB_TPB + tid];
                cd[(f++) * cs] = Time; cd[(f++) * cs] = Tm1; cd[(f++) * cs] = Tm2;
                int g = 0;
#pragma unroll
                for (int i = 0; i < N; i++) ci[(g++) * cs] = count[i];
                for (int i = 0; i < N; i++) ci[(g++) * cs] = cur_s[i * TTVB_TPB + tid];
                ci[(g++) * cs] = trigmask;
                ci[(g++) * cs] = nev;
                ci[(g++) * cs] = status;
                ci[(g++) * cs] = alive ? 1 : 0;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(S.sched + 1 + tile, seg + 1);
            continue;
        }
        // ---- epilogue: remaining observed epochs are missing -> penalty; totals
        if (sys < S.B) {
            double chi2, logl;
            if (!integrate) {
                chi2 = INFINITY; logl = -INFINITY;
                if (S.chi2_planet_out)
                    for (int i = 0; i < N; i++) S.chi2_planet_out[sys * N + i] = chi2;
            } else {
                double tot = 0.0;
                for (int i = 0; i < N; i++) {
                    int cur = cur_s[i * TTVB_TPB + tid];
                    const int end = S.obs_start[i + 1];
                    double a = acc_s[i * TTVB_TPB + tid];
                    for (; cur < end; cur++) a += TTVB_PENALTY;
                    if (S.obs_start[i + 1] > S.obs_start[i]) tot += a;
                    if (S.chi2_planet_out) S.chi2_planet_out[sys * N + i] = a;
                }
                chi2 = tot;
                logl = -0.5 * chi2 - S.second_term;
            }
            if (S.ev_time != nullptr) {
                if (nev > S.cap) { status |= 8; nev = S.cap; }
                S.n_events[sys] = nev;
            }
            if (S.chi2_out) S.chi2_out[sys] = chi2;
            if (S.logl_out) S.logl_out[sys] = logl;
            if (S.status_out) S.status_out[sys] = status;
        }
    }
}

// cube_to_physical for B rows (utils.py:403-438), one thread per row
__global__ void ttvb_cube_to_physical_kernel(const __grid_constant__ DevSys S, double* flat_out,
                                             int* inside_out) {
    const long long sys = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= S.B) return;
    const double* xr = S.x + sys * S.ndim;
    bool inside = true;
    for (int i = 0; i < S.ndim; i++) inside = inside && (S.lo[i] <= xr[i]) && (xr[i] <= S.hi[i]);
    if (inside_out) inside_out[sys] = inside ? 1 : 0;
    double* fo = flat_out + sys * 7 * S.npla;
    for (int k = 0; k < 7 * S.npla; k++) {
        const int m = S.src[k];
        fo[k] = (m < 0) ? S.cval[-1 - m] : (S.bi[m] + xr[m] * S.dbf[m]);
    }
    for (int i = 0; i < S.npla; i++) {
        const double w = (fo[7 * i + 4] + fo[7 * i + 5]) / 2.0;
        const double M = w - fo[7 * i + 5];
        fo[7 * i + 4] = wrap360(w);
        fo[7 * i + 5] = wrap360(M);
    }
}

// FP64 FMA-chain microbenchmark: 8 independent chains per thread
__global__ void ttvb_fp64_peak_kernel(double* out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
}

}  // namespace ttvb
