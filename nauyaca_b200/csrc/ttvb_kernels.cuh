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
// L2 resident): a lane that triggers pushes its three bracketing states; when 64 events
// are queued the whole warp refines them together, two events per lane, and the results
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
// row stride (doubles) of the history ring [slot][field][thread]: odd, so that 32 lanes reading
// 32 different fields of ONE thread (cooperative event push) hit 32 different banks
#define TTVB_RS (TTVB_TPB + 1)
#define TTVB_ST_BAD_TRANSIT_BIT 4

// Self-checks (compute-sanitizer is not available on the GPU pool, profiles/r2a_compute_sanitizer_closed.txt):
// a -DTTVB_CHECKS build verifies, on the device, the indices of every queue / ring / result-area /
// observed-table / event-table access and bounds the inter-CTA waits, counting violations by kind
// in a device array (read back with ttvb_check_failures). The product build compiles them away.
#ifdef TTVB_CHECKS
__device__ int g_ttvb_check[8];
#define TTVB_CHECK(cond, kind) do { if (!(cond)) atomicAdd(&g_ttvb_check[(kind)], 1); } while (0)
#else
#define TTVB_CHECK(cond, kind) do { } while (0)
#endif
enum { CHK_QUEUE_SLOT = 0, CHK_BATCH_HEADER = 1, CHK_FOLD = 2, CHK_OBS_INDEX = 3, CHK_TABLE_INDEX = 4, CHK_HANDOVER_WAIT = 5,
       CHK_UNIT = 6, CHK_KINDS = 8 };

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
    // streaming input (host-pointer calls): number of rows of `x` uploaded so far, bumped by the
    // copy stream while the kernel runs; a tile waits until its rows are there. nullptr: all there.
    const int* ready;
    // thin launches (latency-bound batches that leave schedulers idle): only the first `lanes` lanes
    // of a warp hold a system, so that the batch spreads over more warps and every warp pushes and
    // refines fewer events (plan_launch chooses; 32 = every lane)
    int lanes;
    double* ctx_d;            // [25N+3][ctx_stride]
    int* ctx_i;               // [2N+4][ctx_stride]
    long long ctx_stride;
};

// ---- queue layout of one warp: QCAP slots, slot-major. Header int4 (owner, planet, epoch, seq) per
//      slot, then SD doubles per slot: 3 states x N planets x 6 (x, v), T3, one pad (every pair
//      of doubles is 16-byte aligned: the pusher writes and the refining lanes read 128 bits at a
//      time, and the cooperative copy of the ring states is one coalesced store).
//      QCAP = 64: a batch is refined two events per lane in lockstep.
#define TTVB_QCAP 64
template <int N>
struct QLayout {
    static constexpr int SD = 18 * N + 4;
    static constexpr long long BYTES = (long long)TTVB_QCAP * (16 + SD * 8);
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

// The dynamic shared memory as an array of doubles (every `extern __shared__` array starts at the same
// address). The step loop indexes THIS array with integers: a pointer derived from it is a generic
// pointer to the compiler, which rebuilt its shared-window base (S2R SR_CgaCtaId + LEA, ~35 exposed
// cycles for a lone warp) on every pass of the loop.
extern __shared__ __align__(16) double ttvb_smem_d[];
template <int N>
struct SmemConstsIdx {
    int off;     // o_cst + tid
    __device__ __forceinline__ double gm(int i) const { return ttvb_smem_d[off + (0 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double kc(int i) const { return ttvb_smem_d[off + (1 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double q(int i) const { return ttvb_smem_d[off + (2 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double ieta(int i) const { return ttvb_smem_d[off + (3 * N + i) * TTVB_TPB]; }
    __device__ __forceinline__ double s(int i) const { return ttvb_smem_d[off + (4 * N + i) * TTVB_TPB]; }
};

// The same constants for one pass through the step loop, with the one the kick needs first (q_0, the
// first operand of its first multiplication) already in a register: its shared-memory load is issued
// ahead of the history-ring stores instead of right in front of its use.
template <int N>
struct StepConsts {
    const SmemConstsIdx<N>& c;
    double q0;
    __device__ __forceinline__ double gm(int i) const { return c.gm(i); }
    __device__ __forceinline__ double kc(int i) const { return c.kc(i); }
    __device__ __forceinline__ double q(int i) const { return i == 0 ? q0 : c.q(i); }
    __device__ __forceinline__ double ieta(int i) const { return c.ieta(i); }
    __device__ __forceinline__ double s(int i) const { return c.s(i); }
};

// Stores into the warp's event queue carry an L2 evict-last policy: the queues are rewritten
// every few dozen steps and should stay resident in L2 instead of being written back to HBM.
__device__ __forceinline__ unsigned long long q_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void q_st(double* p, double v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void q_st2(double* p, double a, double b, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void q_st4(int* p, int a, int b, int c, int d, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
                 "l"(pol) : "memory");
}

struct WarpQ {
    int* qi;
    double* qd;
};

// shared-memory carve-up (dynamic); the input staging area aliases the history ring
__host__ __device__ inline size_t smem_bytes(int N, int nobs, int rowlen) {
    size_t b = 0;
    b += (size_t)nobs * (8 + 8 + 4);
    b = (b + 7) & ~(size_t)7;
    size_t hist = (size_t)2 * 6 * N * TTVB_RS * 8, stage = (size_t)TTVB_TPB * rowlen * 8;
    b += hist > stage ? hist : stage;                       // history ring / input staging
    b += (size_t)5 * N * TTVB_TPB * 8;                      // gm, kc, q, ieta, s
    b += (size_t)N * TTVB_TPB * 8;                          // chi2 accumulators
    b += (size_t)(TTVB_TPB / 32) * TTVB_QCAP * 8;           // result area (doubles)
    b += (size_t)N * TTVB_TPB * 4;                          // observed-table cursors
    b += (size_t)(TTVB_TPB / 32) * TTVB_QCAP * 4 * 4;       // result area (ints)
    return b;
}

// Refine up to 64 queued events, two per lane IN LOCKSTEP (slots lane and lane+32: two
// independent dependency chains per lane, which is what hides the FP64 latency here), and
// hand the results to their owners. Per event the operation sequence is that of the scalar
// form (kepler_drift / locate_bracket in ttvb_math.cuh), so the bits are the oracle's.
template <int N, bool COMPACT>
__device__ __noinline__ void process_batch(const DevSys& S, int count, WarpQ Q, const double* kc_s,
                                           const double* q_s, const double* gm_s, int* res_i,
                                           double* res_d, int warp_base_tid, long long warp_base_sys,
                                           const int* obs_epoch_s, const double* obs_time_s,
                                           const double* obs_sigma_s) {
    constexpr int CAP = TTVB_QCAP;
    constexpr int SD = QLayout<N>::SD;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    bool act[2];
    int sl[2], owner[2], planet[2], epoch[2], seq[2], otid[2];
    double T3[2][3];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int slot = lane + 32 * e;
        act[e] = slot < count;
        sl[e] = act[e] ? slot : (lane < count ? lane : 0);     // idle chains shadow a valid event
        const int4 hd = __ldcg(reinterpret_cast<const int4*>(Q.qi) + sl[e]);
        owner[e] = hd.x; planet[e] = hd.y; epoch[e] = hd.z; seq[e] = hd.w;
        TTVB_CHECK(count >= 1 && count <= CAP && sl[e] >= 0 && sl[e] < count, CHK_QUEUE_SLOT);
        TTVB_CHECK(owner[e] >= 0 && owner[e] < 32 && planet[e] >= 0 && planet[e] < N && epoch[e] >= 0 && seq[e] >= 0,
                   CHK_BATCH_HEADER);
        otid[e] = warp_base_tid + owner[e];
        const double* tq = Q.qd + (long long)sl[e] * SD + 18 * N;
        const double2 t01 = __ldcg(reinterpret_cast<const double2*>(tq));
        T3[e][0] = t01.x; T3[e][1] = t01.y; T3[e][2] = __ldcg(tq + 2);
    }
    const double hdt = 0.5 * S.dt;
    // xs/vsn[e][0]: synchronised astrocentric state of the middle node k; [e][1]: of node k-1
    // (which == 0) or k+1 (which == 2)
    double xs[2][2][3] = {}, vsn[2][2][3] = {}, gs[2][2];
    int which[2] = {1, 1};
    bool ok[2] = {true, true};
    // COMPACT: the two passes are a RUN-TIME loop (one copy of the drift code instead of two). The
    // kernel is large and its warps are at different places of it; when the GPU is full of warps that
    // refine often (the 10^7 sweep: 2.3 events per warp-step, `no_instruction` the second stall
    // reason) the smaller footprint is worth 8 %; a sparse or latency-bound launch is ~1 % faster with
    // the passes unrolled. The host picks the variant per launch (dispatch_launch).
#pragma unroll (COMPACT ? 1 : 2)
    for (int pass = 0; pass < 2; pass++) {
        double R[2][3], V[2][3], xt[2][3], vt[2][3];
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int k = 0; k < 3; k++) { R[e][k] = 0.0; V[e][k] = 0.0; xt[e][k] = 0.0; vt[e][k] = 0.0; }
#pragma unroll 1
        for (int j = 0; j < N; j++) {
            const bool on0 = j <= planet[0], on1 = j <= planet[1];
            if (!__any_sync(0xffffffffu, on0 || on1)) break;
            State<2> st;
            MuArray<2> C;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const double2* base = reinterpret_cast<const double2*>(Q.qd + (long long)sl[e] * SD + which[e] * 6 * N + j * 6);
                const double2 a = __ldcg(base), b = __ldcg(base + 1), c = __ldcg(base + 2);
                st.x[e][0] = a.x; st.x[e][1] = a.y; st.x[e][2] = b.x;
                st.v[e][0] = b.y; st.v[e][1] = c.x; st.v[e][2] = c.y;
                C.m[e] = kc_s[j * TTVB_TPB + otid[e]];
            }
            bool conv[2];
            map_A_joint_t<2, MuArray<2>, false, UniformTau>(C, st, UniformTau{-hdt}, nullptr, nullptr, conv);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const bool on = j <= planet[e];
                ok[e] = ok[e] && (conv[e] || !on);
                const double qj = q_s[j * TTVB_TPB + otid[e]];
                const bool last = j == planet[e];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double xr = st.x[e][k] + R[e][k], vr = st.v[e][k] + V[e][k];
                    xt[e][k] = last ? xr : xt[e][k];
                    vt[e][k] = last ? vr : vt[e][k];
                    const double Rn = fma(qj, st.x[e][k], R[e][k]), Vn = fma(qj, st.v[e][k], V[e][k]);
                    R[e][k] = (j < planet[e]) ? Rn : R[e][k];
                    V[e][k] = (j < planet[e]) ? Vn : V[e][k];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const double g = sky_g(xt[e][0], xt[e][1], vt[e][0], vt[e][1]);
            if (pass == 0) {
                gs[e][0] = g;
                which[e] = (g > 0.0) ? 0 : 2;
#pragma unroll
                for (int k = 0; k < 3; k++) { xs[e][0][k] = xt[e][k]; vsn[e][0][k] = vt[e][k]; }
            } else {
                gs[e][1] = g;
#pragma unroll
                for (int k = 0; k < 3; k++) { xs[e][1][k] = xt[e][k]; vsn[e][1][k] = vt[e][k]; }
            }
        }
    }
    // ---- the two sides of both brackets: behind sides together, then ahead sides together
    double mu[2], Tb[2], Ta[2], tau_b[2], tau_a[2], xb[2][3], vb[2][3], xa[2][3], va[2][3];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        mu[e] = S.gm0 + gm_s[planet[e] * TTVB_TPB + otid[e]];
        const bool first = which[e] == 0;     // bracket [T_{k-1}, T_k]; else [T_k, T_{k+1}]
        const double gb = first ? gs[e][1] : gs[e][0], ga = first ? gs[e][0] : gs[e][1];
        Tb[e] = first ? T3[e][0] : T3[e][1];
        Ta[e] = first ? T3[e][1] : T3[e][2];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            xb[e][k] = first ? xs[e][1][k] : xs[e][0][k];
            vb[e][k] = first ? vsn[e][1][k] : vsn[e][0][k];
            xa[e][k] = first ? xs[e][0][k] : xs[e][1][k];
            va[e][k] = first ? vsn[e][0][k] : vsn[e][1][k];
        }
        first_guess(mu[e], S.dt, xb[e], vb[e], gb, xa[e], va[e], ga, tau_b[e], tau_a[e]);
    }
    double rb[2], vbk[2], ra[2], vak[2];
    bool okb[2], oka[2];
    locate_side_joint<2>(mu, xb, vb, tau_b, rb, vbk, okb);
    locate_side_joint<2>(mu, xa, va, tau_a, ra, vak, oka);
    // ---- results; this event's chi2 term, computed here by all lanes at once: events with
    //      rsky > Rstar are not transits (utils.py:113); pos = first observed epoch of the planet
    //      that is >= the event's epoch (observed epochs before it that the owner has not met
    //      are missing from the simulation, utils.py:199-206)
#pragma unroll
    for (int e = 0; e < 2; e++) {
        double tt = TTVB_BAD_TRANSIT, rs = TTVB_BAD_TRANSIT, vs = TTVB_BAD_TRANSIT;
        int flags = 0;
        if (ok[e] && okb[e] && oka[e]) {
            const double tb = Tb[e] + tau_b[e], ta = Ta[e] + tau_a[e];
            const double wa = tau_b[e] / (tau_b[e] - tau_a[e]);
            tt = fma(wa, ta - tb, tb);
            rs = fma(wa, ra[e] - rb[e], rb[e]);
            vs = fma(wa, vak[e] - vbk[e], vbk[e]);
        } else {
            flags = TTVB_ST_BAD_TRANSIT_BIT;
        }
        if (act[e] && S.ev_time != nullptr) {
            const long long sys = warp_base_sys + owner[e];
            if (seq[e] < S.cap) {
                const long long o = sys * (long long)S.cap + seq[e];
                TTVB_CHECK(sys >= 0 && sys < S.B && o >= 0 && o < S.B * (long long)S.cap, CHK_TABLE_INDEX);
                S.ev_planet[o] = planet[e]; S.ev_epoch[o] = epoch[e];
                S.ev_time[o] = tt; S.ev_rsky[o] = rs; S.ev_vsky[o] = vs;
            }
        }
        int info = planet[e], pos = 0;
        double term = 0.0;
        if (act[e] && rs <= S.rstar_au) {
            info |= 0x100;
            int a = S.obs_start[planet[e]];
            const int end = S.obs_start[planet[e] + 1];
            int b = end;
            while (a < b) {
                const int m = (a + b) >> 1;
                if (obs_epoch_s[m] < epoch[e]) a = m + 1; else b = m;
            }
            pos = a;
            TTVB_CHECK(pos >= S.obs_start[planet[e]] && pos <= end && end <= S.nobs, CHK_OBS_INDEX);
            if (pos < end && obs_epoch_s[pos] == epoch[e]) {
                info |= 0x200;
                const double d = (obs_time_s[pos] - tt) / obs_sigma_s[pos];
                term = d * d;
            }
        }
        const int slot = lane + 32 * e;
        res_i[0 * CAP + slot] = info;
        res_i[1 * CAP + slot] = pos;
        res_i[2 * CAP + slot] = act[e] ? flags : 0;
        res_i[3 * CAP + slot] = 0;
        res_d[slot] = term;
    }
    __syncwarp();
    // ---- per owner lane, the set of events it owns (queue order = bit order, low word first)
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const unsigned peers = __match_any_sync(0xffffffffu, act[e] ? owner[e] : 32 + lane);
        if (act[e] && (peers & ((1u << lane) - 1u)) == 0u) res_i[3 * CAP + 32 * e + owner[e]] = (int)peers;
    }
    __syncwarp();
}

// Every lane folds the events it owns into its chi2, in queue order (which keeps the summation
// order of the reference, utils.py:161-216): observed epochs skipped by the simulation cost
// 1e20 each (utils.py:199-206), then the event's own term.
__device__ __forceinline__ void consume_results(int lane, int tid, const int* res_i, const double* res_d,
                                                double* acc_s, int* cur_s, int& status) {
    constexpr int CAP = TTVB_QCAP;
#pragma unroll 1
    for (int e = 0; e < 2; e++) {
        unsigned mine = (unsigned)res_i[3 * CAP + 32 * e + lane];
        while (__any_sync(0xffffffffu, mine != 0u)) {
            if (mine != 0u) {
                const int j = 32 * e + __ffs(mine) - 1;
                mine &= mine - 1u;
                TTVB_CHECK(j >= 0 && j < CAP, CHK_FOLD);
                const int info = res_i[0 * CAP + j];
                status |= res_i[2 * CAP + j];
                if (info & 0x100) {
                    const int pl = info & 0xff, pos = res_i[1 * CAP + j];
                    int cur = cur_s[pl * TTVB_TPB + tid];
                    double a = acc_s[pl * TTVB_TPB + tid];
                    TTVB_CHECK(pl >= 0 && pl < TTVB_MAX_PLANETS_P1 - 1 && cur >= 0 && pos >= 0 && pos - cur <= 100000, CHK_FOLD);
                    while (cur < pos) { a += TTVB_PENALTY; cur++; }
                    if ((info & 0x200) && cur == pos) { a += res_d[j]; cur++; }
                    acc_s[pl * TTVB_TPB + tid] = a;
                    cur_s[pl * TTVB_TPB + tid] = cur;
                }
            }
        }
    }
    __syncwarp();
}

// Per-thread state that crosses the boundary of the step loop (lives in local memory while
// the caller refines a batch of events; in registers inside run_steps).
template <int N>
struct StepCtx {
    State<N> p;
    double prev_dot[N];
    double Time, Tm1, Tm2;
    long long k;              // next step
    int count[N];             // transit epochs counted so far, per planet
    int trigmask, nev, status, alive, qcount, pending_push;
};

// The step loop of one warp: steps cx.k .. k1 (kick, drift, history ring, push of triggered
// events, trigger check). A LEAF function: no calls inside, so its register allocation does
// not depend on the event refinement. Returns true when the warp's event queue is full (the
// caller refines the batch and calls again), false when the steps are done.
template <int N>
__device__ __noinline__ bool run_steps(StepCtx<N>& cx, const double dt, const long long nsteps, const long long k1,
                                       const int o_cst, int* __restrict__ qi, double* __restrict__ qdq) {
    const int tid = threadIdx.x, lane = tid & 31;
    double* const hist = ttvb_smem_d;            // the history ring starts the dynamic shared memory
    const SmemConstsIdx<N> C{o_cst + tid};
    State<N> p = cx.p;
    double prev_dot[N];
    int count[N];
#pragma unroll
    for (int i = 0; i < N; i++) { prev_dot[i] = cx.prev_dot[i]; count[i] = cx.count[i]; }
    int trigmask = cx.trigmask, nev = cx.nev, status = cx.status, qcount = cx.qcount;
    bool alive = cx.alive != 0;
    bool pending = cx.pending_push != 0;
    double Time = cx.Time, Tm1 = cx.Tm1, Tm2 = cx.Tm2;
    long long k = cx.k;
    bool full = false;
    constexpr int CPR = (12 * N + 31) / 32;
    // Two warp votes per step tell what the push phase and the loop head need to know: whether some
    // lane holds an untold event, and whether some lane still has a live system. They are taken at the top of
    // a step, on what the trigger check of the step before has left, and consumed after the arithmetic
    // of the step (votes, not a REDUX reduction: its result comes back through a uniform register that
    // the compiler copies to a vector register at once, and an in-order warp then waits for it).
    unsigned ev_lanes = __ballot_sync(0xffffffffu, trigmask != 0);
    bool warp_alive = __any_sync(0xffffffffu, alive) != 0;
    // The loop head looks at the vote of the step BEFORE the last (a warp whose systems have all
    // stopped takes one more step on dead registers and leaves then).
    bool some_alive = warp_alive;
    // the step counters as 32-bit distances (at most 2e9 steps, ttvb_create): steps left in this call
    // and steps left in the integration, so that the tests of the loop read no 64-bit parameter
    int togo = (int)(k1 - k), left = (int)(nsteps - k);
    // (ptxas lays the blocks of the loop out in program order whatever the source says - the rare
    // Kepler iterations and the push phase moved behind the return with gotos came back in line - so
    // the common path jumps over both; a lone warp pays a fetch redirect for each.)
    // a call that comes back from a refinement in the middle of a push phase re-enters there
    if (pending) { pending = false; goto push_phase; }
#pragma unroll 1
    for (;;) {
        if (togo < 0) break;
        if (!some_alive) break;
        ev_lanes = __ballot_sync(0xffffffffu, trigmask != 0);
        warp_alive = __any_sync(0xffffffffu, alive) != 0;
        // every lane steps; one without a live system computes on whatever its registers
        // hold, on the straight-line path (idle flag), and nothing of it is ever used
        {
            const StepConsts<N> Cs{C, C.q(0)};
            const int hs = (int)((k - 1) & 1) * 6 * N * TTVB_RS + tid;
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    hist[hs + (i * 6 + c) * TTVB_RS] = p.x[i][c];
                    hist[hs + (i * 6 + 3 + c) * TTVB_RS] = p.v[i][c];
                }
            double rj[N], irj[N];
            map_B<N>(Cs, p, dt, rj, irj);
            if (!map_A_after_B<N>(Cs, p, dt, rj, irj, !alive)) { status |= 2; alive = false; trigmask = 0; }
        }
        Tm2 = Tm1; Tm1 = Time; Time += dt;
    push_phase:
        // ---- push the events triggered at step k-1 (they now have p_{k-2}, p_{k-1}, p_k). The two
        //      older states are in the ring: all 32 lanes copy them together, one field each; the
        //      owner adds p_k from its registers and the header.
        some_alive = warp_alive;
        if (ev_lanes != 0u) {
            __syncwarp();
            const unsigned long long pol = q_policy();
            const int par_old = (int)(k & 1) * 6 * N * TTVB_RS;     // ring slot of p_{k-2}; the other holds p_{k-1}
            // cooperative copy of the two ring states of one thread (12N doubles): lane handles
            // element lane + 32 r; element e = s*6N + f is field f of ring state s (0: p_{k-2},
            // 1: p_{k-1}). Computed here (from a fresh read of the lane id) so that nothing of it
            // stays live across the arithmetic of the step.
            int cp_src[CPR], cp_dst[CPR];
            bool cp_ok[CPR];
            {
                unsigned tx;
                asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tx));
    #pragma unroll
                for (int r = 0; r < CPR; r++) {
                    const int e = (int)(tx & 31u) + 32 * r;
                    cp_ok[r] = e < 12 * N;
                    const bool nw = e >= 6 * N;
                    const int f = nw ? e - 6 * N : e;
                    cp_src[r] = f * TTVB_RS + (nw ? (6 * N * TTVB_RS - par_old) : par_old);
                    cp_dst[r] = e;
                }
            }
            // rounds over the LANES that still hold an untold event (usually one round: a lane
            // rarely triggers for two planets in one step). Per round: the ring states of every such
            // lane are copied by the whole warp, then all owners store their part side by side. A
            // lane tells its events in ascending planet order, so the per-system order of the
            // events (and of the chi2 sum) is the reference's.
            // (a fresh ballot: a lane whose system has just stopped drops its pending events, so the
            // vote taken at the top of the step may name lanes that have nothing to tell any more)
            unsigned evl = __ballot_sync(0xffffffffu, trigmask != 0);
            while (evl != 0u) {
                const int room = TTVB_QCAP - qcount;
                const int rank = __popc(evl & ((1u << lane) - 1u));
                {
                    unsigned mm = evl;
                    int taken = 0;
                    while (mm != 0u && taken < room) {
                        const int b = __ffs(mm) - 1;
                        mm &= mm - 1u;
                        const int src = tid - lane + b;
                        TTVB_CHECK(qcount + taken >= 0 && qcount + taken < TTVB_QCAP && b >= 0 && b < 32, CHK_QUEUE_SLOT);
                        double* dst = qdq + (qcount + taken) * QLayout<N>::SD;
                        taken++;
    #pragma unroll
                        for (int r = 0; r < CPR; r++)
                            if (cp_ok[r]) q_st(dst + cp_dst[r], hist[src + cp_src[r]], pol);
                    }
                }
                if (trigmask != 0 && rank < room) {
                    const int i = __ffs(trigmask) - 1;
                    const int slot = qcount + rank;
                    TTVB_CHECK(slot >= 0 && slot < TTVB_QCAP && i < N, CHK_QUEUE_SLOT);
                    int ci = 0;
    #pragma unroll
                    for (int j = 0; j < N; j++) if (j == i) { ci = count[j]; count[j] = ci + 1; }
                    q_st4(qi + 4 * slot, lane, i, ci, nev, pol);
                    nev++;
                    double* qd = qdq + slot * QLayout<N>::SD + 2 * 6 * N;
    #pragma unroll
                    for (int j = 0; j < N; j++) {
                        q_st2(qd + j * 6 + 0, p.x[j][0], p.x[j][1], pol);
                        q_st2(qd + j * 6 + 2, p.x[j][2], p.v[j][0], pol);
                        q_st2(qd + j * 6 + 4, p.v[j][1], p.v[j][2], pol);
                    }
                    q_st2(qd + 6 * N, Tm2, Tm1, pol);
                    q_st(qd + 6 * N + 2, Time, pol);
                    trigmask &= trigmask - 1;
                }
                const int np = __popc(evl);
                qcount += (np < room) ? np : room;
                if (qcount == TTVB_QCAP) { full = true; break; }
                evl = __ballot_sync(0xffffffffu, trigmask != 0);
            }
            __syncwarp();
            if (full) { pending = true; break; }   // come back to the push phase of this step
        }
        // ---- trigger check of step k (unsynchronised Jacobi state, as TTVFast)
        //      (straight-line: what a stopped system or the extra last step leaves in prev_dot is never read)
        {
            const bool chk = alive && left >= 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                const double cd = sky_g(p.x[i][0], p.x[i][1], p.v[i][0], p.v[i][1]);
                const bool trig = chk && prev_dot[i] < 0.0 && cd > 0.0 && p.x[i][2] > 0.0;
                trigmask |= trig ? (1 << i) : 0;
                prev_dot[i] = cd;
            }
        }
        k++; togo--; left--;
    }
    cx.p = p;
#pragma unroll
    for (int i = 0; i < N; i++) { cx.prev_dot[i] = prev_dot[i]; cx.count[i] = count[i]; }
    cx.trigmask = trigmask; cx.nev = nev; cx.status = status; cx.alive = alive ? 1 : 0;
    cx.qcount = qcount; cx.pending_push = pending ? 1 : 0;
    cx.Time = Time; cx.Tm1 = Tm1; cx.Tm2 = Tm2; cx.k = k;
    return full;
}

template <int N, bool COMPACT>
__global__ void __launch_bounds__(TTVB_TPB, TTVB_MINB)
ttvb_main_kernel(const __grid_constant__ DevSys S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- carve shared memory (offsets only: no pointer<->integer casts, so that every access
    //      below is compiled as LDS/STS rather than a generic load/store)
    // layout: [ring | staging][consts 5N][acc N][res_d][obs_time][obs_sigma] doubles, then
    //         [cur N][res_i][obs_epoch] ints
    double* const sd = reinterpret_cast<double*>(smem_raw);
    const int ring_d = 2 * 6 * N * TTVB_RS, stage_d = TTVB_TPB * S.rowlen;
    const int o_cst = ring_d > stage_d ? ring_d : stage_d;
    const int o_acc = o_cst + 5 * N * TTVB_TPB;
    const int o_resd = o_acc + N * TTVB_TPB;
    const int o_obst = o_resd + (TTVB_TPB / 32) * TTVB_QCAP;
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
    int* obs_epoch_s = res_i_all + (TTVB_TPB / 32) * TTVB_QCAP * 4;

    for (int i = tid; i < S.nobs; i += TTVB_TPB) {
        obs_time_s[i] = S.obs_time[i];
        obs_sigma_s[i] = S.obs_sigma[i];
        obs_epoch_s[i] = S.obs_epoch[i];
    }
    int* res_i = res_i_all + warp * TTVB_QCAP * 4;
    double* res_d = res_d_all + warp * TTVB_QCAP;
    WarpQ Q;
    {
        const long long gw = (long long)blockIdx.x * (TTVB_TPB / 32) + warp;
        unsigned char* qb = S.queue + gw * S.queue_stride;
        Q.qi = (int*)qb;
        Q.qd = (double*)(qb + TTVB_QCAP * 16);
    }
    const int spt = (TTVB_TPB / 32) * S.lanes;        // systems per tile
    const long long ntiles = (S.B + spt - 1) / spt;
    const int slot_in_tile = warp * S.lanes + lane;   // of the systems this lane holds, if it holds one
    const int rowlen = S.rowlen;

    __shared__ int s_unit;
    __shared__ int s_rows_ok;      // streamed input: the rows of the tile have arrived (its own word: s_unit may still be being read)
    const int nunits = (int)ntiles * S.nseg;
    for (;;) {
        __syncthreads();   // previous unit done with shared memory; obs table visible
        if (tid == 0) s_unit = atomicAdd(S.sched, 1);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= nunits) break;
        const int seg = unit / (int)ntiles;
        const long long tile = unit - (long long)seg * ntiles;
        TTVB_CHECK(unit >= 0 && seg >= 0 && seg < S.nseg && tile >= 0 && tile < ntiles, CHK_UNIT);
        const long long tile_base = tile * spt;
        const long long sys = tile_base + slot_in_tile;
        const bool mine = lane < S.lanes && sys < S.B;   // this lane holds a system of the batch
        bool alive = mine;
        int status = 0;
        SmemConsts<N> C{cst_s + tid};
        State<N> p = {};
        double prev_dot[N];
        int count[N];
        int trigmask = 0, nev = 0;
        double Time = S.t0, Tm1 = S.t0, Tm2 = S.t0;
        if (seg > 0) {
            // ---- resume: wait until the previous segment of this tile is published, reload the context
            if (tid == 0) {
                volatile int* flag = S.sched + 1 + tile;
#ifdef TTVB_CHECKS
                const long long t_start = clock64();
                while (*flag < seg) {
                    __nanosleep(256);
                    if (clock64() - t_start > 20000000000LL) { TTVB_CHECK(false, CHK_HANDOVER_WAIT); break; }
                }
#else
                while (*flag < seg) __nanosleep(256);
#endif
                __threadfence();     // acquire: the observation of the flag, then a fence, in the observing thread
            }
            __syncthreads();
            __threadfence();
            if (mine) {
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
                for (int j = 0; j < 12 * N; j++) hist[(size_t)j * TTVB_RS + tid] = __ldcg(cd + (f++) * cs);
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
            const long long nrow = (S.B - tile_base < spt) ? (S.B - tile_base) : spt;
            bool arrived = true;
            if (S.ready != nullptr) {
                // streamed input: wait until the copy stream has delivered this tile's rows (rare: the
                // uploads run ahead of the kernel). A bounded wait (~0.3 s), so that uploads that cannot
                // run under the kernel (a failed copy, a profiler serialising the GPU) cannot hang it:
                // the tile gives up, flags it, and the host runs the launch again once the rows are there.
                if (tid == 0) {
                    const int need = (int)(tile_base + nrow);
                    const long long t_start = clock64();
                    int ok = 1;
                    while (*reinterpret_cast<const volatile int*>(S.ready) < need) {
                        __nanosleep(1000);
                        // gave up before? then the rest of the launch gives up at once (the host reruns it)
                        if (reinterpret_cast<const volatile int*>(S.ready)[1] != 0 || clock64() - t_start > 600000000LL) {
                            ok = 0;
                            atomicExch(const_cast<int*>(S.ready) + 1, 1);
                            break;
                        }
                    }
                    s_rows_ok = ok;
                    __threadfence();     // acquire (the row counter was observed by this thread)
                }
                __syncthreads();
                __threadfence();
                arrived = s_rows_ok != 0;
                if (!arrived) { status |= 32; alive = false; }
                __syncthreads();
            }
            const long long nel = nrow * rowlen;
            const double* src = S.x + tile_base * rowlen;
            if (arrived)
                for (long long i = tid; i < nel; i += TTVB_TPB) xstage[i] = __ldcg(src + i);
        }
        __syncthreads();
        // ---- prologue: bounds, cube->physical, constants, elements -> Jacobi state
        if (alive) {
            const double* xr = xstage + (size_t)slot_in_tile * rowlen;
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
            const double* xr = xstage + (size_t)slot_in_tile * rowlen;
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
        // The input staging area ALIASES the history ring: no warp may start stepping (and store into
        // the ring) before every warp of the CTA has read its rows. Only timing kept them apart before:
        // a warp would have had to fall ~10^4 cycles behind in its prologue, which nothing rules out.
        __syncthreads();
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
        const bool integrate = mine && !(status & 1);   // out of bounds: chi2=+inf
        // ---- integration of steps k0..k1 of this unit: the step loop is a leaf function with its
        //      own register allocation; it comes back whenever the warp's event queue is full
        const long long k0 = (long long)seg * S.seg_len + 1;
        const long long k1 = (seg == S.nseg - 1) ? (S.nsteps + 1) : (long long)(seg + 1) * S.seg_len;
        const int warp_base_tid = warp * 32;
        const long long warp_base_sys = tile_base + warp * S.lanes;
        StepCtx<N> cx;
        cx.p = p;
#pragma unroll
        for (int i = 0; i < N; i++) { cx.prev_dot[i] = prev_dot[i]; cx.count[i] = count[i]; }
        cx.trigmask = trigmask; cx.nev = nev; cx.status = status; cx.alive = alive ? 1 : 0;
        cx.Time = Time; cx.Tm1 = Tm1; cx.Tm2 = Tm2;
        cx.qcount = 0; cx.k = k0; cx.pending_push = 0;
        while (run_steps<N>(cx, S.dt, S.nsteps, k1, o_cst, Q.qi, Q.qd)) {
            process_batch<N, COMPACT>(S, TTVB_QCAP, Q, kc_s, q_s, gm_s, res_i, res_d, warp_base_tid, warp_base_sys, obs_epoch_s, obs_time_s, obs_sigma_s);
            consume_results(lane, tid, res_i, res_d, acc_s, cur_s, cx.status);
            cx.qcount = 0;
        }
        // ---- drain the queue
        {
            const int cnt = cx.qcount;   // warp-uniform
            if (cnt > 0) {
                process_batch<N, COMPACT>(S, cnt, Q, kc_s, q_s, gm_s, res_i, res_d, warp_base_tid, warp_base_sys, obs_epoch_s, obs_time_s, obs_sigma_s);
                consume_results(lane, tid, res_i, res_d, acc_s, cur_s, cx.status);
            }
        }
        p = cx.p;
#pragma unroll
        for (int i = 0; i < N; i++) { prev_dot[i] = cx.prev_dot[i]; count[i] = cx.count[i]; }
        trigmask = cx.trigmask; nev = cx.nev; status = cx.status; alive = cx.alive != 0;
        Time = cx.Time; Tm1 = cx.Tm1; Tm2 = cx.Tm2;
        if (seg < S.nseg - 1) {
            // ---- hand the tile over: save the context, publish the segment
            if (mine) {
                double* cd = S.ctx_d + sys;
                int* ci = S.ctx_i + sys;
                const long long cs = S.ctx_stride;
                int f = 0;
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { __stcs(cd + (f++) * cs, p.x[i][c]); }
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { __stcs(cd + (f++) * cs, p.v[i][c]); }
                for (int j = 0; j < 12 * N; j++) __stcs(cd + (f++) * cs, hist[(size_t)j * TTVB_RS + tid]);
#pragma unroll
                for (int i = 0; i < N; i++) __stcs(cd + (f++) * cs, prev_dot[i]);
                for (int i = 0; i < N; i++) __stcs(cd + (f++) * cs, acc_s[i * TTVB_TPB + tid]);
                for (int j = 0; j < 5 * N; j++) __stcs(cd + (f++) * cs, cst_s[(size_t)j * TTVB_TPB + tid]);
                __stcs(cd + (f++) * cs, Time); __stcs(cd + (f++) * cs, Tm1); __stcs(cd + (f++) * cs, Tm2);
                int g = 0;
#pragma unroll
                for (int i = 0; i < N; i++) __stcs(ci + (g++) * cs, count[i]);
                for (int i = 0; i < N; i++) __stcs(ci + (g++) * cs, cur_s[i * TTVB_TPB + tid]);
                __stcs(ci + (g++) * cs, trigmask);
                __stcs(ci + (g++) * cs, nev);
                __stcs(ci + (g++) * cs, status);
                __stcs(ci + (g++) * cs, alive ? 1 : 0);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(S.sched + 1 + tile, seg + 1);
            continue;
        }
        // ---- epilogue: remaining observed epochs are missing -> penalty; totals
        if (mine) {
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

// Seed self-test: the table lookup of ttvb_math.cuh (what the CPU oracle uses) against the
// special-function unit, for every high operand word. A NaN on both sides counts as equal.
__global__ void ttvb_seed_scan_kernel(const uint32_t* rcp_tab, const uint32_t* rsq_tab, unsigned long long* bad,
                                      uint32_t* first_bad) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < (1ull << 32); w += stride) {
        const uint32_t hi = (uint32_t)w;
        const double x = __hiloint2double((int)hi, (int)(hi * 2654435761u));
        const uint32_t hr = (uint32_t)__double2hiint(rcp_seed(x)), hs = (uint32_t)__double2hiint(rsq_seed(x));
        const uint32_t er = seed_rcp_emul(hi, rcp_tab), es = seed_rsq_emul(hi, rsq_tab);
        auto is_nan = [](uint32_t h) { return (h & 0x7FF00000u) == 0x7FF00000u && (h & 0xFFFFFu) != 0u; };
        if (hr != er && !(is_nan(hr) && is_nan(er))) {
            if (atomicAdd(bad, 1ull) == 0ull) { first_bad[0] = hi; first_bad[1] = hr; first_bad[2] = er; }
        }
        if (hs != es && !(is_nan(hs) && is_nan(es))) {
            if (atomicAdd(bad + 1, 1ull) == 0ull) { first_bad[3] = hi; first_bad[4] = hs; first_bad[5] = es; }
        }
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
