/*
 * ttv_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, scalar, one-system-at-a-time restatement of nauyaca's hot path
 *   flat parameter vector -> mid-transit ephemerides -> chi-square -> log-likelihood
 * i.e. of  /root/reference/nauyaca/utils.py:16-447,1193-1219  plus the N-body engine
 * that file calls through `ttvfast.ttvfast(...)` (utils.py:72-74).
 *
 * The engine is the third-party dependency `ttvfast==0.3.0`
 * (reference pyproject.toml:10, poetry.lock:906-907; upstream C = kdeck/TTVFast
 * c_version, Deck, Agol, Holman & Nesvorny 2014, ApJ 787, 132). It is NOT vendored in
 * /root/reference and not installed here, so this file restates its PUBLISHED
 * ALGORITHM: Wisdom-Holman map in Jacobi coordinates, Wisdom's
 * third-order symplectic corrector applied once, universal-variable Kepler drift,
 * transit search on sky-plane r.v sign change, two-sided Keplerian Newton refinement
 * with linear closeness weighting.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the full-precision
 * ephemerides and chi2 values stored in the reference's own notebooks (produced by the
 * real ttvfast 0.3.0): tests/golden/known_answers.json (G1..G4). What those vectors do
 * NOT pin (rsky/vsky values, >5000 events, Kepler non-convergence, BAD_TRANSIT) is
 * "parity unpinned" and listed in DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product (nauyaca_b200/csrc) never links or calls it.
 *
 * Arithmetic discipline: every floating-point operation below is written out
 * explicitly (fma() where a fused multiply-add is meant, nothing else contracted:
 * build with -ffp-contract=off), and no libm transcendental is used (sin/cos/cbrt are
 * the deterministic polynomial/Newton versions below). The CUDA kernel performs the
 * same IEEE-754 operation sequence per system, so GPU results can be compared to this
 * oracle bit for bit, not merely to a tolerance.
 *
 * Reciprocals and reciprocal square roots of the hot loop ("canonical arithmetic v3"): a
 * 20-bit seed refined by ONE third-order step (rcp_fast / rsqrt_fast below, about 0.6 ulp,
 * not correctly rounded). The seed is a pure function of the operand's high 32-bit word; its
 * 2^20 + 2^21 values are data (oracle/mufu_sm100a.npz, handed over with
 * ttvo_set_seed_tables): they are what the sm_100a special-function unit returns
 * (MUFU.RCP64H / MUFU.RSQ64H, dumped exhaustively by scripts/mufu_dump.cu and re-checked on
 * the GPU by tests/test_gpu_seeds.py). The refined values do not depend on the seed beyond
 * the last bit, so the reference-held goldens pin this oracle exactly as before; the table
 * only makes the last bit the same on both sides.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define TTVO_MAXP 9                 /* TTVFast MAX_N_PLANETS */
#define TTVO_G 0.000295994511       /* AU^3 day^-2 Msun^-1, TTVFast's G */
#define TTVO_MEARTH 3.0034893488507934e-06   /* reference constants.py:65 */
#define TTVO_PENALTY 1e+20          /* reference utils.py:206 */
#define TTVO_BAD_TRANSIT (-1.0)
#define TTVO_KEP_MAXIT 12
#define TTVO_TR_MAXIT 12

/* status bits (same values as include/ttvb.h) */
#define ST_OUT_OF_BOUNDS 1
#define ST_KEPLER_FAIL   2
#define ST_BAD_TRANSIT   4
#define ST_EVENT_OVERFLOW 8
#define ST_INVALID_INPUT 16

/* ------------------------------------------------------------------ */
/* deterministic elementary functions (prologue only)                  */
/* ------------------------------------------------------------------ */

/* sin and cos of r, |r| <= ~pi/4, fdlibm-style minimax kernels */
static void sincos_kernel(double r, double *s, double *c)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = r * r;
    double ps = fma(z, fma(z, fma(z, fma(z, fma(z, S6, S5), S4), S3), S2), S1);
    *s = fma(r * z, ps, r);
    double pc = fma(z, fma(z, fma(z, fma(z, fma(z, C6, C5), C4), C3), C2), C1);
    *c = fma(z, fma(z, pc, -0.5), 1.0);
}

static void quadrant_fix(long long k, double sr, double cr, double *s, double *c)
{
    switch ((int)(k & 3)) {
    case 0: *s = sr;  *c = cr;  break;
    case 1: *s = cr;  *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

/* sin/cos of an angle given in degrees; the reduction to [-45,45] deg is exact */
static void det_sincos_deg(double deg, double *s, double *c)
{
    const double DEG2RAD = 1.74532925199432957692e-02;
    double k = rint(deg * (1.0 / 90.0));
    double r = fma(-90.0, k, deg);
    double sr, cr;
    sincos_kernel(r * DEG2RAD, &sr, &cr);
    quadrant_fix((long long)k, sr, cr, s, c);
}

/* sin/cos of an angle in radians, |x| small (a few pi): two-term Cody-Waite reduction */
static void det_sincos_rad(double x, double *s, double *c)
{
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_1 = 1.57079632673412561417e+00;   /* first 33 bits of pi/2 */
    const double PIO2_1T = 6.07710050650619224932e-11;  /* pi/2 - PIO2_1 */
    double k = rint(x * TWO_OVER_PI);
    double r = fma(-k, PIO2_1, x);
    r = fma(-k, PIO2_1T, r);
    double sr, cr;
    sincos_kernel(r, &sr, &cr);
    quadrant_fix((long long)k, sr, cr, s, c);
}

/* cube root of a positive normal double: bit-trick seed + 4 Halley + 1 Newton */
static double det_cbrt(double y)
{
    uint64_t b;
    memcpy(&b, &y, 8);
    uint32_t hi = (uint32_t)(b >> 32);
    hi = hi / 3u + 715094163u;
    b = ((uint64_t)hi) << 32;
    double t;
    memcpy(&t, &b, 8);
    for (int i = 0; i < 4; i++) {
        double t3 = t * t * t;
        t = t * (fma(2.0, y, t3) / fma(2.0, t3, y));
    }
    double t2 = t * t;
    double res = fma(t2, t, -y);
    t = t - res / (3.0 * t2);
    return t;
}

/* Kepler's equation E - e sin E = M (radians, M in [-pi,pi]); monotone Newton */
static double solve_kepler_E(double M, double e, double *sE, double *cE)
{
    const double PI = 3.14159265358979311600e+00;
    double E;
    if (M >= 0.0) { E = M + e; if (E > PI) E = PI; }
    else          { E = M - e; if (E < -PI) E = -PI; }
    double s, c;
    for (int it = 0; it < 60; it++) {
        det_sincos_rad(E, &s, &c);
        double f = fma(-e, s, E) - M;
        double fp = fma(-e, c, 1.0);
        double d = f / fp;
        E = E - d;
        if (fabs(d) <= 1e-15) break;
    }
    det_sincos_rad(E, sE, cE);
    return E;
}

/* ------------------------------------------------------------------ */
/* system state                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    int n;
    double gm0;              /* G*Mstar */
    double gm[TTVO_MAXP];    /* G*m_i */
    double kc[TTVO_MAXP];    /* Jacobi Kepler constant  G*Mstar*eta_i/eta_{i-1}   (A.3) */
    double q[TTVO_MAXP];     /* G*m_i/eta_i : centre-of-mass recurrence weight */
    double ieta[TTVO_MAXP];  /* 1/eta_{i-1} */
    double s[TTVO_MAXP];     /* G*Mstar/eta_{i-1} */
    double mu[TTVO_MAXP];    /* G*(Mstar+m_i): two-body constant of the transit locator (A.6) */
} consts_t;

typedef struct { double x[TTVO_MAXP][3], v[TTVO_MAXP][3]; } state_t;

static double dot3(const double *a, const double *b)
{
    return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]));
}

/* ------------------------------------------------------------------ */
/* seeds of 1/b and 1/sqrt(a), and their one-step refinements           */
/* ------------------------------------------------------------------ */
static const uint32_t *seed_rcp_tab;   /* [2^20]: high word of seed(1.m), m = 20 mantissa bits */
static const uint32_t *seed_rsq_tab;   /* [2][2^20]: seed(1.m), seed(2 * 1.m) */

void ttvo_set_seed_tables(const uint32_t *rcp, const uint32_t *rsq)
{
    seed_rcp_tab = rcp;
    seed_rsq_tab = rsq;
}

static double from_hi(uint32_t hi)
{
    uint64_t b = ((uint64_t)hi) << 32;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static uint32_t hi_of(double d)
{
    uint64_t b;
    memcpy(&b, &d, 8);
    return (uint32_t)(b >> 32);
}

/* seed of 1/b: only the high word of b is looked at; other exponents rescale the table
 * entry by an exact power of two; zero/denormal operands and results are flushed */
static double seed_rcp(double b)
{
    if (!seed_rcp_tab) abort();
    uint32_t hi = hi_of(b), sign = hi & 0x80000000u, ex = (hi >> 20) & 0x7FFu, m = hi & 0xFFFFFu;
    if (ex == 0) return from_hi(sign | 0x7FF00000u);
    if (ex == 0x7FF) return m ? from_hi(0x7FFFFFFFu) : from_hi(sign);
    uint32_t t = seed_rcp_tab[m];
    int e = (int)(t >> 20) - (int)ex + 1023;          /* biased exponent of the result */
    if (e <= 0) return from_hi(sign);
    return from_hi(sign | ((uint32_t)e << 20) | (t & 0xFFFFFu));
}

/* seed of 1/sqrt(a): a = 2^(2k+p) * 1.m  ->  2^-k * seed(2^p * 1.m) */
static double seed_rsq(double a)
{
    if (!seed_rsq_tab) abort();
    uint32_t hi = hi_of(a), sign = hi & 0x80000000u, ex = (hi >> 20) & 0x7FFu, m = hi & 0xFFFFFu;
    if (ex == 0) return from_hi(sign | 0x7FF00000u);
    if (ex == 0x7FF && m) return from_hi(0x7FFFFFFFu);
    if (sign) return from_hi(0xFFF80000u);
    if (ex == 0x7FF) return 0.0;
    int E = (int)ex - 1023, p = E & 1, k = (E - p) / 2;
    uint32_t t = seed_rsq_tab[((size_t)p << 20) + m];
    int e = (int)(t >> 20) - k;
    return from_hi(((uint32_t)e << 20) | (t & 0xFFFFFu));
}

/* 1/b: cubic refinement of the seed (relative error of the seed below 2^-19) */
#ifdef TTVO_NATIVE_RCP
/* Timing build (libttv_oracle_native.so, bench.py's CPU legs only): the host's own divide and
 * square root instead of the emulated GPU seeds, whose 12 MB of table look-ups cost a CPU four
 * times the arithmetic. Same algorithm, last-bit differences; never used as the checker. */
static double rcp_fast(double b) { return 1.0 / b; }
static double rsqrt_fast(double a) { return 1.0 / sqrt(a); }
void ttvo_unused_seed_refs(void) { (void)seed_rcp; (void)seed_rsq; }
#else
static double rcp_fast(double b)
{
    double y0 = seed_rcp(b);
    double e = fma(-b, y0, 1.0);
    e = fma(e, e, e);
    return fma(y0, e, y0);
}

/* 1/sqrt(a): third-order refinement  y0 (1 + e/2 + 3 e^2/8),  e = 1 - a y0^2 */
static double rsqrt_fast(double a)
{
    double y0 = seed_rsq(a);
    double t = y0 * y0;
    double e = fma(-a, t, 1.0);
    double p = fma(e, 0.375, 0.5);
    double h = y0 * e;
    return fma(p, h, y0);
}
#endif

/* ------------------------------------------------------------------ */
/* Stumpff functions c2,c3 by series with argument quartering (A.9)    */
/* ------------------------------------------------------------------ */
static void stumpff_G(double alpha, double X, double *G0, double *G1, double *G2, double *G3)
{
    const double I2 = 1.0 / 2.0, I24 = 1.0 / 24.0, I720 = 1.0 / 720.0, I40320 = 1.0 / 40320.0,
                 I3628800 = 1.0 / 3628800.0, I479001600 = 1.0 / 479001600.0,
                 I87178291200 = 1.0 / 87178291200.0;
    const double I6 = 1.0 / 6.0, I120 = 1.0 / 120.0, I5040 = 1.0 / 5040.0, I362880 = 1.0 / 362880.0,
                 I39916800 = 1.0 / 39916800.0, I6227020800 = 1.0 / 6227020800.0,
                 I1307674368000 = 1.0 / 1307674368000.0;
    double X2 = X * X;
    double z = alpha * X2;
    int nq = 0;
    while (fabs(z) > 0.125 && nq < 64) { z = z * 0.25; nq++; }
    /* c2 = 1/2! - z/4! + z^2/6! - ... ; c3 = 1/3! - z/5! + z^2/7! - ... */
    double c2 = fma(z, fma(z, fma(z, fma(z, fma(z, fma(z, I87178291200, -I479001600), I3628800),
                                          -I40320), I720), -I24), I2);
    double c3 = fma(z, fma(z, fma(z, fma(z, fma(z, fma(z, I1307674368000, -I6227020800), I39916800),
                                          -I362880), I5040), -I120), I6);
    for (int i = 0; i < nq; i++) {
        double c1 = fma(-z, c3, 1.0);
        double c0 = fma(-z, c2, 1.0);
        c3 = fma(c0, c3, c2) * 0.25;
        c2 = 0.5 * c1 * c1;
        z = z * 4.0;
    }
    *G2 = X2 * c2;
    *G3 = X2 * X * c3;
    *G1 = fma(-alpha, *G3, X);
    *G0 = fma(-alpha, *G2, 1.0);
}

/*
 * Universal-variable Kepler drift of one body about a fixed centre:
 * advance (x,v) by tau under mu. Quintic (Schroeder) Newton on
 *   r0*G1 + u*G2 + mu*G3 = tau,   one reciprocal per iteration, started from a fourth-order
 * series guess so that one Stumpff evaluation normally suffices; the G functions are then
 * moved to X+d by a third-order Taylor step (|d| <= 2^-13 |X|: the neglected term is below
 * 2^-57), and the state is advanced with Gauss f,g in increment form.
 * r0 = |x| and ir0 = 1/|x| are passed in (the kick computes them anyway).
 * Returns 0, or ST_KEPLER_FAIL if not converged.
 */
#define TTVO_KEP_TOL 1.220703125e-4    /* 2^-13 */
static int kepler_drift_r(double mu, double tau, double r0, double ir0, double *x, double *v)
{
    double v2 = dot3(v, v);
    double u = dot3(x, v);
    double m = mu * ir0;
    double alpha = fma(2.0, m, -v2);
    double zeta = fma(-alpha, r0, mu);
    double au = alpha * u;
    /* fourth-order series guess (reversion of tau/r0 = X + A2 X^2 + A3 X^3 + A4 X^4) */
    double y = tau * ir0;
    double s = u * ir0;
    double A2 = 0.5 * s;
    double A3 = (m - alpha) * (1.0 / 6.0);
    double A4 = (au * ir0) * (-1.0 / 24.0);
    double B3 = fma(s, A2, -A3);
    double B4 = fma(5.0 * A2, fma(A2, A2, -A3), A4);
    double X = y * fma(y, fma(y, fma(y, -B4, B3), -A2), 1.0);
    double a24 = alpha * (-1.0 / 24.0);
    double G0, G1, G2, G3, r = 0.0;
    int ok = 0;
    for (int it = 0; it < TTVO_KEP_MAXIT; it++) {
        stumpff_G(alpha, X, &G0, &G1, &G2, &G3);
        double F = fma(r0, G1, fma(u, G2, fma(mu, G3, -tau)));
        double F1 = fma(r0, G0, fma(u, G1, mu * G2));
        double F2 = fma(zeta, G1, u * G0);
        double F3 = fma(-au, G1, zeta * G0);
        double iF1 = rcp_fast(F1);
        double h = -F * iF1;
        double fb = F2 * iF1;
        double b = 0.5 * fb;
        double c = F3 * iF1 * (1.0 / 6.0);
        double e4 = a24 * fb;
        /* quintic Schroeder step: h - b h^2 + (2b^2-c) h^3 - (5b^3-5bc+e4) h^4 */
        double k3 = fma(fb, b, -c);
        double k4 = fma(5.0 * b, fma(b, b, -c), e4);
        double d = h * fma(h, fma(h, fma(h, -k4, k3), -b), 1.0);
        if (fabs(d) <= TTVO_KEP_TOL * fabs(X)) {
            /* converged: third-order Taylor update of the G functions to X+d */
            double d2 = 0.5 * d, d3 = d * (1.0 / 3.0);
            double aG1 = -alpha * G1, aG0 = -alpha * G0;
            double nG3 = fma(d, fma(d2, fma(d3, G0, G1), G2), G3);
            double nG2 = fma(d, fma(d2, fma(d3, aG1, G0), G1), G2);
            double nG1 = fma(d, fma(d2, fma(d3, aG0, aG1), G0), G1);
            G3 = nG3; G2 = nG2; G1 = nG1;
            G0 = fma(-alpha, G2, 1.0);
            ok = 1;
            break;
        }
        X = X + d;
    }
    if (!ok) return ST_KEPLER_FAIL;
    double mg2 = mu * G2;
    r = fma(r0, G0, fma(u, G1, mg2));
    double ir = rcp_fast(r);
    double fm1 = -m * G2;
    double g = fma(-mu, G3, tau);
    double fd = -(m * G1) * ir;
    double gdm1 = -mg2 * ir;
    for (int k = 0; k < 3; k++) {
        double xn = fma(g, v[k], fma(fm1, x[k], x[k]));
        double vn = fma(fd, x[k], fma(gdm1, v[k], v[k]));
        x[k] = xn; v[k] = vn;
    }
    return 0;
}

static int kepler_drift(double mu, double tau, double *x, double *v)
{
    double r0sq = dot3(x, x);
    double ir0 = rsqrt_fast(r0sq);
    double r0 = r0sq * ir0;
    return kepler_drift_r(mu, tau, r0, ir0, x, v);
}

/* A(tau): every planet drifts on its Jacobi Kepler orbit (A.4) */
static int map_A(const consts_t *C, state_t *p, double tau, int upto)
{
    int st = 0;
    for (int i = 0; i <= upto; i++) st |= kepler_drift(C->kc[i], tau, p->x[i], p->v[i]);
    return st;
}

/* 1/|x|^3; optionally also |x| = |x|^2 * (1/|x|) and 1/|x| */
static double inv_r3(const double *x, double *r_out, double *ir_out)
{
    double r2 = dot3(x, x);
    double ir = rsqrt_fast(r2);
    if (r_out) { *r_out = r2 * ir; *ir_out = ir; }
    return ir * ir * ir;
}

/*
 * B(tau) followed by A(tau): kick of the Jacobi velocities by the interaction Hamiltonian (A.4)
 *  a_i = kc_i (r'_i/|r'_i|^3 - r_i/|r_i|^3) - (GMstar/eta_{i-1}) sum_{k>i} Gm_k r_k/|r_k|^3
 *        + D_i - sum_{j<i} (Gm_j/eta_{i-1}) D_j ,   D_i = sum_{k!=i} Gm_k (r_k-r_i)/|r_k-r_i|^3
 * (r astrocentric, r' Jacobi). The Jacobi radii |r'_i| the kick needs are the ones the drift
 * starts from, so map_B hands them over (rj, irj).
 */
static void map_B(const consts_t *C, state_t *p, double tau, double *rj, double *irj)
{
    int n = C->n;
    double r[TTVO_MAXP][3], w[TTVO_MAXP][3], D[TTVO_MAXP][3], R[3];
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            if (i == 0) { r[0][k] = p->x[0][k]; R[k] = C->q[0] * p->x[0][k]; }
            else { r[i][k] = p->x[i][k] + R[k]; R[k] = fma(C->q[i], p->x[i][k], R[k]); }
        }
        double ir3 = (i == 0) ? inv_r3(r[0], &rj[0], &irj[0]) : inv_r3(r[i], NULL, NULL);
        for (int k = 0; k < 3; k++) w[i][k] = r[i][k] * ir3;
    }
    /* D_i: the first contribution is a product, the others are accumulated */
    int touched[TTVO_MAXP] = {0};
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            double d[3] = {r[j][0] - r[i][0], r[j][1] - r[i][1], r[j][2] - r[i][2]};
            double ir3 = inv_r3(d, NULL, NULL);
            double gj = C->gm[j] * ir3, gi = -C->gm[i] * ir3;
            for (int k = 0; k < 3; k++) {
                D[i][k] = touched[i] ? fma(gj, d[k], D[i][k]) : gj * d[k];
                D[j][k] = touched[j] ? fma(gi, d[k], D[j][k]) : gi * d[k];
            }
            touched[i] = touched[j] = 1;
        }
    if (n == 1) return;      /* a single planet feels no kick (its Jacobi orbit is the Kepler orbit) */
    /* T_i = sum_{k>i} Gm_k w_k (suffix), U_i = sum_{j<i} Gm_j D_j (prefix) */
    double T[TTVO_MAXP][3], U[3] = {0.0, 0.0, 0.0};
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = n - 1; i >= 0; i--) {
        for (int k = 0; k < 3; k++) {
            T[i][k] = acc[k];
            acc[k] = (i == n - 1) ? C->gm[i] * w[i][k] : fma(C->gm[i], w[i][k], acc[k]);
        }
    }
    for (int i = 0; i < n; i++) {
        double a[3];
        if (i == 0) {
            for (int k = 0; k < 3; k++) a[k] = fma(-C->s[0], T[0][k], D[0][k]);
        } else {
            double ir3 = inv_r3(p->x[i], &rj[i], &irj[i]);
            for (int k = 0; k < 3; k++) {
                double wj = p->x[i][k] * ir3;
                double t = fma(C->kc[i], wj - w[i][k], D[i][k]);
                if (i < n - 1) t = fma(-C->s[i], T[i][k], t);     /* nothing outside the last planet */
                a[k] = fma(-C->ieta[i], U[k], t);
            }
        }
        for (int k = 0; k < 3; k++) {
            U[k] = (i == 0) ? C->gm[0] * D[0][k] : fma(C->gm[i], D[i][k], U[k]);
            p->v[i][k] = fma(tau, a[k], p->v[i][k]);
        }
    }
}

/* one step of the map: B(dt) then A(dt), the drift reusing the kick's Jacobi radii */
static int map_BA(const consts_t *C, state_t *p, double dt)
{
    double rj[TTVO_MAXP], irj[TTVO_MAXP];
    int st = 0;
    map_B(C, p, dt, rj, irj);
    for (int i = 0; i < C->n; i++) st |= kepler_drift_r(C->kc[i], dt, rj[i], irj[i], p->x[i], p->v[i]);
    return st;
}

/* Jacobi -> astrocentric position and velocity of planet i only */
static void astro_of(const consts_t *C, const state_t *p, int i, double *x, double *v)
{
    double R[3] = {0.0, 0.0, 0.0}, V[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < i; j++)
        for (int k = 0; k < 3; k++) {
            R[k] = fma(C->q[j], p->x[j][k], R[k]);
            V[k] = fma(C->q[j], p->v[j][k], V[k]);
        }
    for (int k = 0; k < 3; k++) { x[k] = p->x[i][k] + R[k]; v[k] = p->v[i][k] + V[k]; }
}

/* third-order corrector pieces (A.5) */
static int corr_C(const consts_t *C, state_t *p, double a, double b)
{
    double rj[TTVO_MAXP], irj[TTVO_MAXP];
    int st = map_A(C, p, -a, C->n - 1);
    map_B(C, p, b, rj, irj);
    st |= map_A(C, p, a, C->n - 1);
    return st;
}
static int corr_Z(const consts_t *C, state_t *p, double a, double b)
{
    int st = corr_C(C, p, -a, -b);
    st |= corr_C(C, p, a, b);
    return st;
}

/* ------------------------------------------------------------------ */
/* prologue: masses + orbital elements -> Jacobi Cartesian state (A.2, A.3) */
/* flat = nauyaca order per planet (utils.py:58-70):                    */
/*   mass[Mearth] period ecc inclination argument mean_anomaly ascending_node */
/* ------------------------------------------------------------------ */
static int prologue(const double *flat, int n, double mstar, consts_t *C, state_t *p)
{
    const double FOUR_PI2 = 3.94784176043574344753e+01;
    const double DEG2RAD = 1.74532925199432957692e-02;
    int st = 0;
    C->n = n;
    C->gm0 = TTVO_G * mstar;
    double eta_prev = C->gm0;
    double R[3] = {0.0, 0.0, 0.0}, V[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; i++) {
        const double *e = flat + 7 * i;
        double mass = e[0] * TTVO_MEARTH;
        double P = e[1], ecc = e[2];
        C->gm[i] = TTVO_G * mass;
        double eta = eta_prev + C->gm[i];
        C->ieta[i] = 1.0 / eta_prev;
        C->s[i] = C->gm0 * C->ieta[i];
        C->kc[i] = C->gm0 * eta * C->ieta[i];
        C->q[i] = C->gm[i] / eta;
        C->mu[i] = C->gm0 + C->gm[i];
        eta_prev = eta;
        /* semi-major axis from the JACOBI Kepler constant, velocities from mu_i (A.2) */
        double a3 = C->kc[i] * P * P / FOUR_PI2;
        if (!(a3 > 0.0) || !(a3 < 1e300) || !(ecc >= 0.0) || !(ecc < 1.0)) st |= ST_INVALID_INPUT;
        double a = (st & ST_INVALID_INPUT) ? 1.0 : det_cbrt(a3);
        double nmot = sqrt(C->mu[i] / (a * a * a));
        double Mdeg = e[5];
        Mdeg = fma(-360.0, rint(Mdeg * (1.0 / 360.0)), Mdeg);
        double sE, cE;
        solve_kepler_E(Mdeg * DEG2RAD, ecc, &sE, &cE);
        double b = sqrt(fma(-ecc, ecc, 1.0));
        double idn = 1.0 / fma(-ecc, cE, 1.0);
        double X = a * (cE - ecc), Y = a * b * sE;
        double an = a * nmot;
        double Xd = -an * sE * idn, Yd = an * b * cE * idn;
        double si, ci, sw, cw, sO, cO;
        det_sincos_deg(e[3], &si, &ci);
        det_sincos_deg(e[4], &sw, &cw);
        det_sincos_deg(e[6], &sO, &cO);
        double Px = fma(-sO * sw, ci, cO * cw), Qx = fma(-sO * cw, ci, -cO * sw);
        double Py = fma(cO * sw, ci, sO * cw),  Qy = fma(cO * cw, ci, -sO * sw);
        double Pz = sw * si,                     Qz = cw * si;
        double xa[3] = {fma(X, Px, Y * Qx), fma(X, Py, Y * Qy), fma(X, Pz, Y * Qz)};
        double va[3] = {fma(Xd, Px, Yd * Qx), fma(Xd, Py, Yd * Qy), fma(Xd, Pz, Yd * Qz)};
        for (int k = 0; k < 3; k++) {
            p->x[i][k] = xa[k] - R[k];
            p->v[i][k] = va[k] - V[k];
            R[k] = fma(C->q[i], p->x[i][k], R[k]);
            V[k] = fma(C->q[i], p->v[i][k], V[k]);
        }
    }
    return st;
}

/* ------------------------------------------------------------------ */
/* transit locator (A.6)                                               */
/* ------------------------------------------------------------------ */
static double sky_g(const double *x, const double *v) { return fma(x[1], v[1], x[0] * v[0]); }

/* dg/dt = vx^2 + vy^2 - mu (x^2+y^2)/r^3 along a two-body orbit; also x^2+y^2 and vx^2+vy^2 */
static double sky_gp(double mu, const double *x, const double *v, double *rs2_out, double *vs2_out)
{
    double rs2 = fma(x[1], x[1], x[0] * x[0]);
    double vs2 = fma(v[1], v[1], v[0] * v[0]);
    double r2 = fma(x[2], x[2], rs2);
    double ir = rsqrt_fast(r2);
    *rs2_out = rs2; *vs2_out = vs2;
    return fma(-mu * rs2, ir * ir * ir, vs2);
}

long ttvo_locate_evals;      /* diagnostics: Kepler drifts spent in locate_side */

/* One step of the root search of g = x vx + y vy along a two-body orbit: Newton's step d = -g/g'
 * and its Halley correction d (1 - d g''/(2 g')), with
 *   g'  = vs^2 - mu rs^2 / r^3 ,   g'' = mu/r^3 ( 3 rs^2 (r.v)/r^2 - 4 g )
 * (rs^2 = x^2+y^2, vs^2 = vx^2+vy^2). Returns the corrected step, *d_newton the plain one. */
static double sky_step(double mu, const double *x, const double *v, double *d_newton, double *rs2_out, double *vs2_out)
{
    double g = sky_g(x, v);
    double u = dot3(x, v);
    double rs2 = fma(x[1], x[1], x[0] * x[0]);
    double vs2 = fma(v[1], v[1], v[0] * v[0]);
    double r2 = fma(x[2], x[2], rs2);
    double ir = rsqrt_fast(r2);
    double ir2 = ir * ir;
    double m3 = mu * (ir2 * ir);
    double gp = fma(-m3, rs2, vs2);
    double igp = rcp_fast(gp);
    double d = -g * igp;
    double gpp = m3 * fma(3.0, rs2 * u * ir2, -4.0 * g);
    double c = gpp * igp * d;
    *d_newton = d; *rs2_out = rs2; *vs2_out = vs2;
    return fma(-0.5 * c, d, d);
}

/* Root of g(tau) = 0 along the two-body orbit through (x0,v0); returns status. The search stops
 * when the NEWTON step is below 1e-4 d: the Halley-corrected iterate it then moves to is within
 * ~1e-13 d of the root (cubic convergence, |d| is 7e-5 d at most after the Hermite first guess
 * at dt = 0.35 d), so one Kepler drift per side is the rule. */
#define TTVO_TR_TOL 1e-4
static int locate_side(double mu, const double *x0, const double *v0, double *tau,
                       double *rsky, double *vsky)
{
    double t = *tau;
    for (int it = 0; it < TTVO_TR_MAXIT; it++) {
        double x[3] = {x0[0], x0[1], x0[2]}, v[3] = {v0[0], v0[1], v0[2]};
        if (kepler_drift(mu, t, x, v)) return ST_BAD_TRANSIT;
        ttvo_locate_evals++;
        double d, rs2, vs2;
        double dh = sky_step(mu, x, v, &d, &rs2, &vs2);
        t = t + dh;
        if (fabs(d) <= TTVO_TR_TOL) {
            *tau = t; *rsky = sqrt(rs2); *vsky = sqrt(vs2);
            return 0;
        }
    }
    return ST_BAD_TRANSIT;
}

typedef struct {
    int cap, n;
    int *planet, *epoch;
    double *time, *rsky, *vsky;
} events_t;

/*
 * One trigger = one recorded event. p3[0..2] are the unsynchronised Jacobi states after
 * steps k-1, k, k+1 and T3 the accumulated times; the trigger fired at step k. Node k is
 * synchronised first; its sky-plane g decides whether the root lies in [T_{k-1},T_k]
 * (TTVFast's direct branch) or in [T_k,T_{k+1}] (its "lost in the unwinding" branch).
 */
static int locate_event(const consts_t *C, const state_t *p3, const double *T3, double dt,
                        int i, double *t_out, double *rsky_out, double *vsky_out)
{
    state_t mid = p3[1];
    double hdt = 0.5 * dt;
    if (map_A(C, &mid, -hdt, i)) return ST_BAD_TRANSIT;
    double xm[3], vm[3];
    astro_of(C, &mid, i, xm, vm);
    double gm = sky_g(xm, vm);
    int other = (gm > 0.0) ? 0 : 2;
    state_t oth = p3[other];
    if (map_A(C, &oth, -hdt, i)) return ST_BAD_TRANSIT;
    double xo[3], vo[3];
    astro_of(C, &oth, i, xo, vo);
    double go = sky_g(xo, vo);
    const double *xb, *vb, *xa, *va;
    double gb, ga, Tb, Ta;
    if (other == 0) { xb = xo; vb = vo; gb = go; Tb = T3[0]; xa = xm; va = vm; ga = gm; Ta = T3[1]; }
    else            { xb = xm; vb = vm; gb = gm; Tb = T3[1]; xa = xo; va = vo; ga = go; Ta = T3[2]; }
    /* first guess of the root: INVERSE cubic Hermite interpolation of t(g) between the two nodes
     * (values 0 and dt, slopes 1/g' at the nodes) evaluated at g = 0; s is the secant fraction.
     * Outside [0, dt] (or NaN) it falls back to the secant point. */
    double den = ga - gb;
    double s = -gb * rcp_fast(den);
    double u1, u2;
    double qb = den * rcp_fast(sky_gp(C->mu[i], xb, vb, &u1, &u2));
    double qa = den * rcp_fast(sky_gp(C->mu[i], xa, va, &u1, &u2));
    double oms = 1.0 - s, s2 = s * s;
    double h10 = s * oms * oms, h01 = s2 * fma(-2.0, s, 3.0), h11 = -s2 * oms;
    double tau_b = fma(h10, qb, fma(h01, dt, h11 * qa));
    if (!(tau_b >= 0.0 && tau_b <= dt)) tau_b = s * dt;
    double tau_a = tau_b - dt;
    double rb, vbk, ra, vak;
    int st = locate_side(C->mu[i], xb, vb, &tau_b, &rb, &vbk);
    st |= locate_side(C->mu[i], xa, va, &tau_a, &ra, &vak);
    if (st) return ST_BAD_TRANSIT;
    double tb = Tb + tau_b, ta = Ta + tau_a;
    double wa = tau_b / (tau_b - tau_a);
    *t_out = fma(wa, ta - tb, tb);
    *rsky_out = fma(wa, ra - rb, rb);
    *vsky_out = fma(wa, vak - vbk, vbk);
    return 0;
}

/* number of iterations of `Time=t0; while (Time<total) Time+=dt` (A.7) */
long ttvo_nsteps(double t0, double total, double dt)
{
    long n = 0;
    double T = t0;
    if (!(dt > 0.0)) return -1;
    while (T < total) { T += dt; n++; if (n > 2000000000L) return -1; }
    return n;
}

/*
 * Engine: flat physical parameters (7 per planet, nauyaca order, degrees, Mearth) ->
 * event list (planet, epoch, time, rsky, vsky), i.e. what `ttvfast.ttvfast(...,
 * input_flag=1)['positions']` returns before the -2 padding (utils.py:72-79).
 * Events are listed per trigger step, planets ascending within a step.
 */
int ttvo_ephemeris(const double *flat, int npla, double mstar, double t0, double dt, double total,
                   int cap, int *n_events, int *planet, int *epoch, double *time, double *rsky,
                   double *vsky)
{
    consts_t C;
    state_t p, hist[3];
    double Thist[3];
    int status = 0, count[TTVO_MAXP], trig[TTVO_MAXP];
    double prev_dot[TTVO_MAXP];
    *n_events = 0;
    if (npla < 1 || npla > TTVO_MAXP) return ST_INVALID_INPUT;
    long nsteps = ttvo_nsteps(t0, total, dt);
    if (nsteps < 0) return ST_INVALID_INPUT;
    status |= prologue(flat, npla, mstar, &C, &p);
    if (status & ST_INVALID_INPUT) return status;
    /* corrector: real -> map, applied once (A.5) */
    const double alpha = 4.18330013267037774e-01;        /* sqrt(7/40) */
    const double beta = 4.98011920555997350e-02;         /* 1/(48 alpha) */
    double ca = alpha * dt, cb = 0.5 * beta * dt;
    int cst = corr_Z(&C, &p, ca, cb);
    cst |= corr_Z(&C, &p, -ca, -cb);
    if (cst) return status | ST_KEPLER_FAIL;
    for (int i = 0; i < npla; i++) {
        prev_dot[i] = sky_g(p.x[i], p.v[i]);
        count[i] = 0; trig[i] = 0;
    }
    if (map_A(&C, &p, 0.5 * dt, npla - 1)) return status | ST_KEPLER_FAIL;
    double Time = t0;
    hist[1] = p; Thist[1] = Time;     /* becomes "k-1" after the first shift */
    hist[2] = p; Thist[2] = Time;
    /* steps 1..nsteps, plus one ghost step so that a trigger at the last step can still
       use the bracket [T_n, T_n+dt] (TTVFast's else-branch advances one extra step) */
    for (long k = 1; k <= nsteps + 1; k++) {
        hist[0] = hist[1]; Thist[0] = Thist[1];
        hist[1] = hist[2]; Thist[1] = Thist[2];
        if (map_BA(&C, &p, dt)) { status |= ST_KEPLER_FAIL; break; }
        Time += dt;
        hist[2] = p; Thist[2] = Time;
        /* events triggered at step k-1 now have their three states */
        for (int i = 0; i < npla; i++) {
            if (!trig[i]) continue;
            trig[i] = 0;
            double t, rs, vs;
            if (locate_event(&C, hist, Thist, dt, i, &t, &rs, &vs)) {
                status |= ST_BAD_TRANSIT;
                t = TTVO_BAD_TRANSIT; rs = TTVO_BAD_TRANSIT; vs = TTVO_BAD_TRANSIT;
            }
            if (*n_events < cap) {
                int e = *n_events;
                planet[e] = i; epoch[e] = count[i]; time[e] = t; rsky[e] = rs; vsky[e] = vs;
                *n_events = e + 1;
            } else status |= ST_EVENT_OVERFLOW;
            count[i]++;
        }
        if (k <= nsteps)
            for (int i = 0; i < npla; i++) {
                double cd = sky_g(p.x[i], p.v[i]);
                if (prev_dot[i] < 0.0 && cd > 0.0 && p.x[i][2] > 0.0) trig[i] = 1;
                prev_dot[i] = cd;
            }
    }
    return status;
}

/* ------------------------------------------------------------------ */
/* host plumbing of utils.py restated                                   */
/* ------------------------------------------------------------------ */

/* utils.py:402 -- f = lambda x: x-360 if x>360 else (360+x if x<0 else x) */
static double wrap360(double x) { return x > 360.0 ? x - 360.0 : (x < 0.0 ? 360.0 + x : x); }

/* utils.py:1193-1219 intervals(): 1 if every lo<=x<=hi */
int ttvo_inside(const double *x, const double *lo, const double *hi, int ndim)
{
    for (int i = 0; i < ndim; i++) if (!(lo[i] <= x[i] && x[i] <= hi[i])) return 0;
    return 1;
}

/* utils.py:440-447 _insert_constants(): const_idx ascending flat indices */
void ttvo_insert_constants(const double *x, int ndim, const int *const_idx, const double *const_val,
                           int nconst, double *flat)
{
    int total = ndim + nconst, ic = 0, ix = 0;
    for (int k = 0; k < total; k++) {
        if (ic < nconst && const_idx[ic] == k) flat[k] = const_val[ic++];
        else flat[k] = x[ix++];
    }
}

/* utils.py:403-438 cube_to_physical() */
void ttvo_cube_to_physical(const double *cube, int ndim, const double *bi, const double *bf,
                           const int *const_idx, const double *const_val, int nconst, int npla,
                           double *flat)
{
    double tmp[7 * TTVO_MAXP];
    for (int i = 0; i < ndim; i++) tmp[i] = bi[i] + cube[i] * (bf[i] - bi[i]);
    ttvo_insert_constants(tmp, ndim, const_idx, const_val, nconst, flat);
    for (int p = 0; p < npla; p++) {
        double x4 = flat[7 * p + 4], x5 = flat[7 * p + 5];
        double w = (x4 + x5) / 2.0;
        double M = w - x5;
        flat[7 * p + 4] = wrap360(w);
        flat[7 * p + 5] = wrap360(M);
    }
}

/*
 * utils.py:84-123 _ephemeris() + utils.py:161-216 _chi2(): keep events with
 * rsky <= rstarAU, index by (planet, epoch); for every observed (planet, epoch) in table
 * order add ((t_obs - t_sim)/sigma)^2 or 1e20 if absent; per-planet partial sums added
 * in planet order. obs_planet[] must be grouped by planet in the order of
 * PSystem.transit_times; chi2_planet (nullable) receives the per-planet sums in order of
 * first appearance.
 */
double ttvo_chi2(int n_events, const int *planet, const int *epoch, const double *time,
                 const double *rsky, double rstar_au, int nobs, const int *obs_planet,
                 const int *obs_epoch, const double *obs_time, const double *obs_sigma,
                 double *chi2_planet)
{
    double tot = 0.0;
    int o = 0, np = 0;
    while (o < nobs) {
        int pl = obs_planet[o];
        double c = 0.0;
        for (; o < nobs && obs_planet[o] == pl; o++) {
            int found = 0;
            double ts = 0.0;
            /* dict(zip(epoch,time)): a later duplicate key would overwrite; keys are unique */
            for (int e = 0; e < n_events; e++)
                if (planet[e] == pl && epoch[e] == obs_epoch[o] && rsky[e] <= rstar_au) {
                    ts = time[e]; found = 1;
                }
            if (found) { double d = (obs_time[o] - ts) / obs_sigma[o]; c += d * d; }
            else c += TTVO_PENALTY;
        }
        if (chi2_planet) chi2_planet[np] = c;
        np++;
        tot += c;
    }
    return tot;
}

typedef struct {
    int npla, ndim, nconst, nobs;
    double mstar, t0, ftime, dt, rstar_au, second_term;
    const double *bi, *bf, *lo, *hi;          /* [ndim] */
    const int *const_idx; const double *const_val;
    const int *obs_planet, *obs_epoch; const double *obs_time, *obs_sigma;
} ttvo_system;

/*
 * utils.py:333-399 log_likelihood_func / utils.py:219-257 calculate_chi2 for one vector.
 * mode 0: x is a hypercube vector (bounds lo/hi = PSystem.hypercube) -> cube_to_physical
 * mode 1: x is physical WITHOUT constants (bounds lo/hi = PSystem.bounds) -> insert
 * mode 2: x is the full 7*npla physical vector, no bounds check (calculate_chi2_physical)
 * Out of bounds: chi2=+inf, logl=-inf, status ST_OUT_OF_BOUNDS.
 */
int ttvo_loglike(const ttvo_system *S, const double *x, int mode, double *chi2_out, double *logl_out)
{
    enum { CAP = 5000 };      /* ttvfast-python allocates 5000 event slots (utils.py:76-79) */
    double flat[7 * TTVO_MAXP];
    if (mode == 0 || mode == 1) {
        if (!ttvo_inside(x, S->lo, S->hi, S->ndim)) {
            if (chi2_out) *chi2_out = INFINITY;
            if (logl_out) *logl_out = -INFINITY;
            return ST_OUT_OF_BOUNDS;
        }
        if (mode == 0)
            ttvo_cube_to_physical(x, S->ndim, S->bi, S->bf, S->const_idx, S->const_val, S->nconst,
                                  S->npla, flat);
        else
            ttvo_insert_constants(x, S->ndim, S->const_idx, S->const_val, S->nconst, flat);
    } else {
        memcpy(flat, x, sizeof(double) * 7 * S->npla);
    }
    int *pl = (int *)malloc(sizeof(int) * 2 * CAP);
    double *tm = (double *)malloc(sizeof(double) * 3 * CAP);
    int n = 0;
    int st = ttvo_ephemeris(flat, S->npla, S->mstar, S->t0, S->dt, S->ftime, CAP, &n, pl, pl + CAP,
                            tm, tm + CAP, tm + 2 * CAP);
    double chi2 = ttvo_chi2(n, pl, pl + CAP, tm, tm + CAP, S->rstar_au, S->nobs, S->obs_planet,
                            S->obs_epoch, S->obs_time, S->obs_sigma, NULL);
    free(pl); free(tm);
    if (chi2_out) *chi2_out = chi2;
    if (logl_out) *logl_out = -0.5 * chi2 - S->second_term;
    return st;
}

/* batch driver used by bench.py's cpu_baseline leg ("port", 1 thread per call) */
int ttvo_loglike_batch(const ttvo_system *S, const double *x, long B, int mode, double *chi2_out,
                       double *logl_out, int *status_out)
{
    for (long b = 0; b < B; b++) {
        double c, l;
        int st = ttvo_loglike(S, x + b * (mode == 2 ? 7 * S->npla : S->ndim), mode, &c, &l);
        if (chi2_out) chi2_out[b] = c;
        if (logl_out) logl_out[b] = l;
        if (status_out) status_out[b] = st;
    }
    return 0;
}

/* exposed for unit tests of the deterministic elementary functions */
void ttvo_test_sincos_deg(double d, double *s, double *c) { det_sincos_deg(d, s, c); }
void ttvo_test_sincos_rad(double x, double *s, double *c) { det_sincos_rad(x, s, c); }
double ttvo_test_cbrt(double y) { return det_cbrt(y); }
int ttvo_test_kepler_drift(double mu, double tau, double *x, double *v) { return kepler_drift(mu, tau, x, v); }
