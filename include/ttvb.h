/*
 * ttvb.h -- C ABI of libttvb.so, the B200 (sm_100a) batched TTV likelihood engine.
 *
 * Drop-in boundary for ONE path of EliabCanul/nauyaca: flat planetary parameter vector ->
 * mid-transit ephemerides -> chi-square -> log-likelihood, evaluated for a whole batch of
 * parameter vectors per call. Plain pointers and sizes only; no C++/torch types.
 * Every entry point names the reference interface it replaces (paths are relative to the
 * reference checkout, nauyaca/...).
 *
 * Conventions
 *  - all functions return 0 on success, <0 on error (TTVB_E_*); the message of the last
 *    error of the calling thread is available from ttvb_last_error(). Nothing throws or
 *    exits across this boundary.
 *  - a handle is bound to one CUDA device; calls on one handle are ordered on the stream
 *    given to the call (NULL = the legacy default stream). One handle per host thread. The
 *    handle owns scratch state shared by all its launches, so launches of ONE handle never
 *    overlap: a call given a different stream than the previous one first waits (on the device,
 *    cudaStreamWaitEvent) for that previous launch. Use one handle per stream for concurrency.
 *  - `x`, and every output array, may be HOST or DEVICE pointers (detected with
 *    cudaPointerGetAttributes). With host buffers the call copies in chunks (cudaMemcpyAsync
 *    from/to the caller's buffers; pin them for full PCIe speed) through device staging owned
 *    by the handle and returns after the results are in the host arrays. With device buffers
 *    the call is asynchronous on `stream`.
 *  - there is NO CPU fallback: without a CUDA device ttvb_create fails with
 *    TTVB_E_NO_DEVICE.
 */
#ifndef TTVB_H
#define TTVB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTVB_MAX_PLANETS 9        /* TTVFast MAX_N_PLANETS */
#define TTVB_EVENT_CAP 5000       /* slots ttvfast-python allocates (utils.py:76-79) */

/* error codes */
#define TTVB_OK 0
#define TTVB_E_INVALID (-1)       /* bad argument */
#define TTVB_E_NO_DEVICE (-2)     /* no CUDA device / device index out of range */
#define TTVB_E_CUDA (-3)          /* CUDA runtime error, see ttvb_last_error() */
#define TTVB_E_UNSUPPORTED (-4)   /* npla outside 1..TTVB_MAX_PLANETS */

/* per-system status bits written to status_out */
#define TTVB_ST_OUT_OF_BOUNDS 1   /* outside hypercube/bounds: chi2=+inf, logl=-inf (utils.py:245-246,368-369) */
#define TTVB_ST_KEPLER_FAIL 2     /* Kepler drift did not converge; integration stopped, remaining epochs penalised */
#define TTVB_ST_BAD_TRANSIT 4     /* a transit refinement failed; its time is -1 (TTVFast BAD_TRANSIT) */
#define TTVB_ST_EVENT_OVERFLOW 8  /* more events than `cap` in ttvb_ephemeris */
#define TTVB_ST_INVALID_INPUT 16  /* non-finite / non-elliptic elements */
#define TTVB_ST_INPUT_TIMEOUT 32  /* host-pointer call: the rows of this system never arrived on the device (a
                                   * failed upload; the call itself returns an error). Never set otherwise. */

/* interpretation of the rows of `x` */
#define TTVB_MODE_CUBE 0          /* x[B][ndim] in the unit hypercube: intervals(PS.hypercube) then
                                     cube_to_physical (utils.py:245-251, 366-374, 403-438) */
#define TTVB_MODE_PHYSICAL 1      /* x[B][ndim] physical, constants NOT included: intervals(PS.bounds)
                                     then _insert_constants (utils.py:377-387, 440-447) */
#define TTVB_MODE_PHYSICAL_FULL 2 /* x[B][7*npla] physical incl. constants, no bounds check
                                     (calculate_chi2_physical, utils.py:260-330; run_TTVFast :16-81) */

/*
 * Per-system constants = the attributes of nauyaca.PlanetarySystem consumed by the path
 * (planetarysystem.py:184-350). All arrays are HOST pointers, copied by ttvb_create.
 */
typedef struct ttvb_system {
    int32_t npla;              /* PS.npla */
    int32_t ndim;              /* PS.ndim  (= 7*npla - nconst) */
    int32_t nconst;            /* len(PS.constant_params) */
    int32_t nobs;              /* total observed transits over all planets */
    double mstar;              /* PS.mstar [Msun] */
    double t0, ftime, dt;      /* PS.t0, PS.ftime, PS.dt [days] (planetarysystem.py:266-303) */
    double rstar_au;           /* PS.rstarAU (planetarysystem.py:204) */
    double second_term;        /* sum(PS.second_term_logL.values()) (planetarysystem.py:345-348) */
    const double *bi, *bf;     /* [ndim] PS.bi, PS.bf (planetarysystem.py:213) */
    const double *cube_lo, *cube_hi;  /* [ndim] PS.hypercube (planetarysystem.py:209) */
    const double *phys_lo, *phys_hi;  /* [ndim] PS.bounds */
    const int32_t *const_idx;  /* [nconst] keys of PS.constant_params, ascending flat indices */
    const double *const_val;   /* [nconst] values of PS.constant_params */
    const int32_t *obs_planet; /* [nobs] planet index (PS.planets_IDs), grouped by planet */
    const int32_t *obs_epoch;  /* [nobs] epoch keys of PS.transit_times[planet], table order */
    const double *obs_time;    /* [nobs] PS.transit_times[planet][epoch] */
    const double *obs_sigma;   /* [nobs] PS.sigma_obs[planet][epoch] (planetarysystem.py:325) */
} ttvb_system;

typedef struct ttvb_handle ttvb_handle;

/* Message of the last error raised on the calling thread ("" if none). */
const char *ttvb_last_error(void);

/* Library/ABI version (major*1000 + minor). */
int ttvb_version(void);

/* Number of CUDA devices visible (0 if none / no driver). */
int ttvb_device_count(void);

/*
 * Upload the constants of one planetary system to `device` and allocate the handle's
 * workspaces. Replaces the per-call re-pickling of PSystem into the multiprocessing
 * workers (optimizers.py:185-189, mcmc.py:193-211). TTVB_E_INVALID for inconsistent sizes,
 * dt <= 0, non-finite times, mstar <= 0, sigma <= 0 or duplicate observed epochs;
 * TTVB_E_UNSUPPORTED for more than TTVB_MAX_PLANETS planets; TTVB_E_NO_DEVICE without a GPU.
 */
int ttvb_create(const ttvb_system *sys, int device, ttvb_handle **out);
int ttvb_destroy(ttvb_handle *h);

/*
 * Replace the hypercube bounds used by TTVB_MODE_CUBE. The reference's optimizer shrinks
 * PSystem.hypercube in place between its global and local stages (optimizers.py:90) while
 * bi/bf stay fixed; this keeps one handle valid across that mutation. lo, hi: [ndim] host.
 */
int ttvb_set_cube_bounds(ttvb_handle *h, const double *lo, const double *hi);

/* Number of integrator steps = iterations of `Time=t0; while(Time<total) Time+=dt` in the
 * engine behind ttvfast.ttvfast (utils.py:72-74). */
int ttvb_nsteps(const ttvb_handle *h, int64_t *nsteps);

/*
 * Batched objective. Replaces, for B vectors at once,
 *   log_likelihood_func(x, PS, cube, insert_constants)   utils.py:333-399  -> logl_out
 *   calculate_chi2(x, PS)                                 utils.py:219-257  -> chi2_out
 *   calculate_chi2_physical(x, PS, individual=...)        utils.py:260-330  -> chi2_out, chi2_planet_out
 * x: [B][ndim] (modes 0,1) or [B][7*npla] (mode 2), row-major doubles.
 * chi2_out[B], logl_out[B], status_out[B] : any may be NULL.
 * chi2_planet_out: NULL or [B][npla] per-planet chi2 (`individual=True`; planets without
 * observations get 0).
 */
int ttvb_loglike(ttvb_handle *h, const double *x, int64_t B, int mode, double *chi2_out,
                 double *logl_out, double *chi2_planet_out, int32_t *status_out, void *stream);

/*
 * Batched engine seam. Replaces run_TTVFast(flat, mstar, t0, ftime, dt) (utils.py:16-81),
 * i.e. ttvfast.ttvfast(..., input_flag=1)['positions'] before the -2 padding, for B vectors.
 * Outputs are [B][cap] row-major; n_events[B] holds the number of valid leading entries.
 * Events are ordered by trigger step, planets ascending within a step; epoch counts every
 * conjunction of that planet (also those with rsky > Rstar).
 */
int ttvb_ephemeris(ttvb_handle *h, const double *x, int64_t B, int mode, int32_t cap,
                   int32_t *n_events, int32_t *planet, int32_t *epoch, double *time, double *rsky,
                   double *vsky, int32_t *status_out, void *stream);

/*
 * cube_to_physical(PS, x) (utils.py:403-438) for B vectors: flat_out[B][7*npla].
 * inside_out (nullable) [B]: 1 if the row is inside PS.hypercube (utils.py:1193-1219).
 */
int ttvb_cube_to_physical(ttvb_handle *h, const double *x, int64_t B, double *flat_out,
                          int32_t *inside_out, void *stream);

/* Device time of the likelihood/ephemeris kernel of the LAST call on this handle, measured
 * with CUDA events on the call's stream (synchronises on the end event). */
int ttvb_last_kernel_ms(ttvb_handle *h, float *ms);

/* Number of kernels launched by this handle since creation (bench.py "gpu_launches"). */
int ttvb_launch_count(const ttvb_handle *h, int64_t *count);

/* Host-pointer calls of two waves or more are STREAMED: the kernel starts at once and its tiles wait,
 * on the device, for rows that the copy stream uploads meanwhile. Where uploads cannot run under a
 * kernel (a profiler serialising the GPU, CUDA_LAUNCH_BLOCKING=1) the tiles give up after a bounded
 * wait (~0.3 s), the launch is run again on the rows that have arrived by then (same results) and
 * streaming is switched off for the handle. `count` = launches that were run again. TTVB_NO_STREAM=1
 * in the environment switches streaming off from the start. */
int ttvb_stream_retries(const ttvb_handle *h, int64_t *count);

/* FP64 FMA-chain microbenchmark on the handle's device: achieved TFLOP/s (2 flop per FMA).
 * Used as the measured FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
int ttvb_fp64_peak(ttvb_handle *h, double *tflops, float *ms);

/* ------------------------------------------------------------------------------------------
 * Parallel-tempering ensemble sampler resident on the device.
 * Replaces, for the reference's default flat prior (MCMC.logprior == 0.0, mcmc.py:505-509), the
 * iteration of ptemcee.Sampler.sample() that mcmc.py:195-221 drives: two half-ensemble stretch
 * moves, temperature swaps and the Vousden ladder adaptation, with the likelihood of every
 * half-step evaluated by the same kernel as ttvb_loglike. Walkers, log-likelihoods and the
 * ladder stay on the GPU between iterations.
 */
typedef struct ttvb_pt ttvb_pt;

/* betas: [ntemps] host, descending from 1 (ptemcee's ladder). a: stretch scale (mcmc.py:190: 10).
 * adaptation_lag / adaptation_time: mcmc.py:188-189 (1000, 100). mode: TTVB_MODE_CUBE or
 * TTVB_MODE_PHYSICAL (walkers without the constants), as mcmc.py:129-137 decides. */
int ttvb_pt_create(ttvb_handle *h, int32_t ntemps, int32_t nwalkers, const double *betas, double a,
                   double adaptation_lag, double adaptation_time, uint64_t seed, int mode, ttvb_pt **out);
int ttvb_pt_destroy(ttvb_pt *pt);
/* Upload walkers p0[ntemps][nwalkers][ndim] (host) and evaluate their log-likelihoods. */
int ttvb_pt_set_walkers(ttvb_pt *pt, const double *p0, void *stream);
/* Run `iterations` PT iterations (asynchronous on `stream`). adapt != 0: adapt the ladder. */
int ttvb_pt_run(ttvb_pt *pt, int64_t iterations, int adapt, void *stream);
/* The same iteration in pieces, for ensembles whose likelihood evaluations are sharded over
 * several GPUs: every rank holds the whole ensemble (same seed, same p0; the
 * moves are deterministic functions of seed and iteration), evaluates the proposals [lo, hi)
 * of the half-step's block of ntemps*nwalkers/2, the caller all-gathers the block of proposal
 * log-likelihoods (ttvb_pt_proposal_logl: device pointer, length, and the capacity available
 * for a padded in-place gather) over NCCL, and every rank accepts / swaps / adapts alike.
 * One iteration = begin(0) gather end(0) begin(1) gather end(1). */
int ttvb_pt_half_begin(ttvb_pt *pt, int half, int64_t lo, int64_t hi, void *stream);
int ttvb_pt_half_end(ttvb_pt *pt, int half, int adapt, void *stream);
int ttvb_pt_proposal_logl(ttvb_pt *pt, double **dev_ptr, int64_t *n, int64_t *capacity);
/* Copy the state to host arrays (any may be NULL); synchronises `stream`.
 * p[ntemps][nwalkers][ndim], logl[ntemps][nwalkers], betas[ntemps], nprop / nprop_accepted
 * [ntemps][nwalkers], nswap / nswap_accepted [ntemps] (ptemcee's counters). */
int ttvb_pt_get(ttvb_pt *pt, double *p, double *logl, double *betas, double *nprop, double *nprop_accepted,
                double *nswap, double *nswap_accepted, void *stream);

/* Self-test hooks of the sampler's random-number machinery (host code, no GPU needed):
 * Philox4x32-10 block for a counter/key, and the Feistel pairing permutation of {0..n-1}. */
int ttvb_selftest_philox(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
int ttvb_selftest_perm(uint32_t n, uint32_t k0, uint32_t k1, uint32_t *perm_out);

/* Self-test of the reciprocal / reciprocal-square-root SEEDS the hot loop starts from
 * (csrc/ttvb_math.cuh, "canonical arithmetic v3"): runs the table lookup that the CPU oracle and
 * the host harness use (rcp_tab[2^20], rsq_tab[2^21] high words, host pointers; see
 * oracle/mufu_sm100a.npz) on device `device` against the special-function unit for ALL 2^32 high
 * operand words. mismatches[0/1] = number of high words where the rcp / rsqrt lookup differs from
 * the hardware (NaN results compare as equal); first_bad[0..2] / [3..5] = (operand hi, hardware
 * hi, lookup hi) of one offending word each, if any. */
int ttvb_selftest_seed_scan(int device, const uint32_t *rcp_tab, const uint32_t *rsq_tab, uint64_t mismatches[2],
                            uint32_t first_bad[6]);

/* Self-checks build (nvcc -DTTVB_CHECKS; compute-sanitizer is not available on the GPU pool): the
 * kernels verify the indices of their queue / ring / result-area / observed-table / event-table
 * accesses and bound their inter-CTA waits, counting violations by kind:
 * counts[0] queue slot, [1] batch header, [2] fold, [3] observed-table index, [4] event-table index,
 * [5] hand-over wait timed out, [6] work-unit decoding. Returns TTVB_E_UNSUPPORTED (counts = -1) from
 * the product build, which compiles the checks away. tests/test_gpu_checks.py runs the cases a
 * sanitizer run would have covered against such a build. */
int ttvb_check_failures(int device, int32_t counts[8]);

#ifdef __cplusplus
}
#endif
#endif /* TTVB_H */
