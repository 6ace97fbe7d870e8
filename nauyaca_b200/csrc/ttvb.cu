// ttvb.cu -- C ABI of libttvb.so (include/ttvb.h) over the sm_100a kernels.
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/ttvb.h"
#include "ttvb_kernels.cuh"
#include "ttvb_pt.cuh"


namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(TTVB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

constexpr long long CHUNK = 1LL << 20;   // systems per launch (bounds staging memory)
#ifndef TTVB_COMPACT_MAX_N
#define TTVB_COMPACT_MAX_N 3             // planet counts that also get the compact-refinement build of the kernel
#endif

}  // namespace

struct ttvb_handle {
    int device = 0;
    int num_sms = 0;
    int blocks_per_sm = 0;
    ttvb::DevSys sys{};          // host template; per-call fields filled at launch
    int npla = 0, ndim = 0, nobs = 0;
    std::vector<double> cube_lo, cube_hi, phys_lo, phys_hi;
    int* d_obs_epoch = nullptr;
    double* d_obs_time = nullptr;
    double* d_obs_sigma = nullptr;
    unsigned char* d_queue = nullptr;
    long long queue_stride = 0;
    int max_grid = 0;
    // device staging for host-pointer calls (grown on demand)
    unsigned char* d_stage = nullptr;
    size_t stage_bytes = 0;
    int tpb = TTVB_TPB;
    int* d_sched = nullptr;      // unit counter + per-tile segment flags
    size_t sched_ints = 0;
    unsigned char* d_ctx = nullptr;   // hand-over context of time-segmented runs
    size_t ctx_bytes = 0;
    int force_nseg = 0;          // TTVB_NSEG (0 = choose per launch)
    int force_lanes = 0;         // TTVB_LANES (0 = choose per launch): systems on the first so many lanes of a warp
    int variant = 0;             // TTVB_VARIANT: 0 = choose per launch, 1 = unrolled refinement, 2 = compact refinement
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_done = nullptr;   // end of the last launch that used the handle's scratch state
    // host-pointer calls: copy-in / copy-out streams and the events of the two staging buffers
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int* h_ready = nullptr;          // pinned: row counts the copy-in stream writes to the device counter
    size_t h_ready_cap = 0;          // (one slot per upload piece of a call: a slot is read when its copy EXECUTES)
    bool no_stream = false;          // TTVB_NO_STREAM=1: upload every launch before it starts
    long long stream_retries = 0;    // streamed launches that had to be run again (uploads did not overlap)
    bool test_stall = false;         // TTVB_STREAM_TEST_STALL=1 (tests): hold the uploads back until the kernel gives up
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_krn[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool timed = false;
    bool in_capture = false;     // the call is being captured into a CUDA graph: no events, no waits on outside work
    long long launches = 0;
};

namespace {

template <int N>
int configure(ttvb_handle* h) {
    // worst-case dynamic shared memory over the three modes
    size_t smem = ttvb::smem_bytes(N, h->nobs, 7 * N);
    // The attribute belongs to the FUNCTION, not to the handle: a second handle with a smaller
    // observed table must not lower it under the first one's launches. Opt in to the device's
    // limit once; every launch then asks for what it needs.
    int optin = 0;
    CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, ttvb::ttvb_main_kernel<N, false>));
    optin -= (int)fa.sharedSizeBytes;                 // the opt-in limit covers static + dynamic shared memory
    if (smem > (size_t)optin) return fail(TTVB_E_UNSUPPORTED, "observed table too large for shared memory (%zu B needed, %d B available)", smem, optin);
    CU(cudaFuncSetAttribute(ttvb::ttvb_main_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ttvb::ttvb_main_kernel<N, false>, TTVB_TPB, smem));
    if constexpr (N <= TTVB_COMPACT_MAX_N) {          // the compact build exists for the common planet counts
        int nb2 = 0;
        CU(cudaFuncSetAttribute(ttvb::ttvb_main_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, ttvb::ttvb_main_kernel<N, true>, TTVB_TPB, smem));
        nb = std::min(nb, nb2);
    }
    if (nb < 1) return fail(TTVB_E_CUDA, "kernel does not fit on an SM (smem %zu B)", smem);
    h->blocks_per_sm = nb;
    h->tpb = TTVB_TPB;
    h->queue_stride = ttvb::QLayout<N>::BYTES;
    return TTVB_OK;
}

// Work scheduling of one launch. A launch is cut into nseg time segments when that evens out
// the last, partially filled round of resident CTAs (every system costs the same, so a launch
// of T tiles on G resident CTAs takes ceil(T/G) rounds; with nseg segments it takes
// ceil(nseg*T/G)/nseg). Sets S.nseg/seg_len/sched/ctx and returns the grid size.
int plan_launch(ttvb_handle* h, ttvb::DevSys& S, cudaStream_t st, int* grid_out) {
    // Thin launch: a batch that fits the resident grid (at most blocks_per_sm warps per scheduler) is latency-bound,
    // and a good part of that latency is the event handling of a warp (a push phase on most steps,
    // one refinement per 64 events), which grows with the number of systems the warp holds. Such a
    // batch is spread over all the schedulers, only the first `lanes` lanes of every warp holding a
    // system (a warp with few active lanes issues no faster, but it pushes and refines that many
    // fewer events; two warps on a scheduler would each run ~1.4x longer, so one CTA per SM is the
    // limit). Same bits: a system's arithmetic does not depend on its lane. TTVB_LANES forces it.
    int lanes = 32;
    if (S.ready == nullptr) {
        if (h->force_lanes > 0) lanes = std::min(h->force_lanes, 32);
        else {
            // k resident CTAs per SM are needed at 32 lanes per warp; while the batch fits the resident
            // grid (k <= blocks_per_sm) it is spread evenly over k CTAs on EVERY SM (one more CTA on
            // a few SMs only would make those, and so the launch, as slow as a full second wave)
            const long long warps = (long long)h->num_sms * (h->tpb / 32);
            const long long k = (S.B + warps * 32 - 1) / (warps * 32);
            if (k <= h->blocks_per_sm) lanes = (int)std::min<long long>(32, (S.B + warps * k - 1) / (warps * k));
        }
    }
    S.lanes = lanes;
    const long long spt = (long long)(h->tpb / 32) * lanes;
    const long long ntiles = (S.B + spt - 1) / spt;
    const long long G = h->max_grid;
    int best = 1;
    if (h->force_nseg > 0) {
        best = h->force_nseg;
    } else if (ntiles > G) {
        double best_eff = (double)ntiles / (double)G / (double)((ntiles + G - 1) / G);
        for (int s = 2; s <= 8; s++) {
            if ((S.nsteps + 1) / s < 1000) break;             // keep segments long
            const long long units = ntiles * s;
            const double eff = (double)units / (double)G / (double)((units + G - 1) / G);
            if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
        }
    }
    if ((S.nsteps + 1) / best < 1) best = 1;
    S.nseg = best;
    S.seg_len = (S.nsteps + 1 + best - 1) / best;
    const size_t need_ints = (size_t)ntiles + 1;
    if (need_ints > h->sched_ints) {
        if (h->d_sched) CU(cudaFree(h->d_sched));
        h->d_sched = nullptr; h->sched_ints = 0;
        CU(cudaMalloc(&h->d_sched, need_ints * 4));
        h->sched_ints = need_ints;
    }
    CU(cudaMemsetAsync(h->d_sched, 0, need_ints * 4, st));
    S.sched = h->d_sched;
    S.ctx_d = nullptr; S.ctx_i = nullptr; S.ctx_stride = S.B;
    if (best > 1) {
        const size_t nd = 25 * (size_t)h->npla + 3, ni = 2 * (size_t)h->npla + 4;
        const size_t need = (size_t)S.B * (nd * 8 + ni * 4);
        if (need > h->ctx_bytes) {
            if (h->d_ctx) CU(cudaFree(h->d_ctx));
            h->d_ctx = nullptr; h->ctx_bytes = 0;
            CU(cudaMalloc(&h->d_ctx, need));
            h->ctx_bytes = need;
        }
        S.ctx_d = (double*)h->d_ctx;
        S.ctx_i = (int*)(h->d_ctx + (size_t)S.B * nd * 8);
    }
    const long long units = ntiles * best;
    *grid_out = (int)std::min<long long>(units, G);
    return TTVB_OK;
}

template <int N>
int launch_main(ttvb_handle* h, const ttvb::DevSys& S, int grid, cudaStream_t st) {
    size_t smem = ttvb::smem_bytes(N, S.nobs, S.rowlen);
    // Two builds of the kernel that differ only in how compact the refinement code is (process_batch):
    // a launch that fills the GPU with warps that refine often (more than ~1.5 events per warp-step
    // over two waves or more: the 10^7 sweep) is instruction-cache bound and takes the compact one.
    const bool dense = h->variant == 2 ||
        (h->variant == 0 && S.B >= 2LL * h->max_grid * h->tpb && 32.0 * S.nobs >= 1.5 * (double)(S.nsteps + 1));
    if constexpr (N <= TTVB_COMPACT_MAX_N) {
        if (dense) ttvb::ttvb_main_kernel<N, true><<<grid, TTVB_TPB, smem, st>>>(S);
        else ttvb::ttvb_main_kernel<N, false><<<grid, TTVB_TPB, smem, st>>>(S);
    } else {
        (void)dense;
        ttvb::ttvb_main_kernel<N, false><<<grid, TTVB_TPB, smem, st>>>(S);
    }
    CU(cudaGetLastError());
    h->launches++;
    return TTVB_OK;
}

int dispatch_configure(ttvb_handle* h) {
    switch (h->npla) {
#ifndef TTVB_DEV_ONLY_23
        case 1: return configure<1>(h);
#endif
        case 2: return configure<2>(h);
        case 3: return configure<3>(h);
#ifndef TTVB_DEV_ONLY_23
        case 4: return configure<4>(h);
        case 5: return configure<5>(h);
        case 6: return configure<6>(h);
        case 7: return configure<7>(h);
        case 8: return configure<8>(h);
        case 9: return configure<9>(h);
#endif
    }
    return fail(TTVB_E_UNSUPPORTED, "npla=%d outside 1..%d", h->npla, TTVB_MAX_PLANETS);
}

int dispatch_launch(ttvb_handle* h, const ttvb::DevSys& S, int grid, cudaStream_t st) {
    switch (h->npla) {
#ifndef TTVB_DEV_ONLY_23
        case 1: return launch_main<1>(h, S, grid, st);
#endif
        case 2: return launch_main<2>(h, S, grid, st);
        case 3: return launch_main<3>(h, S, grid, st);
#ifndef TTVB_DEV_ONLY_23
        case 4: return launch_main<4>(h, S, grid, st);
        case 5: return launch_main<5>(h, S, grid, st);
        case 6: return launch_main<6>(h, S, grid, st);
        case 7: return launch_main<7>(h, S, grid, st);
        case 8: return launch_main<8>(h, S, grid, st);
        case 9: return launch_main<9>(h, S, grid, st);
#endif
    }
    return fail(TTVB_E_UNSUPPORTED, "npla=%d outside 1..%d", h->npla, TTVB_MAX_PLANETS);
}

bool is_device_ptr(const void* p) {
    if (p == nullptr) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure_stage(ttvb_handle* h, size_t bytes) {
    if (bytes <= h->stage_bytes) return TTVB_OK;
    if (h->d_stage) { CU(cudaFree(h->d_stage)); h->d_stage = nullptr; h->stage_bytes = 0; }
    CU(cudaMalloc(&h->d_stage, bytes));
    h->stage_bytes = bytes;
    return TTVB_OK;
}

int ensure_copy_streams(ttvb_handle* h) {
    if (h->s_in) return TTVB_OK;
    CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_krn[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
    return TTVB_OK;
}

void set_mode(ttvb_handle* h, ttvb::DevSys& S, int mode) {
    S.mode = mode;
    S.rowlen = (mode == TTVB_MODE_PHYSICAL_FULL) ? 7 * h->npla : h->ndim;
    const std::vector<double>& lo = (mode == TTVB_MODE_CUBE) ? h->cube_lo : h->phys_lo;
    const std::vector<double>& hi = (mode == TTVB_MODE_CUBE) ? h->cube_hi : h->phys_hi;
    for (int i = 0; i < h->ndim; i++) { S.lo[i] = lo[i]; S.hi[i] = hi[i]; }
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

const char* ttvb_last_error(void) { return g_err; }

int ttvb_version(void) { return 1000; }

int ttvb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ttvb_create(const ttvb_system* sys, int device, ttvb_handle** out) {
    if (!sys || !out) return fail(TTVB_E_INVALID, "null argument");
    *out = nullptr;
    if (sys->npla < 1 || sys->npla > TTVB_MAX_PLANETS)
        return fail(TTVB_E_UNSUPPORTED, "npla=%d outside 1..%d", sys->npla, TTVB_MAX_PLANETS);
    if (sys->nconst < 0 || sys->ndim < 0 || sys->ndim + sys->nconst != 7 * sys->npla)
        return fail(TTVB_E_INVALID, "ndim (%d) + nconst (%d) != 7*npla (%d)", sys->ndim, sys->nconst, 7 * sys->npla);
    if (sys->nobs < 0) return fail(TTVB_E_INVALID, "nobs < 0");
    if (!(sys->dt > 0.0) || !std::isfinite(sys->t0) || !std::isfinite(sys->ftime))
        return fail(TTVB_E_INVALID, "need dt > 0 and finite t0, ftime");
    if (!(sys->mstar > 0.0) || !std::isfinite(sys->mstar)) return fail(TTVB_E_INVALID, "need a finite mstar > 0");
    int ndev = ttvb_device_count();
    if (ndev <= 0) return fail(TTVB_E_NO_DEVICE, "no CUDA device available (libttvb has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TTVB_E_NO_DEVICE, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));

    ttvb_handle* h = new (std::nothrow) ttvb_handle();
    if (!h) return fail(TTVB_E_INVALID, "out of host memory");
    h->device = device;
    if (const char* e = std::getenv("TTVB_NSEG")) h->force_nseg = std::atoi(e);   // experiments / tests
    if (const char* e = std::getenv("TTVB_NO_STREAM")) h->no_stream = std::atoi(e) != 0;
    if (const char* e = std::getenv("TTVB_LANES")) h->force_lanes = std::atoi(e);
    if (const char* e = std::getenv("TTVB_VARIANT")) h->variant = std::atoi(e);
    if (const char* e = std::getenv("TTVB_STREAM_TEST_STALL")) h->test_stall = std::atoi(e) != 0;
    h->npla = sys->npla; h->ndim = sys->ndim; h->nobs = sys->nobs;
    ttvb::DevSys& S = h->sys;
    S.npla = sys->npla; S.ndim = sys->ndim; S.nconst = sys->nconst; S.nobs = sys->nobs;
    S.gm0 = TTVB_G * sys->mstar;
    S.t0 = sys->t0; S.dt = sys->dt; S.rstar_au = sys->rstar_au; S.second_term = sys->second_term;
    // number of steps: same floating-point accumulation as the engine loop
    {
        long long n = 0;
        double T = sys->t0;
        while (T < sys->ftime) {
            T += sys->dt; n++;
            if (n > 2000000000LL) { delete h; return fail(TTVB_E_INVALID, "more than 2e9 steps"); }
        }
        S.nsteps = n;
    }
    // flat-slot source map (utils.py:440-447: constants inserted at ascending flat indices)
    {
        int ic = 0, ix = 0;
        for (int k = 0; k < 7 * sys->npla; k++) {
            if (ic < sys->nconst && sys->const_idx[ic] == k) {
                S.src[k] = -1 - ic; S.cval[ic] = sys->const_val[ic]; ic++;
            } else {
                S.src[k] = ix++;
            }
        }
        if (ic != sys->nconst || ix != sys->ndim) {
            delete h;
            return fail(TTVB_E_INVALID, "const_idx must be ascending, unique and inside 0..7*npla-1");
        }
    }
    h->cube_lo.resize(sys->ndim); h->cube_hi.resize(sys->ndim);
    h->phys_lo.resize(sys->ndim); h->phys_hi.resize(sys->ndim);
    for (int i = 0; i < sys->ndim; i++) {
        S.bi[i] = sys->bi[i];
        S.dbf[i] = sys->bf[i] - sys->bi[i];     // utils.py:426 (bf - bi)
        h->cube_lo[i] = sys->cube_lo[i]; h->cube_hi[i] = sys->cube_hi[i];
        h->phys_lo[i] = sys->phys_lo[i]; h->phys_hi[i] = sys->phys_hi[i];
    }
    // observed table: grouped by planet (ascending), epochs ascending within a planet
    std::vector<int> order(sys->nobs);
    for (int i = 0; i < sys->nobs; i++) {
        order[i] = i;
        if (sys->obs_planet[i] < 0 || sys->obs_planet[i] >= sys->npla || !(sys->obs_sigma[i] > 0.0)) {
            delete h;
            return fail(TTVB_E_INVALID, "observation %d: planet index out of range or sigma <= 0", i);
        }
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (sys->obs_planet[a] != sys->obs_planet[b]) return sys->obs_planet[a] < sys->obs_planet[b];
        return sys->obs_epoch[a] < sys->obs_epoch[b];
    });
    std::vector<int> ep(sys->nobs);
    std::vector<double> tm(sys->nobs), sg(sys->nobs);
    for (int i = 0; i <= sys->npla; i++) S.obs_start[i] = 0;
    for (int i = 0; i < sys->nobs; i++) {
        int o = order[i];
        ep[i] = sys->obs_epoch[o]; tm[i] = sys->obs_time[o]; sg[i] = sys->obs_sigma[o];
        S.obs_start[sys->obs_planet[o] + 1]++;
        if (i > 0 && sys->obs_planet[order[i - 1]] == sys->obs_planet[o] && ep[i - 1] == ep[i]) {
            delete h;
            return fail(TTVB_E_INVALID, "duplicate observed epoch %d for planet %d", ep[i], sys->obs_planet[o]);
        }
    }
    for (int i = 0; i < sys->npla; i++) S.obs_start[i + 1] += S.obs_start[i];

    int rc = TTVB_OK;
    auto bail = [&](int code) { ttvb_destroy(h); return code; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(TTVB_E_CUDA, "cudaGetDeviceProperties failed"));
    h->num_sms = prop.multiProcessorCount;
    size_t nb = std::max<size_t>(1, sys->nobs);
    if (cudaMalloc(&h->d_obs_epoch, nb * 4) != cudaSuccess || cudaMalloc(&h->d_obs_time, nb * 8) != cudaSuccess ||
        cudaMalloc(&h->d_obs_sigma, nb * 8) != cudaSuccess)
        return bail(fail(TTVB_E_CUDA, "cudaMalloc(observed table) failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (sys->nobs > 0) {
        if (cudaMemcpy(h->d_obs_epoch, ep.data(), nb * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->d_obs_time, tm.data(), nb * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->d_obs_sigma, sg.data(), nb * 8, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(TTVB_E_CUDA, "upload of the observed table failed"));
    }
    S.obs_epoch = h->d_obs_epoch; S.obs_time = h->d_obs_time; S.obs_sigma = h->d_obs_sigma;
#ifndef TTVB_NO_CONST_COEF
    {
        // the constant-bank literals of the hot loop (ttvb_math.cuh)
        const double kc[16] = {1.0 / 2.0, -1.0 / 24.0, 1.0 / 720.0, -1.0 / 40320.0, 1.0 / 3628800.0, -1.0 / 479001600.0,
                               1.0 / 87178291200.0, 1.0 / 6.0, -1.0 / 120.0, 1.0 / 5040.0, -1.0 / 362880.0,
                               1.0 / 39916800.0, -1.0 / 6227020800.0, 1.0 / 1307674368000.0, 1.0 / 3.0, 0.0};
        if (cudaMemcpyToSymbol(ttvb::TTVB_STUMPFF_C, kc, sizeof(kc)) != cudaSuccess)
            return bail(fail(TTVB_E_CUDA, "upload of the series coefficients failed"));
    }
#endif
    rc = dispatch_configure(h);
    if (rc != TTVB_OK) return bail(rc);
    h->max_grid = h->num_sms * h->blocks_per_sm;
    size_t qbytes = (size_t)h->max_grid * (h->tpb / 32) * (size_t)h->queue_stride;
    if (cudaMalloc(&h->d_queue, qbytes) != cudaSuccess) return bail(fail(TTVB_E_CUDA, "cudaMalloc(event queues, %zu B) failed", qbytes));
    S.queue = h->d_queue; S.queue_stride = h->queue_stride;

    if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(TTVB_E_CUDA, "cudaEventCreate failed"));
    *out = h;
    return TTVB_OK;
}

int ttvb_destroy(ttvb_handle* h) {
    if (!h) return TTVB_OK;
    cudaSetDevice(h->device);
    if (h->d_obs_epoch) cudaFree(h->d_obs_epoch);
    if (h->d_obs_time) cudaFree(h->d_obs_time);
    if (h->d_obs_sigma) cudaFree(h->d_obs_sigma);
    if (h->d_queue) cudaFree(h->d_queue);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_sched) cudaFree(h->d_sched);
    if (h->d_ctx) cudaFree(h->d_ctx);

    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    for (int i = 0; i < 2; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_krn[i]) cudaEventDestroy(h->ev_krn[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->h_ready) cudaFreeHost(h->h_ready);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
    return TTVB_OK;
}

int ttvb_set_cube_bounds(ttvb_handle* h, const double* lo, const double* hi) {
    if (!h || !lo || !hi) return fail(TTVB_E_INVALID, "null argument");
    for (int i = 0; i < h->ndim; i++) { h->cube_lo[i] = lo[i]; h->cube_hi[i] = hi[i]; }
    return TTVB_OK;
}

int ttvb_nsteps(const ttvb_handle* h, int64_t* nsteps) {
    if (!h || !nsteps) return fail(TTVB_E_INVALID, "null argument");
    *nsteps = h->sys.nsteps;
    return TTVB_OK;
}

// Shared driver of ttvb_loglike / ttvb_ephemeris.
static int run_batch(ttvb_handle* h, const double* x, int64_t B, int mode, double* chi2_out, double* logl_out,
                     double* chi2_planet_out, int32_t* status_out, int32_t cap, int32_t* n_events,
                     int32_t* ev_planet, int32_t* ev_epoch, double* ev_time, double* ev_rsky, double* ev_vsky,
                     void* stream) {
    if (!h) return fail(TTVB_E_INVALID, "null handle");
    if (B < 0) return fail(TTVB_E_INVALID, "B < 0");
    if (mode < 0 || mode > 2) return fail(TTVB_E_INVALID, "mode %d not in {0,1,2}", mode);
    if (B == 0) return TTVB_OK;
    if (!x) return fail(TTVB_E_INVALID, "x is null");
    const bool table = ev_time != nullptr;
    if (table && (!n_events || !ev_planet || !ev_epoch || !ev_rsky || !ev_vsky || cap < 1))
        return fail(TTVB_E_INVALID, "ephemeris outputs must all be non-null and cap >= 1");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool dev = is_device_ptr(x);
    const void* outs[] = {chi2_out, logl_out, chi2_planet_out, status_out, n_events, ev_planet, ev_epoch, ev_time, ev_rsky, ev_vsky};
    for (const void* p : outs)
        if (p && is_device_ptr(p) != dev)
            return fail(TTVB_E_INVALID, "x and the output arrays must be all host or all device pointers");

    ttvb::DevSys S = h->sys;
    set_mode(h, S, mode);
    const int rowlen = S.rowlen, N = h->npla;
    S.cap = cap;
    if (!h->in_capture) h->timed = false;
    // The handle's scratch state (unit counter and segment flags, event queues, hand-over
    // context, staging) is shared by all its launches: a call on another stream first waits for
    // the previous launch of this handle, so two streams can never run on it at the same time.
    if (!h->in_capture) CU(cudaStreamWaitEvent(st, h->ev_done, 0));

    if (dev) {
        S.B = B; S.x = x; S.ready = nullptr;
        S.chi2_out = chi2_out; S.logl_out = logl_out; S.chi2_planet_out = chi2_planet_out; S.status_out = status_out;
        S.n_events = n_events; S.ev_planet = ev_planet; S.ev_epoch = ev_epoch;
        S.ev_time = ev_time; S.ev_rsky = ev_rsky; S.ev_vsky = ev_vsky;
        int grid = 0;
        int rc = plan_launch(h, S, st, &grid);
        if (rc != TTVB_OK) return rc;
        if (h->in_capture) return dispatch_launch(h, S, grid, st);
        CU(cudaEventRecord(h->ev0, st));
        rc = dispatch_launch(h, S, grid, st);
        if (rc != TTVB_OK) return rc;
        CU(cudaEventRecord(h->ev1, st));
        CU(cudaEventRecord(h->ev_done, st));
        h->timed = true;
        return TTVB_OK;
    }

    // host pointers: STREAMED. The batch is cut into launches of up to ~2^20 systems that alternate
    // between two staging buffers. The kernel of a launch starts at once and its tiles wait (on the
    // device) for their rows: the copy-in stream uploads them in pieces of whole waves, 1 wave first
    // and doubling (one wave = as many systems as the resident CTAs hold), bumping a row counter
    // after each piece. So only the first small upload is exposed, there is no kernel boundary
    // inside a launch, the uploads of launch k+1 run under the kernel of launch k, and the results
    // of launch k are downloaded (copy-out stream) under launch k+1. The kernels themselves stay in
    // order on `st` (they share the handle's scheduling state). Returns after the results are on
    // the host.
    const long long wave = (long long)h->max_grid * h->tpb;
    long long cap_sys = std::max<long long>(wave, (CHUNK / wave) * wave);
    if (table) {
        // an event table needs cap*28 bytes per system: bound the staging of one launch to ~0.5 GiB
        const size_t per_sys = (size_t)cap * 32 + (size_t)rowlen * 8 + 64;
        cap_sys = std::max<long long>(1, std::min<long long>(cap_sys, (long long)(((size_t)1 << 29) / per_sys)));
    }
    const long long chunk = std::min<long long>(B, cap_sys);
    size_t off_x = 0, bytes = align256((size_t)chunk * rowlen * 8);
    size_t off_chi2 = bytes; bytes += align256((size_t)chunk * 8);
    size_t off_logl = bytes; bytes += align256((size_t)chunk * 8);
    size_t off_pl = bytes;   bytes += align256((size_t)chunk * N * 8);
    size_t off_st = bytes;   bytes += align256((size_t)chunk * 4);
    size_t off_ne = bytes, off_ep = 0, off_epo = 0, off_et = 0, off_er = 0, off_ev = 0;
    if (table) {
        bytes += align256((size_t)chunk * 4);
        off_ep = bytes;  bytes += align256((size_t)chunk * cap * 4);
        off_epo = bytes; bytes += align256((size_t)chunk * cap * 4);
        off_et = bytes;  bytes += align256((size_t)chunk * cap * 8);
        off_er = bytes;  bytes += align256((size_t)chunk * cap * 8);
        off_ev = bytes;  bytes += align256((size_t)chunk * cap * 8);
    }
    const size_t off_ready = bytes; bytes += 256;
    const long long nchunks = (B + chunk - 1) / chunk;
    const int nbuf = nchunks > 1 ? 2 : 1;
    int rc = ensure_stage(h, bytes * nbuf);
    if (rc != TTVB_OK) return rc;
    rc = ensure_copy_streams(h);
    if (rc != TTVB_OK) return rc;
    {
        // one pinned slot per upload piece of this call (at most ~log2(chunk/wave)+2 pieces per launch)
        size_t per_launch = 2;
        for (long long w = wave; w < chunk; w *= 2) per_launch++;
        const size_t need = (size_t)nchunks * (per_launch + 1);          // + the launch's time-out flag
        if (need > h->h_ready_cap) {
            if (h->h_ready) CU(cudaFreeHost(h->h_ready));
            h->h_ready = nullptr; h->h_ready_cap = 0;
            CU(cudaMallocHost(&h->h_ready, need * sizeof(int)));
            h->h_ready_cap = need;
        }
    }
    size_t slot = 0;
    CU(cudaEventRecord(h->ev0, st));
    // what is left to do for a launch once its kernel is in the stream: look at its time-out flag
    // (streamed launches), run it again if its uploads could not overlap, then download its results
    struct Pending { bool valid = false, streamed = false; int bf = 0; long long b0 = 0, nb = 0; int* hflag = nullptr; ttvb::DevSys S; };
    auto finish = [&](Pending& q) -> int {
        if (!q.valid) return TTVB_OK;
        unsigned char* d = h->d_stage + (size_t)q.bf * bytes;
        if (q.streamed) {
            CU(cudaEventSynchronize(h->ev_krn[q.bf]));
            if (*q.hflag != 0) {
                // the uploads did not run under the kernel (a tool serialising the GPU): its tiles gave up
                // after a bounded wait; the rows have arrived by now, run the launch again without waiting
                CU(cudaStreamSynchronize(h->s_in));
                q.S.ready = nullptr;
                int grid = 0;
                int rc2 = plan_launch(h, q.S, st, &grid);
                if (rc2 != TTVB_OK) return rc2;
                rc2 = dispatch_launch(h, q.S, grid, st);
                if (rc2 != TTVB_OK) return rc2;
                CU(cudaEventRecord(h->ev_krn[q.bf], st));
                h->stream_retries++;
                h->no_stream = true;
            }
        }
        cudaStream_t so = h->s_out;
        const long long b0 = q.b0, nb = q.nb;
        CU(cudaStreamWaitEvent(so, h->ev_krn[q.bf], 0));
        if (chi2_out) CU(cudaMemcpyAsync(chi2_out + b0, d + off_chi2, (size_t)nb * 8, cudaMemcpyDeviceToHost, so));
        if (logl_out) CU(cudaMemcpyAsync(logl_out + b0, d + off_logl, (size_t)nb * 8, cudaMemcpyDeviceToHost, so));
        if (chi2_planet_out) CU(cudaMemcpyAsync(chi2_planet_out + b0 * N, d + off_pl, (size_t)nb * N * 8, cudaMemcpyDeviceToHost, so));
        if (status_out) CU(cudaMemcpyAsync(status_out + b0, d + off_st, (size_t)nb * 4, cudaMemcpyDeviceToHost, so));
        if (table) {
            CU(cudaMemcpyAsync(n_events + b0, d + off_ne, (size_t)nb * 4, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ev_planet + b0 * cap, d + off_ep, (size_t)nb * cap * 4, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ev_epoch + b0 * cap, d + off_epo, (size_t)nb * cap * 4, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ev_time + b0 * cap, d + off_et, (size_t)nb * cap * 8, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ev_rsky + b0 * cap, d + off_er, (size_t)nb * cap * 8, cudaMemcpyDeviceToHost, so));
            CU(cudaMemcpyAsync(ev_vsky + b0 * cap, d + off_ev, (size_t)nb * cap * 8, cudaMemcpyDeviceToHost, so));
        }
        CU(cudaEventRecord(h->ev_out[q.bf], so));
        q.valid = false;
        return TTVB_OK;
    };
    Pending prev;
    for (long long k = 0; k < nchunks; k++) {
        const int bf = (int)(k % nbuf);
        const long long b0 = k * chunk, nb = std::min<long long>(chunk, B - b0);
        unsigned char* d = h->d_stage + (size_t)bf * bytes;
        int* d_ready = (int*)(d + off_ready);
        // Small launches (under two waves) are uploaded first and launched after: there is nothing to
        // hide. A streamed kernel needs its uploads to run CONCURRENTLY with it, which tools that
        // serialise the GPU (ncu, CUDA_LAUNCH_BLOCKING) do not allow: the first launch that finds this
        // out (finish(): time-out flag, rerun) switches streaming off for the handle.
        const bool streamed = nb >= 2 * wave && !h->no_stream;
        // ---- copy-in stream: the buffer is free once launch k-2 has been downloaded; zero the row
        //      counter (and the time-out flag next to it), then upload the rows piece by piece
        if (k >= nbuf) CU(cudaStreamWaitEvent(h->s_in, h->ev_out[bf], 0));
        CU(cudaMemsetAsync(d_ready, 0, 8, h->s_in));
        if (!streamed)
            CU(cudaMemcpyAsync(d + off_x, x + b0 * rowlen, (size_t)nb * rowlen * 8, cudaMemcpyHostToDevice, h->s_in));
        CU(cudaEventRecord(h->ev_in[bf], h->s_in));
        // ---- caller's stream: the kernel (the tiles of a streamed launch wait for their rows)
        S.B = nb; S.x = (const double*)(d + off_x);
        S.ready = streamed ? d_ready : nullptr;
        S.chi2_out = chi2_out ? (double*)(d + off_chi2) : nullptr;
        S.logl_out = logl_out ? (double*)(d + off_logl) : nullptr;
        S.chi2_planet_out = chi2_planet_out ? (double*)(d + off_pl) : nullptr;
        S.status_out = status_out ? (int*)(d + off_st) : nullptr;
        if (table) {
            S.n_events = (int*)(d + off_ne); S.ev_planet = (int*)(d + off_ep); S.ev_epoch = (int*)(d + off_epo);
            S.ev_time = (double*)(d + off_et); S.ev_rsky = (double*)(d + off_er); S.ev_vsky = (double*)(d + off_ev);
        } else {
            S.n_events = nullptr; S.ev_planet = nullptr; S.ev_epoch = nullptr;
            S.ev_time = nullptr; S.ev_rsky = nullptr; S.ev_vsky = nullptr;
        }
        CU(cudaStreamWaitEvent(st, h->ev_in[bf], 0));          // counter zeroed / rows uploaded (and buffer free)
        int grid = 0;
        rc = plan_launch(h, S, st, &grid);
        if (rc != TTVB_OK) return rc;
        rc = dispatch_launch(h, S, grid, st);
        if (rc != TTVB_OK) return rc;
        Pending cur;
        cur.valid = true; cur.streamed = streamed; cur.bf = bf; cur.b0 = b0; cur.nb = nb; cur.S = S;
        // ---- uploads of a streamed launch (after the launch call, so that a pageable source, whose
        //      copies block the host, cannot delay the launch)
        if (streamed) {
            if (h->test_stall) {                 // what a serialising profiler does: no upload while the kernel runs
                cudaStreamSynchronize(st);
            }
            long long done = 0, piece = std::min<long long>(wave, nb);
            while (done < nb) {
                long long n = std::min<long long>(piece, nb - done);
                if (nb - done - n < wave / 2) n = nb - done;               // no tiny last piece
                CU(cudaMemcpyAsync(d + off_x + (size_t)done * rowlen * 8, x + (b0 + done) * rowlen, (size_t)n * rowlen * 8,
                                   cudaMemcpyHostToDevice, h->s_in));
                done += n;
                if (slot >= h->h_ready_cap) return fail(TTVB_E_INVALID, "internal: upload piece count");
                int* hv = h->h_ready + slot++;
                *hv = (int)done;
                CU(cudaMemcpyAsync(d_ready, hv, 4, cudaMemcpyHostToDevice, h->s_in));
                piece *= 2;
            }
            if (slot >= h->h_ready_cap) return fail(TTVB_E_INVALID, "internal: upload piece count");
            cur.hflag = h->h_ready + slot++;
            *cur.hflag = 0;
            CU(cudaMemcpyAsync(cur.hflag, d_ready + 1, 4, cudaMemcpyDeviceToHost, st));    // after the kernel
        }
        CU(cudaEventRecord(h->ev_krn[bf], st));
        // the previous launch is finished off only now, with this one already in the stream
        rc = finish(prev);
        if (rc != TTVB_OK) return rc;
        prev = cur;
    }
    rc = finish(prev);
    if (rc != TTVB_OK) return rc;
    // the caller's stream ends after the last download; ev1 closes the timed region there
    for (int bf = 0; bf < nbuf; bf++) CU(cudaStreamWaitEvent(st, h->ev_out[bf], 0));
    CU(cudaEventRecord(h->ev1, st));
    CU(cudaEventRecord(h->ev_done, st));
    h->timed = true;
    CU(cudaStreamSynchronize(st));
    return TTVB_OK;
}

int ttvb_loglike(ttvb_handle* h, const double* x, int64_t B, int mode, double* chi2_out, double* logl_out,
                 double* chi2_planet_out, int32_t* status_out, void* stream) {
    return run_batch(h, x, B, mode, chi2_out, logl_out, chi2_planet_out, status_out, 0, nullptr, nullptr, nullptr,
                     nullptr, nullptr, nullptr, stream);
}

int ttvb_ephemeris(ttvb_handle* h, const double* x, int64_t B, int mode, int32_t cap, int32_t* n_events,
                   int32_t* planet, int32_t* epoch, double* time, double* rsky, double* vsky, int32_t* status_out,
                   void* stream) {
    if (!time) return fail(TTVB_E_INVALID, "time output is null");
    return run_batch(h, x, B, mode, nullptr, nullptr, nullptr, status_out, cap, n_events, planet, epoch, time, rsky,
                     vsky, stream);
}

int ttvb_cube_to_physical(ttvb_handle* h, const double* x, int64_t B, double* flat_out, int32_t* inside_out,
                          void* stream) {
    if (!h || !x || !flat_out) return fail(TTVB_E_INVALID, "null argument");
    if (B <= 0) return B == 0 ? TTVB_OK : fail(TTVB_E_INVALID, "B < 0");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool dev = is_device_ptr(x);
    if (is_device_ptr(flat_out) != dev || (inside_out && is_device_ptr(inside_out) != dev))
        return fail(TTVB_E_INVALID, "x and the output arrays must be all host or all device pointers");
    ttvb::DevSys S = h->sys;
    set_mode(h, S, TTVB_MODE_CUBE);
    const int nf = 7 * h->npla;
    if (dev) {
        S.B = B; S.x = x; S.ready = nullptr;
        ttvb::ttvb_cube_to_physical_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(S, flat_out, inside_out);
        CU(cudaGetLastError());
        h->launches++;
        return TTVB_OK;
    }
    const long long chunk = std::min<long long>(B, CHUNK);
    size_t bytes = align256((size_t)chunk * h->ndim * 8);
    size_t off_f = bytes; bytes += align256((size_t)chunk * nf * 8);
    size_t off_i = bytes; bytes += align256((size_t)chunk * 4);
    int rc = ensure_stage(h, bytes);
    if (rc != TTVB_OK) return rc;
    unsigned char* d = h->d_stage;
    for (long long b0 = 0; b0 < B; b0 += chunk) {
        const long long nb = std::min<long long>(chunk, B - b0);
        CU(cudaMemcpyAsync(d, x + b0 * h->ndim, (size_t)nb * h->ndim * 8, cudaMemcpyHostToDevice, st));
        S.B = nb; S.x = (const double*)d;
        ttvb::ttvb_cube_to_physical_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(S, (double*)(d + off_f),
                                                                                       inside_out ? (int*)(d + off_i) : nullptr);
        CU(cudaGetLastError());
        h->launches++;
        CU(cudaMemcpyAsync(flat_out + b0 * nf, d + off_f, (size_t)nb * nf * 8, cudaMemcpyDeviceToHost, st));
        if (inside_out) CU(cudaMemcpyAsync(inside_out + b0, d + off_i, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return TTVB_OK;
}

int ttvb_last_kernel_ms(ttvb_handle* h, float* ms) {
    if (!h || !ms) return fail(TTVB_E_INVALID, "null argument");
    if (!h->timed) return fail(TTVB_E_INVALID, "no timed call on this handle yet");
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return TTVB_OK;
}

int ttvb_stream_retries(const ttvb_handle* h, int64_t* count) {
    if (!h || !count) return fail(TTVB_E_INVALID, "null argument");
    *count = h->stream_retries;
    return TTVB_OK;
}

int ttvb_launch_count(const ttvb_handle* h, int64_t* count) {
    if (!h || !count) return fail(TTVB_E_INVALID, "null argument");
    *count = h->launches;
    return TTVB_OK;
}

// ---------------------------------------------------------------- device-resident PT sampler
}  // extern "C"

struct ttvb_pt {
    ttvb_handle* h = nullptr;
    ttvb::PTState S{};
    int mode = 0;
    double lag = 1000.0, adapt_time = 100.0;
    long long time = 0;
    int hb = 1;
    double* d_block = nullptr;     // one allocation carved into the arrays of S
    // one PT iteration captured as a CUDA graph (per value of `adapt`), replayed by ttvb_pt_run
    cudaStream_t gs = nullptr;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    long long launches_per_iter = 0;
    bool use_graph = true;
};

extern "C" {

int ttvb_pt_create(ttvb_handle* h, int32_t ntemps, int32_t nwalkers, const double* betas, double a, double lag,
                   double adapt_time, uint64_t seed, int mode, ttvb_pt** out) {
    if (!h || !betas || !out) return fail(TTVB_E_INVALID, "null argument");
    *out = nullptr;
    if (ntemps < 1 || nwalkers < 2 || (nwalkers & 1)) return fail(TTVB_E_INVALID, "need ntemps >= 1 and an even number of walkers");
    if (nwalkers < 2 * h->ndim) return fail(TTVB_E_INVALID, "The number of walkers must be greater than 2*dimension");
    if (mode != TTVB_MODE_CUBE && mode != TTVB_MODE_PHYSICAL) return fail(TTVB_E_INVALID, "mode must be cube or physical");
    if (!(a > 1.0)) return fail(TTVB_E_INVALID, "stretch scale a must be > 1");
    CU(cudaSetDevice(h->device));
    ttvb_pt* pt = new (std::nothrow) ttvb_pt();
    if (!pt) return fail(TTVB_E_INVALID, "out of host memory");
    pt->h = h; pt->mode = mode; pt->lag = lag; pt->adapt_time = adapt_time;
    const size_t nt = ntemps, nw = nwalkers, dim = h->ndim, half = nw / 2;
    const size_t n_p = nt * nw * dim, n_l = nt * nw, n_q = nt * half * dim, n_h = nt * half;
    const size_t pad = 64;         // room behind qlogl for the padded all-gather of sharded runs
    const size_t total = n_p + n_l + nt + n_q + n_h + pad + n_h + 2 * n_l + 2 * nt + nt + 1;
    if (cudaMalloc(&pt->d_block, total * 8) != cudaSuccess) { delete pt; return fail(TTVB_E_CUDA, "cudaMalloc(PT state) failed"); }
    cudaMemset(pt->d_block, 0, total * 8);
    double* c = pt->d_block;
    ttvb::PTState& S = pt->S;
    S.ntemps = ntemps; S.nwalkers = nwalkers; S.dim = (int)dim; S.a = a;
    S.p = c; c += n_p; S.logl = c; c += n_l; S.betas = c; c += nt; S.q = c; c += n_q; S.qlogl = c; c += n_h + pad;
    S.lnz = c; c += n_h; S.nprop = c; c += n_l; S.nprop_acc = c; c += n_l; S.nswap = c; c += nt; S.nswap_acc = c; c += nt;
    S.ratios = c; c += nt;
    S.time_dev = reinterpret_cast<long long*>(c); c += 1;          // zeroed with the block
    S.seed0 = (uint32_t)seed; S.seed1 = (uint32_t)(seed >> 32);
    if (const char* e = std::getenv("TTVB_PT_GRAPH")) pt->use_graph = std::atoi(e) != 0;
    int bits = 1;
    while ((1u << bits) < (unsigned)nwalkers) bits++;
    pt->hb = (bits + 1) / 2;
    if (cudaMemcpy(S.betas, betas, nt * 8, cudaMemcpyHostToDevice) != cudaSuccess) { ttvb_pt_destroy(pt); return fail(TTVB_E_CUDA, "upload of the ladder failed"); }
    *out = pt;
    return TTVB_OK;
}

int ttvb_pt_destroy(ttvb_pt* pt) {
    if (!pt) return TTVB_OK;
    for (int i = 0; i < 2; i++) if (pt->gexec[i]) cudaGraphExecDestroy(pt->gexec[i]);
    if (pt->ev_a) cudaEventDestroy(pt->ev_a);
    if (pt->ev_b) cudaEventDestroy(pt->ev_b);
    if (pt->gs) cudaStreamDestroy(pt->gs);
    if (pt->d_block) cudaFree(pt->d_block);
    delete pt;
    return TTVB_OK;
}

int ttvb_pt_set_walkers(ttvb_pt* pt, const double* p0, void* stream) {
    if (!pt || !p0) return fail(TTVB_E_INVALID, "null argument");
    ttvb_handle* h = pt->h;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const ttvb::PTState& S = pt->S;
    const int64_t B = (int64_t)S.ntemps * S.nwalkers;
    CU(cudaMemcpyAsync(S.p, p0, (size_t)B * S.dim * 8, cudaMemcpyHostToDevice, st));
    return run_batch(h, S.p, B, pt->mode, nullptr, S.logl, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                     nullptr, nullptr, stream);
}

// One PT iteration on stream `st`: propose -> likelihood of the proposals -> accept, twice, then the
// fused swap / bookkeeping / ladder kernel (which also advances the device-side iteration counter).
static int pt_iteration(ttvb_pt* pt, int adapt, cudaStream_t st) {
    ttvb_handle* h = pt->h;
    const ttvb::PTState& S = pt->S;
    const int half = S.nwalkers / 2;
    const int64_t B = (int64_t)S.ntemps * half;
    const int tb = 128, gb = (int)((B + tb - 1) / tb);
    for (int j = 0; j < 2; j++) {
        ttvb::pt_propose_kernel<<<gb, tb, 0, st>>>(S, j);
        int rc = run_batch(h, S.q, B, pt->mode, nullptr, S.qlogl, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                           nullptr, nullptr, nullptr, (void*)st);
        if (rc != TTVB_OK) return rc;
        ttvb::pt_accept_kernel<<<gb, tb, 0, st>>>(S, j);
        h->launches += 2;
    }
    ttvb::pt_swap_ladder_kernel<<<1, 1024, 0, st>>>(S, pt->hb, adapt, pt->lag, pt->adapt_time);
    h->launches++;
    CU(cudaGetLastError());
    return TTVB_OK;
}

int ttvb_pt_run(ttvb_pt* pt, int64_t iterations, int adapt, void* stream) {
    if (!pt) return fail(TTVB_E_INVALID, "null argument");
    ttvb_handle* h = pt->h;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int ai = adapt ? 1 : 0;
    int64_t n = 0;
    if (pt->use_graph && iterations >= 3) {
        // The first iteration runs plainly (it sizes the handle's scratch allocations); the second
        // is CAPTURED on an internal stream (the caller's may be the legacy default stream, which
        // cannot be captured) and the graph replayed for all the others. Nothing in an iteration
        // depends on its number except through the device-side counter, so one graph serves them all.
        int rc = pt_iteration(pt, adapt, st);
        if (rc != TTVB_OK) return rc;
        pt->time++; n++;
        if (!pt->gs) {
            CU(cudaStreamCreateWithFlags(&pt->gs, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&pt->ev_a, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&pt->ev_b, cudaEventDisableTiming));
        }
        CU(cudaEventRecord(pt->ev_a, st));
        CU(cudaStreamWaitEvent(pt->gs, pt->ev_a, 0));
        if (!pt->gexec[ai]) {
            const long long l0 = h->launches;
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(pt->gs, cudaStreamCaptureModeRelaxed));
            h->in_capture = true;
            rc = pt_iteration(pt, adapt, pt->gs);
            h->in_capture = false;
            cudaError_t ce = cudaStreamEndCapture(pt->gs, &graph);
            if (rc != TTVB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(TTVB_E_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&pt->gexec[ai], graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(TTVB_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
            pt->launches_per_iter = h->launches - l0;
            h->launches = l0;                        // capturing launched nothing
        }
        for (; n < iterations; n++) {
            CU(cudaGraphLaunch(pt->gexec[ai], pt->gs));
            h->launches += pt->launches_per_iter;
            pt->time++;
        }
        CU(cudaEventRecord(pt->ev_b, pt->gs));
        CU(cudaStreamWaitEvent(st, pt->ev_b, 0));
        CU(cudaEventRecord(h->ev_done, st));
        return TTVB_OK;
    }
    for (; n < iterations; n++) {
        int rc = pt_iteration(pt, adapt, st);
        if (rc != TTVB_OK) return rc;
        pt->time++;
    }
    return TTVB_OK;
}

// ---- one PT iteration in pieces, for runs where the proposals of a half-step are sharded
//      over several GPUs (every rank holds the whole ensemble and makes the same moves)
int ttvb_pt_half_begin(ttvb_pt* pt, int half, int64_t lo, int64_t hi, void* stream) {
    if (!pt) return fail(TTVB_E_INVALID, "null argument");
    ttvb_handle* h = pt->h;
    const ttvb::PTState& S = pt->S;
    const int64_t B = (int64_t)S.ntemps * (S.nwalkers / 2);
    if (half < 0 || half > 1 || lo < 0 || hi > B || lo > hi) return fail(TTVB_E_INVALID, "bad half or proposal range");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int tb = 128, gb = (int)((B + tb - 1) / tb);
    ttvb::pt_propose_kernel<<<gb, tb, 0, st>>>(S, half);
    h->launches++;
    CU(cudaGetLastError());
    if (hi == lo) return TTVB_OK;
    return run_batch(h, S.q + lo * S.dim, hi - lo, pt->mode, nullptr, S.qlogl + lo, nullptr, nullptr, 0, nullptr,
                     nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int ttvb_pt_half_end(ttvb_pt* pt, int half, int adapt, void* stream) {
    if (!pt) return fail(TTVB_E_INVALID, "null argument");
    if (half < 0 || half > 1) return fail(TTVB_E_INVALID, "bad half");
    ttvb_handle* h = pt->h;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const ttvb::PTState& S = pt->S;
    const int64_t B = (int64_t)S.ntemps * (S.nwalkers / 2);
    const int tb = 128, gb = (int)((B + tb - 1) / tb);
    ttvb::pt_accept_kernel<<<gb, tb, 0, st>>>(S, half);
    h->launches++;
    if (half == 1) {
        ttvb::pt_swap_ladder_kernel<<<1, 1024, 0, st>>>(S, pt->hb, adapt, pt->lag, pt->adapt_time);
        h->launches++;
        pt->time++;
    }
    CU(cudaGetLastError());
    return TTVB_OK;
}

int ttvb_pt_proposal_logl(ttvb_pt* pt, double** dev_ptr, int64_t* n, int64_t* capacity) {
    if (!pt || !dev_ptr) return fail(TTVB_E_INVALID, "null argument");
    const ttvb::PTState& S = pt->S;
    const int64_t B = (int64_t)S.ntemps * (S.nwalkers / 2);
    *dev_ptr = S.qlogl;
    if (n) *n = B;
    if (capacity) *capacity = B + 64;
    return TTVB_OK;
}

int ttvb_pt_get(ttvb_pt* pt, double* p, double* logl, double* betas, double* nprop, double* nprop_acc, double* nswap,
                double* nswap_acc, void* stream) {
    if (!pt) return fail(TTVB_E_INVALID, "null argument");
    CU(cudaSetDevice(pt->h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const ttvb::PTState& S = pt->S;
    const size_t nl = (size_t)S.ntemps * S.nwalkers;
    if (p) CU(cudaMemcpyAsync(p, S.p, nl * S.dim * 8, cudaMemcpyDeviceToHost, st));
    if (logl) CU(cudaMemcpyAsync(logl, S.logl, nl * 8, cudaMemcpyDeviceToHost, st));
    if (betas) CU(cudaMemcpyAsync(betas, S.betas, (size_t)S.ntemps * 8, cudaMemcpyDeviceToHost, st));
    if (nprop) CU(cudaMemcpyAsync(nprop, S.nprop, nl * 8, cudaMemcpyDeviceToHost, st));
    if (nprop_acc) CU(cudaMemcpyAsync(nprop_acc, S.nprop_acc, nl * 8, cudaMemcpyDeviceToHost, st));
    if (nswap) CU(cudaMemcpyAsync(nswap, S.nswap, (size_t)S.ntemps * 8, cudaMemcpyDeviceToHost, st));
    if (nswap_acc) CU(cudaMemcpyAsync(nswap_acc, S.nswap_acc, (size_t)S.ntemps * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return TTVB_OK;
}

int ttvb_selftest_philox(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    if (!counter || !key || !out) return fail(TTVB_E_INVALID, "null argument");
    uint32_t o[4];
    ttvb::Philox::gen(counter[0], counter[1], counter[2], counter[3], key[0], key[1], o);
    for (int i = 0; i < 4; i++) out[i] = o[i];
    return TTVB_OK;
}

int ttvb_selftest_perm(uint32_t n, uint32_t k0, uint32_t k1, uint32_t* perm_out) {
    if (!perm_out || n < 1) return fail(TTVB_E_INVALID, "bad argument");
    int bits = 1;
    while ((1u << bits) < n) bits++;
    const int hb = (bits + 1) / 2;
    for (uint32_t w = 0; w < n; w++) perm_out[w] = ttvb::feistel_perm(w, n, hb, k0, k1);
    return TTVB_OK;
}

int ttvb_selftest_seed_scan(int device, const uint32_t* rcp_tab, const uint32_t* rsq_tab, uint64_t mismatches[2],
                            uint32_t first_bad[6]) {
    if (!rcp_tab || !rsq_tab || !mismatches || !first_bad) return fail(TTVB_E_INVALID, "null argument");
    int ndev = ttvb_device_count();
    if (ndev <= 0) return fail(TTVB_E_NO_DEVICE, "no CUDA device available (libttvb has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(TTVB_E_NO_DEVICE, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));
    uint32_t *d_rcp = nullptr, *d_rsq = nullptr, *d_first = nullptr;
    unsigned long long* d_bad = nullptr;
    CU(cudaMalloc(&d_rcp, 4u << 20)); CU(cudaMalloc(&d_rsq, 8u << 20));
    CU(cudaMalloc(&d_bad, 16)); CU(cudaMalloc(&d_first, 24));
    CU(cudaMemcpy(d_rcp, rcp_tab, 4u << 20, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_rsq, rsq_tab, 8u << 20, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_bad, 0, 16)); CU(cudaMemset(d_first, 0, 24));
    ttvb::ttvb_seed_scan_kernel<<<148 * 16, 256>>>(d_rcp, d_rsq, d_bad, d_first);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    unsigned long long bad[2];
    CU(cudaMemcpy(bad, d_bad, 16, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(first_bad, d_first, 24, cudaMemcpyDeviceToHost));
    mismatches[0] = bad[0]; mismatches[1] = bad[1];
    cudaFree(d_rcp); cudaFree(d_rsq); cudaFree(d_bad); cudaFree(d_first);
    return TTVB_OK;
}

int ttvb_check_failures(int device, int32_t counts[8]) {
    if (!counts) return fail(TTVB_E_INVALID, "null argument");
#ifdef TTVB_CHECKS
    CU(cudaSetDevice(device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(counts, g_ttvb_check, 8 * sizeof(int)));
    return TTVB_OK;
#else
    (void)device;
    for (int i = 0; i < 8; i++) counts[i] = -1;
    return fail(TTVB_E_UNSUPPORTED, "this libttvb was built without -DTTVB_CHECKS");
#endif
}

int ttvb_fp64_peak(ttvb_handle* h, double* tflops, float* ms_out) {
    if (!h || !tflops) return fail(TTVB_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    const int threads = 256, blocks = h->num_sms * 8, iters = 4096;
    double* d = nullptr;
    CU(cudaMalloc(&d, (size_t)threads * blocks * 8));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(a, 0));
        ttvb::ttvb_fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(b, 0));
        CU(cudaEventSynchronize(b));
        h->launches++;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    CU(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    double flops = (double)threads * blocks * (double)iters * 64.0 * 2.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return TTVB_OK;
}

}  // extern "C"
