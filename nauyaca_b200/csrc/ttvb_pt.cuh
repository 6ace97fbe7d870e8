// ttvb_pt.cuh -- parallel-tempering ensemble moves on the device.
//
// The sampler the reference drives (ptemcee 1.0.0 through mcmc.py:195-221) spends its time in
// the likelihood of ntemps*nwalkers/2 proposals per half-step; everything around it (stretch
// proposals, accept/reject, temperature swaps, ladder adaptation) is a few passes over small
// arrays. Keeping those passes on the GPU removes the host round trip of every half-step: the
// walkers, their log-likelihoods and the ladder stay resident, and one PT iteration is a
// stream of small kernels around two launches of ttvb_main_kernel.
//
// Algorithm per iteration (ptemcee's), flat prior inside the bounds
// (the reference's default MCMC.logprior == 0.0, mcmc.py:505-509; bounds are enforced by the
// likelihood returning -inf):
//   for half j in (0,1):   q = p_s[js] + z (p_u - p_s[js]),  ln z ~ U(-ln a, ln a), js ~ U{0..n/2-1}
//                          accept if ln u < dim ln z + beta (logl_q - logl)
//   for i = ntemps-1 .. 1: pair walker perm_i(w) of chain i with perm'_i(w) of chain i-1,
//                          swap if ln u < (beta_{i-1}-beta_i)(logl_i - logl_{i-1})
//   ladder: Vousden et al. (2016) adaptation with decay lag/(t+lag)/time.
// Random numbers: Philox4x32-10 (counter-based, one counter per walker and purpose, so the
// result does not depend on launch geometry). Random pairings are Feistel permutations with
// cycle walking (no sort needed, every walker computes its partner independently).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttvb {

struct Philox {
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // Philox4x32-10
    __host__ __device__ static inline void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t (&out)[4]) {
        uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
        for (int r = 0; r < 10; r++) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    // uniform in (0,1) from 53 random bits
    __host__ __device__ static inline double u01(uint32_t a, uint32_t b) {
        const uint64_t m = (((uint64_t)a << 32) | b) >> 11;
        return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
    }
};

// purposes (third counter word)
enum { PT_RNG_STRETCH = 1, PT_RNG_ACCEPT = 2, PT_RNG_SWAP = 3, PT_RNG_PERM = 4 };

// pseudo-random permutation of {0..n-1}: 4-round Feistel network on 2*hb bits, cycle walking
__host__ __device__ inline uint32_t feistel_perm(uint32_t w, uint32_t n, int hb, uint32_t k0, uint32_t k1) {
    const uint32_t mask = (1u << hb) - 1u;
    uint32_t x = w;
    do {
        uint32_t l = x >> hb, r = x & mask;
#pragma unroll
        for (int rd = 0; rd < 4; rd++) {
            uint32_t o[4];
            Philox::gen(r, (uint32_t)rd, PT_RNG_PERM, 0u, k0, k1, o);
            const uint32_t nl = r, nr = l ^ (o[0] & mask);
            l = nl; r = nr;
        }
        x = (l << hb) | r;
    } while (x >= n);
    return x;
}

struct PTState {
    int ntemps, nwalkers, dim;
    double a;                    // stretch scale
    double* p;                   // [ntemps][nwalkers][dim]
    double* logl;                // [ntemps][nwalkers]
    double* betas;               // [ntemps]
    double* q;                   // [ntemps*nwalkers/2][dim] proposals
    double* qlogl;               // [ntemps*nwalkers/2]
    double* lnz;                 // [ntemps*nwalkers/2]
    double* nprop;               // [ntemps][nwalkers]
    double* nprop_acc;
    double* nswap;               // [ntemps]
    double* nswap_acc;
    double* ratios;              // [ntemps-1] swap acceptance of the current iteration
    long long* time_dev;         // iteration counter ON THE DEVICE: every kernel reads it from here, so that
                                 // one iteration is the same work whatever its number (CUDA-graph replay)
    uint32_t seed0, seed1;
};

// stretch proposals of half `j` at iteration `it`
__global__ void pt_propose_kernel(PTState S, int j) {
    const long long it = *S.time_dev;
    const int half = S.nwalkers / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S.ntemps * half) return;
    const int t = idx / half, w = idx - t * half;
    uint32_t o[4];
    Philox::gen((uint32_t)idx, (uint32_t)(2 * it + j), PT_RNG_STRETCH, (uint32_t)((2 * it + j) >> 32), S.seed0, S.seed1, o);
    const double lna = log(S.a);
    const double lz = (2.0 * Philox::u01(o[0], o[1]) - 1.0) * lna;
    const double z = exp(lz);
    int js = (int)(Philox::u01(o[2], o[3]) * half);
    if (js >= half) js = half - 1;
    const double* pu = S.p + ((size_t)t * S.nwalkers + (2 * w + j)) * S.dim;
    const double* ps = S.p + ((size_t)t * S.nwalkers + (2 * js + ((j + 1) & 1))) * S.dim;
    double* q = S.q + (size_t)idx * S.dim;
    for (int d = 0; d < S.dim; d++) q[d] = ps[d] + z * (pu[d] - ps[d]);
    S.lnz[idx] = lz;
}

// accept / reject the proposals of half `j`
__global__ void pt_accept_kernel(PTState S, int j) {
    const long long it = *S.time_dev;
    const int half = S.nwalkers / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S.ntemps * half) return;
    const int t = idx / half, w = idx - t * half;
    const int wi = t * S.nwalkers + (2 * w + j);
    uint32_t o[4];
    Philox::gen((uint32_t)idx, (uint32_t)(2 * it + j), PT_RNG_ACCEPT, (uint32_t)((2 * it + j) >> 32), S.seed0, S.seed1, o);
    const double beta = S.betas[t];
    const double ql = S.qlogl[idx], ol = S.logl[wi];
    // tempered likelihoods; beta * (-inf) with beta == 0 is NaN -> -inf, as ptemcee does
    double qt = beta * ql, ot = beta * ol;
    if (isnan(qt)) qt = -INFINITY;
    if (isnan(ot)) ot = -INFINITY;
    const double logpaccept = S.dim * S.lnz[idx] + qt - ot;
    const bool acc = log(Philox::u01(o[0], o[1])) < logpaccept;
    if (acc) {
        const double* q = S.q + (size_t)idx * S.dim;
        double* p = S.p + (size_t)wi * S.dim;
        for (int d = 0; d < S.dim; d++) p[d] = q[d];
        S.logl[wi] = ql;
    }
    S.nprop[wi] += 1.0;
    if (acc) S.nprop_acc[wi] += 1.0;
}

// Temperature swaps of one iteration (chains i and i-1, hottest pair first), the swap bookkeeping,
// the ladder adaptation and the advance of the iteration counter: ONE block. The pairs of a level
// are disjoint (two permutations of the walkers), the levels are separated by block barriers.
__global__ void __launch_bounds__(1024) pt_swap_ladder_kernel(PTState S, int hb, int adapt, double lag, double adapt_time) {
    const long long it = *S.time_dev;
    const int nt = S.ntemps;
    __shared__ int s_acc;
    for (int i = nt - 1; i >= 1; i--) {
        if (threadIdx.x == 0) s_acc = 0;
        __syncthreads();
        const uint32_t k0 = S.seed0 ^ (uint32_t)(it * 2654435761u), k1 = S.seed1 + (uint32_t)(it >> 32);
        const double dbeta = S.betas[i - 1] - S.betas[i];
        for (int w = threadIdx.x; w < S.nwalkers; w += blockDim.x) {
            const uint32_t a = feistel_perm((uint32_t)w, (uint32_t)S.nwalkers, hb, k0, k1 ^ (uint32_t)(2 * i));
            const uint32_t b = feistel_perm((uint32_t)w, (uint32_t)S.nwalkers, hb, k0, k1 ^ (uint32_t)(2 * i + 1));
            uint32_t o[4];
            Philox::gen((uint32_t)w, (uint32_t)it, PT_RNG_SWAP, (uint32_t)i, S.seed0, S.seed1, o);
            const int ia = i * S.nwalkers + (int)a, ib = (i - 1) * S.nwalkers + (int)b;
            const double la = S.logl[ia], lb = S.logl[ib];
            const double paccept = dbeta * (la - lb);
            if (paccept > log(Philox::u01(o[0], o[1]))) {
                double* pa = S.p + (size_t)ia * S.dim;
                double* pb = S.p + (size_t)ib * S.dim;
                for (int d = 0; d < S.dim; d++) { const double tmp = pa[d]; pa[d] = pb[d]; pb[d] = tmp; }
                S.logl[ia] = lb;
                S.logl[ib] = la;
                atomicAdd(&s_acc, 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) S.ratios[i - 1] = (double)s_acc;
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    for (int i = nt - 1; i >= 1; i--) {
        const double nacc = S.ratios[i - 1];
        S.nswap[i] += S.nwalkers; S.nswap[i - 1] += S.nwalkers;
        S.nswap_acc[i] += nacc; S.nswap_acc[i - 1] += nacc;
        S.ratios[i - 1] = nacc / S.nwalkers;
    }
    if (adapt && nt > 2) {
        const double decay = lag / ((double)it + lag);
        const double kappa = decay / adapt_time;
        // deltaTs = diff(1/betas[:-1]) * exp(kappa (ratios[:-1] - ratios[1:])); betas[1:-1] = 1/(cumsum + 1/betas[0])
        double cum = 1.0 / S.betas[0];
        double prev_inv = 1.0 / S.betas[0];
        for (int k = 1; k <= nt - 2; k++) {
            const double inv_k = 1.0 / S.betas[k];
            const double dT = (inv_k - prev_inv) * exp(kappa * (S.ratios[k - 1] - S.ratios[k]));
            prev_inv = inv_k;
            cum += dT;
            S.betas[k] = 1.0 / cum;
        }
    }
    for (int i = 0; i < nt - 1; i++) S.ratios[i] = 0.0;
    *S.time_dev = it + 1;
}

}  // namespace ttvb
