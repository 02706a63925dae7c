// gwin_b200 -- device-resident ensemble / parallel-tempering step.
//
// The reference leaves the ensemble step to emcee (gwin/sampler/emcee.py:74-76, 286-289): per
// half-step it proposes k/2 (or ntemps*k/2) points on the host, maps the likelihood over a pool and
// accepts on the host.  With the likelihood at ~1 us per walker the host side of that loop is what
// limits a multi-GPU run, so the three host stages are kernels here and the walkers never leave
// the device:
//
//   ens_propose_kernel   stretch move (emcee 2.x EnsembleSampler / PTSampler variants)
//   ens_prior_kernel     boundary conditions + log prior + waveform-parameter rows for the
//                        likelihood kernel (the distributions gwin's docs use: uniform, sin/cos
//                        angle, uniform angle (cyclic), uniform radius; optional mchirp,q -> m1,m2)
//   ens_accept_kernel    Metropolis accept with beta*logl + logp
//   ens_swap_kernel      adjacent-temperature swaps
//
// Randomness is counter based (Philox4x32-10 keyed by the seed, counter = (iteration, stage,
// index)): every rank of a multi-GPU run regenerates the same numbers, so proposals and decisions
// are identical everywhere without position traffic; only the sharded log-likelihoods are
// all-gathered (by the caller, NCCL on device buffers).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/gwin_b200.h"

#define G_TWOPI 6.283185307179586476925286766559005768
#define G_PI    3.141592653589793238462643383279502884

extern int gwin_fail(int code, const char* fmt, ...);

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t (&out)[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// two uniforms in (0, 1) with 52 random bits each
__device__ __forceinline__ void uniform2(const uint32_t (&r)[4], double& u0, double& u1) {
    const uint64_t a = ((uint64_t)r[0] << 32 | r[1]) >> 12, b = ((uint64_t)r[2] << 32 | r[3]) >> 12;
    u0 = ((double)a + 0.5) * (1.0 / 4503599627370496.0);
    u1 = ((double)b + 0.5) * (1.0 / 4503599627370496.0);
}

enum { STAGE_PROPOSE = 1, STAGE_ACCEPT = 2, STAGE_SWAP = 3 };

// index of walker j of the half being updated / of the complementary half
__device__ __forceinline__ int half_index(int j, int half, int nwalkers, int interleaved) {
    return interleaved ? 2 * j + half : half * (nwalkers / 2) + j;
}

// ---------------------------------------------------------------------------
// stretch-move proposal for one half of every temperature
//   move 0: emcee EnsembleSampler  z = ((a-1)u+1)^2/a,      q = c - z (c - s),  factor (ndim-1) ln z
//   move 1: emcee PTSampler        z = exp(U(-ln a, ln a)), q = c + z (s - c),  factor  ndim    ln z
// ---------------------------------------------------------------------------
__global__ void ens_propose_kernel(uint64_t seed, uint32_t iter, const uint32_t* __restrict__ iter_dev,
                                   int half, int ntemps, int nwalkers, int ndim,
                                   double a, int move, int interleaved,
                                   const double* __restrict__ p, double* __restrict__ q,
                                   double* __restrict__ logfac) {
    const int nh = nwalkers / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntemps * nh) return;
    if (iter_dev) iter += *iter_dev;
    const int k = i / nh, j = i % nh;
    uint32_t r[4];
    philox4x32(seed, iter, (uint32_t)(STAGE_PROPOSE * 4 + half), (uint32_t)i, 0u, r);
    double u0, u1;
    uniform2(r, u0, u1);
    const int jc = min((int)(u1 * nh), nh - 1);
    double z, lf;
    if (move == 0) { const double t = (a - 1.0) * u0 + 1.0; z = t * t / a; lf = (ndim - 1.0) * log(z); }
    else           { const double la = log(a); const double lz = la * (2.0 * u0 - 1.0); z = exp(lz); lf = ndim * lz; }
    const double* s = p + ((size_t)k * nwalkers + half_index(j, half, nwalkers, interleaved)) * ndim;
    const double* c = p + ((size_t)k * nwalkers + half_index(jc, 1 - half, nwalkers, interleaved)) * ndim;
    double* out = q + (size_t)i * ndim;
    for (int d = 0; d < ndim; ++d) out[d] = c[d] - z * (c[d] - s[d]);
    logfac[i] = lf;
}

// ---------------------------------------------------------------------------
// prior + parameter rows
// ---------------------------------------------------------------------------
struct gwin_prior_spec_dev {
    int ndim;
    int type[GWIN_ENS_MAXDIM];
    int target[GWIN_ENS_MAXDIM];
    double lo[GWIN_ENS_MAXDIM], hi[GWIN_ENS_MAXDIM];
    double mu[GWIN_ENS_MAXDIM], sigma[GWIN_ENS_MAXDIM];
    double fixed[GWIN_NPARAM];
    int derived_masses;
    int derived_type[2];
    double derived_lo[2], derived_hi[2], derived_mu[2], derived_sigma[2];
};

__device__ __forceinline__ double wrap_cyclic(double x, double lo, double hi) {
    const double w = hi - lo;
    double y = fmod(x - lo, w);
    if (y < 0.0) y += w;
    if (y >= w) y = 0.0;            // (a rounding error below lo wraps to hi itself: the domain is half-open)
    return y + lo;
}
__device__ __forceinline__ double wrap_reflect(double x, double lo, double hi) {
    const double w = hi - lo, period = 2.0 * w;
    double y = fmod(x - lo, period);
    if (y < 0.0) y += period;
    if (y > w) y = period - y;
    return y + lo;
}

// boundary condition (x is updated) and log density of one dimension
__device__ __forceinline__ double prior_term(int type, double& x, double lo, double hi, double mu, double sg) {
    switch (type) {
        case GWIN_PRIOR_UNIFORM:
            return (x >= lo && x <= hi) ? -log(hi - lo) : -INFINITY;
        case GWIN_PRIOR_UNIFORM_ANGLE:          // cyclic
            x = wrap_cyclic(x, lo, hi);
            return -log(hi - lo);
        case GWIN_PRIOR_SIN_ANGLE:              // p ~ sin x on [lo, hi] within [0, pi], reflected
            x = wrap_reflect(x, 0.0, G_PI);
            return (x >= lo && x <= hi) ? log(fabs(sin(x)) / fabs(cos(lo) - cos(hi))) : -INFINITY;
        case GWIN_PRIOR_COS_ANGLE:              // p ~ cos x on [lo, hi] within [-pi/2, pi/2], reflected
            x = wrap_reflect(x, -0.5 * G_PI, 0.5 * G_PI);
            return (x >= lo && x <= hi) ? log(fabs(cos(x)) / fabs(sin(hi) - sin(lo))) : -INFINITY;
        case GWIN_PRIOR_UNIFORM_RADIUS:         // p ~ x^2
            return (x >= lo && x <= hi) ? log(3.0 * x * x / (hi * hi * hi - lo * lo * lo)) : -INFINITY;
        case GWIN_PRIOR_GAUSSIAN: {             // normal (mu, sigma) truncated to [lo, hi]
            const double z = (x - mu) / sg;
            // mass inside the bounds, from complementary error functions on the side of the mean the interval
            // lies on (as distributions.Gaussian._norm: erf differences cancel far out in a tail)
            const double a = (lo - mu) / (sg * 1.4142135623730951), b = (hi - mu) / (sg * 1.4142135623730951);
            const double nrm = a > 0.0 ? 0.5 * (erfc(a) - erfc(b)) : (b < 0.0 ? 0.5 * (erfc(-b) - erfc(-a)) : 0.5 * (erf(b) - erf(a)));
            return (x >= lo && x <= hi) ? -0.5 * z * z - log(sg * 2.5066282746310002 * nrm) : -INFINITY;
        }
        case GWIN_PRIOR_NONE:
            return 0.0;
        default:
            return -INFINITY;
    }
}

// q [W][ndim] (sampling parameters, untouched) -> logp [W], skip [W], params [13][W]
__global__ void ens_prior_kernel(const __grid_constant__ gwin_prior_spec_dev S, int W, const double* __restrict__ q,
                                 double* __restrict__ logp, uint8_t* __restrict__ skip,
                                 double* __restrict__ params, double* __restrict__ calib) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double row[GWIN_NPARAM];
#pragma unroll
    for (int i = 0; i < GWIN_NPARAM; ++i) row[i] = S.fixed[i];
    double mchirp = 0.0, qratio = 0.0;
    double lp = 0.0;
    for (int d = 0; d < S.ndim; ++d) {
        double x = q[(size_t)w * S.ndim + d];
        const double lo = S.lo[d], hi = S.hi[d];
        lp += prior_term(S.type[d], x, lo, hi, S.mu[d], S.sigma[d]);
        const int tg = S.target[d];
        if (tg >= 0 && tg < GWIN_NPARAM) row[tg] = x;
        else if (tg == GWIN_TARGET_MCHIRP) mchirp = x;
        else if (tg == GWIN_TARGET_Q) qratio = x;
        else if (tg >= GWIN_TARGET_CALIB0 && calib) calib[(size_t)(tg - GWIN_TARGET_CALIB0) * W + w] = x;
    }
    if (mchirp > 0.0 && qratio > 0.0) {             // waveform transform (mchirp, q = m1/m2) -> (m1, m2)
        const double m2 = mchirp * pow(1.0 + qratio, 0.2) / pow(qratio, 0.6);
        row[0] = qratio * m2;
        row[1] = m2;
    }
    if (S.derived_masses) {
        // sampling transform: the walkers move in (mchirp, q), the prior is on (mass1, mass2):
        // + log |d(m1, m2)/d(mchirp, q)| = log(mchirp ((1 + q) / q^3)^(2/5))
        if (mchirp > 0.0 && qratio > 0.0) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
                lp += prior_term(S.derived_type[i], row[i], S.derived_lo[i], S.derived_hi[i],
                                 S.derived_mu[i], S.derived_sigma[i]);
            lp += log(mchirp * pow((1.0 + qratio) / (qratio * qratio * qratio), 0.4));
        } else {
            lp = -INFINITY;
        }
    }
    if (isnan(lp)) lp = -INFINITY;
    logp[w] = lp;
    skip[w] = (lp == -INFINITY) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < GWIN_NPARAM; ++i) params[(size_t)i * W + w] = row[i];
}

// ---------------------------------------------------------------------------
// accept:  ln r < logfac + (beta*logl_q + logp_q) - lnprob_s
// ---------------------------------------------------------------------------
__global__ void ens_accept_kernel(uint64_t seed, uint32_t iter, const uint32_t* __restrict__ iter_dev,
                                  int half, int ntemps, int nwalkers, int ndim,
                                  int interleaved, const double* __restrict__ betas,
                                  const double* __restrict__ q, const double* __restrict__ logfac,
                                  const double* __restrict__ q_logl, const double* __restrict__ q_logp,
                                  double* __restrict__ p, double* __restrict__ logl,
                                  double* __restrict__ lnprob, int* __restrict__ naccept) {
    const int nh = nwalkers / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntemps * nh) return;
    const int k = i / nh, j = i % nh;
    if (iter_dev) iter += *iter_dev;
    uint32_t r[4];
    philox4x32(seed, iter, (uint32_t)(STAGE_ACCEPT * 4 + half), (uint32_t)i, 0u, r);
    double u0, u1;
    uniform2(r, u0, u1);
    const size_t wi = (size_t)k * nwalkers + half_index(j, half, nwalkers, interleaved);
    const double lp = q_logp[i];
    // emcee: points outside the prior are never evaluated and carry logl = 0
    const double ll = (lp == -INFINITY) ? 0.0 : q_logl[i];
    const double newprob = (lp == -INFINITY) ? -INFINITY : betas[k] * ll + lp;
    const bool accept = log(u0) < logfac[i] + newprob - lnprob[wi];
    if (accept && !isnan(newprob)) {
        for (int d = 0; d < ndim; ++d) p[wi * ndim + d] = q[(size_t)i * ndim + d];
        logl[wi] = ll;
        lnprob[wi] = newprob;
        naccept[wi] += 1;
    }
}

// ---------------------------------------------------------------------------
// temperature swap between chains i and i-1 (emcee PTSampler._temperature_swaps); iperm / i1perm
// are random permutations of the walkers supplied by the caller.
// ---------------------------------------------------------------------------
__global__ void ens_swap_kernel(uint64_t seed, uint32_t iter, const uint32_t* __restrict__ iter_dev,
                                int level, int nwalkers, int ndim,
                                const double* __restrict__ betas, const int* __restrict__ iperm,
                                const int* __restrict__ i1perm, double* __restrict__ p,
                                double* __restrict__ logl, double* __restrict__ lnprob,
                                int* __restrict__ nswap_accepted) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwalkers) return;
    if (iter_dev) iter += *iter_dev;
    uint32_t r[4];
    philox4x32(seed, iter, (uint32_t)(STAGE_SWAP * 4), (uint32_t)(level * nwalkers + j), 0u, r);
    double u0, u1;
    uniform2(r, u0, u1);
    const double dbeta = betas[level - 1] - betas[level];
    const size_t a = (size_t)level * nwalkers + iperm[j], b = (size_t)(level - 1) * nwalkers + i1perm[j];
    const double la = logl[a], lb = logl[b];
    if (dbeta * (la - lb) > log(u0)) {
        for (int d = 0; d < ndim; ++d) {
            const double t = p[a * ndim + d];
            p[a * ndim + d] = p[b * ndim + d];
            p[b * ndim + d] = t;
        }
        const double pa = lnprob[a], pb = lnprob[b];
        logl[a] = lb; logl[b] = la;
        lnprob[a] = pb - dbeta * lb;
        lnprob[b] = pa + dbeta * la;
        atomicAdd(nswap_accepted + level, 1);
        atomicAdd(nswap_accepted + level - 1, 1);
    }
}

// All levels ntemps-1 .. 1 in ONE launch: a single CTA walks the levels in emcee's order
// (hottest pair first) and synchronises between them, because level l-1 exchanges with chain
// l-1 as level l left it.  perms: [ntemps-1][2][nwalkers]; entry (ntemps-1-level) holds the two
// permutations whose elements are paired.
__global__ void __launch_bounds__(1024)
ens_swap_all_kernel(uint64_t seed, uint32_t iter, const uint32_t* __restrict__ iter_dev,
                    int ntemps, int nwalkers, int ndim,
                    const double* __restrict__ betas, const int* __restrict__ perms, int64_t perm_stride,
                    double* __restrict__ p, double* __restrict__ logl, double* __restrict__ lnprob,
                    int* __restrict__ nswap_accepted) {
    if (iter_dev) { iter += *iter_dev; perms += (size_t)(*iter_dev) * perm_stride; }
    for (int level = ntemps - 1; level >= 1; --level) {
        const int* iperm = perms + (size_t)(ntemps - 1 - level) * 2 * nwalkers;
        const int* i1perm = iperm + nwalkers;
        const double dbeta = betas[level - 1] - betas[level];
        for (int j = threadIdx.x; j < nwalkers; j += blockDim.x) {
            uint32_t r[4];
            philox4x32(seed, iter, (uint32_t)(STAGE_SWAP * 4), (uint32_t)(level * nwalkers + j), 0u, r);
            double u0, u1;
            uniform2(r, u0, u1);
            const size_t a = (size_t)level * nwalkers + iperm[j], b = (size_t)(level - 1) * nwalkers + i1perm[j];
            const double la = logl[a], lb = logl[b];
            if (dbeta * (la - lb) > log(u0)) {
                for (int d = 0; d < ndim; ++d) {
                    const double t = p[a * ndim + d];
                    p[a * ndim + d] = p[b * ndim + d];
                    p[b * ndim + d] = t;
                }
                const double pa = lnprob[a], pb = lnprob[b];
                logl[a] = lb; logl[b] = la;
                lnprob[a] = pb - dbeta * lb;
                lnprob[b] = pa + dbeta * la;
                atomicAdd(nswap_accepted + level, 1);
                atomicAdd(nswap_accepted + level - 1, 1);
            }
        }
        __syncthreads();            // level l-1 reads what level l wrote to chain l-1
    }
}

// The same, for ensembles whose log-likelihoods fit in shared memory (ntemps x nwalkers x 8 bytes) with at
// most two walkers per thread: the random numbers, their logarithms and the pairings of every level are
// prepared before the level loop (they do not depend on the chains), the log-likelihoods -- all a decision
// reads -- live in shared memory, and only accepted swaps touch global memory (positions, lnprob).  Same
// Philox counters, same decisions, same arithmetic: the chains are bit-identical to the kernel above.
#define SWAP_MAXLV 8
__global__ void __launch_bounds__(1024)
ens_swap_all_fast_kernel(uint64_t seed, uint32_t iter, const uint32_t* __restrict__ iter_dev,
                         int ntemps, int nwalkers, int ndim,
                         const double* __restrict__ betas, const int* __restrict__ perms, int64_t perm_stride,
                         double* __restrict__ logl, double* __restrict__ lnprob,
                         int* __restrict__ nswap_accepted, int* __restrict__ idx_out) {
    extern __shared__ double s_logl[];                       // [ntemps][nwalkers]
    int* s_idx = reinterpret_cast<int*>(s_logl + (size_t)ntemps * nwalkers);   // row of p that sits at (k, w) now
    __shared__ int s_acc[SWAP_MAXLV + 1];
    if (iter_dev) { iter += *iter_dev; perms += (size_t)(*iter_dev) * perm_stride; }
    for (int i = threadIdx.x; i < ntemps * nwalkers; i += blockDim.x) { s_logl[i] = logl[i]; s_idx[i] = i; }
    if (threadIdx.x <= SWAP_MAXLV) s_acc[threadIdx.x] = 0;
    const int j0 = threadIdx.x, j1 = threadIdx.x + blockDim.x;
    double lu[SWAP_MAXLV][2];
#pragma unroll
    for (int q = 0; q < SWAP_MAXLV; ++q) {
        const int level = ntemps - 1 - q;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = h ? j1 : j0;
            lu[q][h] = 0.0;
            if (level >= 1 && j < nwalkers) {
                uint32_t r[4];
                philox4x32(seed, iter, (uint32_t)(STAGE_SWAP * 4), (uint32_t)(level * nwalkers + j), 0u, r);
                double u0, u1;
                uniform2(r, u0, u1);
                lu[q][h] = log(u0);
            }
        }
    }
    // pairings of the first level; those of the next level are fetched while a level is being decided
    int na[2], nb[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = h ? j1 : j0;
        na[h] = j < nwalkers ? perms[j] : 0;
        nb[h] = j < nwalkers ? perms[nwalkers + j] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SWAP_MAXLV; ++q) {
        const int level = ntemps - 1 - q;
        if (level >= 1) {                                      // (uniform)
            const double dbeta = betas[level - 1] - betas[level];
            const int ca[2] = {na[0], na[1]}, cb[2] = {nb[0], nb[1]};
            if (level >= 2) {
                const int* nxt = perms + (size_t)(q + 1) * 2 * nwalkers;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = h ? j1 : j0;
                    na[h] = j < nwalkers ? nxt[j] : 0;
                    nb[h] = j < nwalkers ? nxt[nwalkers + j] : 0;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = h ? j1 : j0;
                if (j < nwalkers) {
                    const int a = level * nwalkers + ca[h], b = (level - 1) * nwalkers + cb[h];
                    const double la = s_logl[a], lb = s_logl[b];
                    if (dbeta * (la - lb) > lu[q][h]) {
                        // the positions are not moved here (one SM's load/store queue would be the bottleneck):
                        // the exchange is recorded and applied to all rows at once afterwards
                        { const int t = s_idx[a]; s_idx[a] = s_idx[b]; s_idx[b] = t; }
                        const double qa = lnprob[a], qb = lnprob[b];
                        s_logl[a] = lb; s_logl[b] = la;
                        lnprob[a] = qb - dbeta * lb;
                        lnprob[b] = qa + dbeta * la;
                        atomicAdd(s_acc + level, 1);
                        atomicAdd(s_acc + level - 1, 1);
                    }
                }
            }
            __syncthreads();            // level l-1 reads what level l wrote to chain l-1
        }
    }
    for (int i = threadIdx.x; i < ntemps * nwalkers; i += blockDim.x) { logl[i] = s_logl[i]; idx_out[i] = s_idx[i]; }
    if (threadIdx.x < ntemps && s_acc[threadIdx.x]) atomicAdd(nswap_accepted + threadIdx.x, s_acc[threadIdx.x]);
}
// rows of p to where the swaps sent them: tmp[r] = p[idx[r]], then p = tmp
__global__ void ens_rows_gather_kernel(int n, int ndim, const int* __restrict__ idx, const double* __restrict__ p,
                                       double* __restrict__ tmp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * ndim) tmp[i] = p[(size_t)idx[i / ndim] * ndim + i % ndim];
}
__global__ void ens_rows_copy_kernel(int n, const double* __restrict__ tmp, double* __restrict__ p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = tmp[i];
}

// History row `it (+ *it_dev)` of the chain [W][niter][ndim], lnprob [W][niter], logl [W][niter]
__global__ void ens_record_kernel(int W, int ndim, int niter, uint32_t it, const uint32_t* __restrict__ it_dev,
                                  const double* __restrict__ p, const double* __restrict__ lnprob,
                                  const double* __restrict__ logl, double* __restrict__ chain,
                                  double* __restrict__ lnp_h, double* __restrict__ lnl_h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * ndim) return;
    if (it_dev) it += *it_dev;
    if (it >= (uint32_t)niter) return;
    const int w = i / ndim, d = i % ndim;
    chain[((size_t)w * niter + it) * ndim + d] = p[i];
    if (d == 0) {
        lnp_h[(size_t)w * niter + it] = lnprob[w];
        lnl_h[(size_t)w * niter + it] = logl[w];
    }
}
__global__ void ens_tick_kernel(uint32_t* counter) { *counter += 1u; }

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
#define CKE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return gwin_fail(GWIN_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)

static int check_dims(int ntemps, int nwalkers, int ndim) {
    if (ntemps < 1 || nwalkers < 2 || (nwalkers & 1) || ndim < 1 || ndim > GWIN_ENS_MAXDIM)
        return gwin_fail(GWIN_EINVAL, "need ntemps >= 1, even nwalkers >= 2, 1 <= ndim <= %d", GWIN_ENS_MAXDIM);
    return GWIN_OK;
}

extern "C" int gwin_ens_propose(uint64_t seed, uint32_t iter, const uint32_t* iter_dev,
                                int half, int ntemps, int nwalkers, int ndim,
                                double a, int move, int interleaved,
                                const double* p, double* q, double* logfac, void* stream) {
    int rc = check_dims(ntemps, nwalkers, ndim);
    if (rc) return rc;
    if (!p || !q || !logfac) return gwin_fail(GWIN_EINVAL, "NULL buffer");
    const int n = ntemps * (nwalkers / 2);
    ens_propose_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, iter, iter_dev, half, ntemps, nwalkers, ndim,
                                                                           a, move, interleaved, p, q, logfac);
    CKE(cudaGetLastError());
    return GWIN_OK;
}

extern "C" int gwin_ens_prior(const gwin_prior_spec* spec, int64_t W, const double* q,
                              double* logp, uint8_t* skip, double* params, void* stream) {
    return gwin_ens_prior_calib(spec, W, q, logp, skip, params, nullptr, 0, stream);
}

extern "C" int gwin_ens_prior_calib(const gwin_prior_spec* spec, int64_t W, const double* q,
                                    double* logp, uint8_t* skip, double* params,
                                    double* calib, int n_calib, void* stream) {
    if (!spec || !q || !logp || !skip || !params) return gwin_fail(GWIN_EINVAL, "NULL argument");
    if (spec->ndim < 1 || spec->ndim > GWIN_ENS_MAXDIM) return gwin_fail(GWIN_EINVAL, "bad ndim");
    if (W <= 0) return GWIN_OK;
    if (n_calib < 0 || (n_calib > 0 && !calib)) return gwin_fail(GWIN_EINVAL, "bad calibration buffer");
    gwin_prior_spec_dev S;
    S.ndim = spec->ndim;
    for (int d = 0; d < GWIN_ENS_MAXDIM; ++d) {
        S.type[d] = spec->type[d]; S.target[d] = spec->target[d];
        S.lo[d] = spec->lo[d]; S.hi[d] = spec->hi[d];
        S.mu[d] = spec->mu[d]; S.sigma[d] = spec->sigma[d];
        if (d < spec->ndim && spec->target[d] >= GWIN_TARGET_CALIB0 && n_calib > 0) {
            const int v = spec->target[d] - GWIN_TARGET_CALIB0;
            if (v >= n_calib) return gwin_fail(GWIN_EINVAL, "dimension %d targets calibration value %d of %d", d, v, n_calib);
        }
        if (d < spec->ndim && spec->type[d] == GWIN_PRIOR_GAUSSIAN && !(spec->sigma[d] > 0.0))
            return gwin_fail(GWIN_EINVAL, "dimension %d: sigma must be positive", d);
    }
    for (int i = 0; i < GWIN_NPARAM; ++i) S.fixed[i] = spec->fixed[i];
    S.derived_masses = spec->derived_masses;
    for (int i = 0; i < 2; ++i) {
        S.derived_type[i] = spec->derived_type[i];
        S.derived_lo[i] = spec->derived_lo[i]; S.derived_hi[i] = spec->derived_hi[i];
        S.derived_mu[i] = spec->derived_mu[i]; S.derived_sigma[i] = spec->derived_sigma[i];
    }
    ens_prior_kernel<<<(unsigned)((W + 127) / 128), 128, 0, (cudaStream_t)stream>>>(S, (int)W, q, logp, skip, params,
                                                                                    n_calib > 0 ? calib : nullptr);
    CKE(cudaGetLastError());
    return GWIN_OK;
}

extern "C" int gwin_ens_accept(uint64_t seed, uint32_t iter, const uint32_t* iter_dev,
                               int half, int ntemps, int nwalkers, int ndim,
                               int interleaved, const double* betas, const double* q, const double* logfac,
                               const double* q_logl, const double* q_logp,
                               double* p, double* logl, double* lnprob, int* naccept, void* stream) {
    int rc = check_dims(ntemps, nwalkers, ndim);
    if (rc) return rc;
    const int n = ntemps * (nwalkers / 2);
    ens_accept_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, iter, iter_dev, half, ntemps, nwalkers, ndim,
                                                                          interleaved, betas, q, logfac, q_logl, q_logp,
                                                                          p, logl, lnprob, naccept);
    CKE(cudaGetLastError());
    return GWIN_OK;
}

extern "C" int gwin_ens_swap(uint64_t seed, uint32_t iter, const uint32_t* iter_dev, int level, int nwalkers, int ndim,
                             const double* betas, const int* iperm, const int* i1perm,
                             double* p, double* logl, double* lnprob, int* nswap_accepted, void* stream) {
    if (level < 1 || nwalkers < 1 || ndim < 1) return gwin_fail(GWIN_EINVAL, "bad swap arguments");
    ens_swap_kernel<<<(nwalkers + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, iter, iter_dev, level, nwalkers, ndim, betas,
                                                                              iperm, i1perm, p, logl, lnprob,
                                                                              nswap_accepted);
    CKE(cudaGetLastError());
    return GWIN_OK;
}

extern "C" int gwin_ens_swap_all(uint64_t seed, uint32_t iter, const uint32_t* iter_dev,
                                 int ntemps, int nwalkers, int ndim,
                                 const double* betas, const int* perms, int64_t perm_stride,
                                 double* p, double* logl, double* lnprob, int* nswap_accepted, void* stream) {
    if (ntemps < 2) return GWIN_OK;
    if (nwalkers < 1 || ndim < 1 || !betas || !perms || !p || !logl || !lnprob || !nswap_accepted)
        return gwin_fail(GWIN_EINVAL, "bad swap arguments");
    const int threads = nwalkers >= 1024 ? 1024 : ((nwalkers + 31) / 32) * 32;
    const size_t smem = (sizeof(double) + sizeof(int)) * (size_t)ntemps * nwalkers;
    static const bool fast_on = !(getenv("GWIN_SWAP_FAST") && atoi(getenv("GWIN_SWAP_FAST")) == 0);
    if (fast_on && ntemps - 1 <= SWAP_MAXLV && nwalkers <= 2 * threads && smem <= 220 * 1024) {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        static bool attr_set_dev[64] = {};
        static double* tmp_dev[64] = {};
        static int* idx_dev[64] = {};
        static size_t cap_dev[64] = {};
        if (!attr_set_dev[dev]) {
            cudaFuncSetAttribute(ens_swap_all_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
            attr_set_dev[dev] = true;
        }
        const size_t n = (size_t)ntemps * nwalkers;
        if (n * ndim > cap_dev[dev]) {
            // (scratch for the rows on their way: grown outside any stream capture -- a sampler's first
            //  iteration always runs uncaptured)
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing((cudaStream_t)stream, &cap);
            if (cap != cudaStreamCaptureStatusNone) return gwin_fail(GWIN_EINVAL, "swap scratch has to grow: not inside a stream capture");
            CKE(cudaDeviceSynchronize());
            cudaFree(tmp_dev[dev]); cudaFree(idx_dev[dev]);
            tmp_dev[dev] = nullptr; idx_dev[dev] = nullptr; cap_dev[dev] = 0;
            CKE(cudaMalloc(&tmp_dev[dev], sizeof(double) * n * ndim));
            CKE(cudaMalloc(&idx_dev[dev], sizeof(int) * n * ndim));
            cap_dev[dev] = n * ndim;
        }
        ens_swap_all_fast_kernel<<<1, threads, smem, (cudaStream_t)stream>>>(seed, iter, iter_dev, ntemps, nwalkers, ndim, betas, perms,
                                                                             perm_stride, logl, lnprob, nswap_accepted, idx_dev[dev]);
        const int nn = (int)(n * ndim);
        ens_rows_gather_kernel<<<(nn + 255) / 256, 256, 0, (cudaStream_t)stream>>>((int)n, ndim, idx_dev[dev], p, tmp_dev[dev]);
        ens_rows_copy_kernel<<<(nn + 255) / 256, 256, 0, (cudaStream_t)stream>>>(nn, tmp_dev[dev], p);
    } else {
        ens_swap_all_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(seed, iter, iter_dev, ntemps, nwalkers, ndim, betas, perms,
                                                                     perm_stride, p, logl, lnprob, nswap_accepted);
    }
    CKE(cudaGetLastError());
    return GWIN_OK;
}

extern "C" int gwin_ens_record(int64_t W, int ndim, int64_t niter, uint32_t it, const uint32_t* it_dev,
                               const double* p, const double* lnprob, const double* logl,
                               double* chain, double* lnp_h, double* lnl_h, void* stream) {
    if (W < 1 || ndim < 1 || niter < 1 || W * ndim > (1 << 30) || niter > (1u << 31) - 1)
        return gwin_fail(GWIN_EINVAL, "bad history dimensions");
    if (!p || !lnprob || !logl || !chain || !lnp_h || !lnl_h) return gwin_fail(GWIN_EINVAL, "NULL buffer");
    const int n = (int)(W * ndim);
    ens_record_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((int)W, ndim, (int)niter, it, it_dev,
                                                                        p, lnprob, logl, chain, lnp_h, lnl_h);
    CKE(cudaGetLastError());
    return GWIN_OK;
}

extern "C" int gwin_ens_tick(uint32_t* counter, void* stream) {
    if (!counter) return gwin_fail(GWIN_EINVAL, "counter is NULL");
    ens_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    CKE(cudaGetLastError());
    return GWIN_OK;
}
