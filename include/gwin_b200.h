/*
 * gwin_b200 -- C ABI of the B200-native batched log-likelihood-ratio path.
 *
 * The reference (gwastro/gwin) is pure Python and has no FFI; its "plugin API"
 * for this path is the model protocol (gwin/models/gaussian_noise.py:219-350,
 * gwin/models/marginalized_gaussian_noise.py:456-511) called through
 * CallModel (gwin/models/__init__.py:101-137) by pool.map.  This header is the
 * C boundary a maintainer would bind beneath those classes (ctypes stub in
 * INTEGRATION.md).  Plain pointers and sizes only; no torch types.
 *
 * All functions return 0 on success or a negative GWIN_E* code; the message of
 * the last failure on the calling thread is available from gwin_last_error().
 * No exceptions cross the ABI.
 */
#ifndef GWIN_B200_H
#define GWIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWIN_OK            0
#define GWIN_EINVAL       -1   /* bad argument (mirrors the reference's ValueError) */
#define GWIN_ECUDA        -2   /* CUDA runtime failure */
#define GWIN_ENOMEM       -3
#define GWIN_ENODEVICE    -4   /* no usable sm_100 device: there is NO CPU fallback */

#define GWIN_MAX_DET       4
#define GWIN_NPARAM       13   /* per-walker parameter row, see gwin_loglr_batch */

/* loglr flavour */
#define GWIN_MODE_GAUSSIAN     0  /* gaussian_noise.py:304-350:  sum_det Re<d|h> - <h|h>/2        */
#define GWIN_MODE_MARG_PHASE   1  /* marginalized_gaussian_noise.py:456-459, 461-511:
                                     log I0(|sum_det <d|h>|) - sum_det <h|h>/2                     */
#define GWIN_MODE_MARG_DIST    2  /* marginalized_gaussian_noise.py:441-448 (_margdist_loglr):
                                     logsumexp_j(mf/D_j - opt/(2 D_j^2), b = w_j),
                                     mf = |sum_det <d|h>|, opt = sum_det <h|h>                     */
#define GWIN_MODE_MARG_DIST_PHASE 3 /* :431-439 (_margdistphase_loglr): as above with
                                     log I0(mf) in place of mf (the reference divides the LOG
                                     by D_j; kept)                                                  */

/* kernel selection (diagnostic; production uses GWIN_KERNEL_AUTO) */
#define GWIN_KERNEL_AUTO       0
#define GWIN_KERNEL_DIRECT     1  /* per-bin FP64 phase + sincos                                  */
#define GWIN_KERNEL_RECURRENCE 2  /* block-anchored chirp-phasor recurrence + Horner accumulation */
#define GWIN_KERNEL_TABLE      3  /* per-plan moment tables + NUFFT interpolation wherever sub-blocks of
                                     >= 128 bins qualify, recurrence below and for the walkers' last
                                     bins (what AUTO selects on long segments; uncalibrated only) */

typedef struct gwin_plan gwin_plan;

/*
 * Plan build == GaussianNoise.__init__ (gaussian_noise.py:219-265):
 *   kmin,kmax = get_cutoff_indices(f_lower, f_upper, delta_f, (n_f-1)*2)   :247
 *   weight    = sqrt(norm/psd), norm defaults to 4*delta_f                  :251-262
 *   data[kmin:kmax] *= weight[kmin:kmax]                                    :264-265
 * and the lazily cached lognl (:275-291).  The plan keeps device replicas of
 * the (over-whitened, conjugated, amplitude-pre-scaled) data tables; the
 * caller's arrays are only read (the reference copies them, base_data.py:87).
 *
 *   device          CUDA ordinal
 *   n_det           1..GWIN_MAX_DET detectors
 *   n_f             bins per frequency series (all detectors equal, :240-242)
 *   delta_f         bin spacing [Hz]
 *   epoch_gps       data epoch == waveform-generator epoch (:235-238)
 *   f_lower         likelihood low-frequency cutoff; f_upper <= 0 means None
 *   f_lower_wf      waveform generator's f_lower (bins < ceil(f_lower_wf/df) are 0)
 *   det_response    [n_det][9] row-major detector response tensors
 *   det_location    [n_det][3] vertex locations [m], Earth-fixed frame
 *   data            [n_det][n_f][2] host, (re,im) interleaved, un-whitened d~(f)
 *   psd             [n_det][n_f] host, or NULL for unit PSD (:254-256)
 *   norm            <= 0 selects the default 4*delta_f
 */
int gwin_plan_create(gwin_plan** out, int device, int n_det, int64_t n_f,
                     double delta_f, double epoch_gps,
                     double f_lower, double f_upper, double f_lower_wf,
                     const double* det_response, const double* det_location,
                     const double* data, const double* psd, double norm);

void gwin_plan_destroy(gwin_plan* plan);

/* kmin, kmax as computed by the plan (gaussian_noise.py:247-250). */
int gwin_plan_cutoffs(const gwin_plan* plan, int64_t* kmin, int64_t* kmax);

/* lognl (gaussian_noise.py:275-291): total and per detector [n_det]. */
int gwin_plan_lognl(const gwin_plan* plan, double* lognl_total, double* lognl_det);

/* Tuning knobs: kernel = GWIN_KERNEL_*, phase_tol = max block-polynomial phase
 * error [rad] of the recurrence kernel (default 1e-10), mchirp_min [Msun] =
 * smallest chirp mass the block schedule must cover (default 0.5; the host entry points and the
 * table path derive it from the batches unless it is set here).  0 leaves a value unchanged. */
int gwin_plan_configure(gwin_plan* plan, int kernel, double phase_tol, double mchirp_min);

/*
 * Coverage of the moment tables of GWIN_KERNEL_TABLE / AUTO (no counterpart in the reference: a
 * property of this implementation).  The tables are built for a range of the
 * Newtonian chirp coefficient (i.e. of the chirp mass) and a bound on how far the other phase
 * coefficients (mass ratio, spins, tides) lie from those of an equal-mass, non-spinning chirp.
 * By default the library measures both on the first batch that uses the tables (one small kernel
 * and one synchronisation; the host entry points measure again after a batch that sent many walkers
 * down the slow path).  A caller that knows its prior declares it instead:
 * gwin_plan_cover_host / _device take W parameter points (layout as gwin_loglr_batch_strided; e.g.
 * a few thousand prior draws), widen their range by a margin, and add it to what is covered;
 * W = 0 returns to automatic coverage.  gwin_plan_cover_device synchronises `stream`.
 * Walkers outside the coverage are never wrong, only slow: see gwin_plan_slow_path.
 */
int gwin_plan_cover_host(gwin_plan* plan, int64_t W, const double* params,
                         int64_t stride_walker, int64_t stride_param);
int gwin_plan_cover_device(gwin_plan* plan, int64_t W, const double* params,
                           int64_t stride_walker, int64_t stride_param, void* stream);

/* Sub-blocks, reference chirps and bytes of the current moment tables (0 if none), and the first
 * bin they take (kmax if none). */
int gwin_plan_table_info(gwin_plan* plan, int64_t* n_subblocks, int64_t* n_references,
                         int64_t* bytes, int64_t* k_split);

/*
 * Slow path.  The recurrence schedule holds for chirp masses >= mchirp_min and the moment tables for
 * their coverage; a walker outside either (seen when its coefficients are set up, or -- for the
 * tables -- when its own truncation estimate of a sub-block is too large) is evaluated by the
 * direct kernel instead (per-bin phase + sincos: ~20x slower, same tolerance).  Returns the number of
 * such walkers since the previous call and resets it (synchronises the device).  A count that
 * keeps growing means the coverage should be re-declared.
 */
int gwin_plan_slow_path(gwin_plan* plan, int64_t* n_walkers);

/*
 * Distance-marginalisation grid == MarginalizedGaussianNoise._setup_prior
 * (marginalized_gaussian_noise.py:368-375): dist[n] = numpy.linspace(min, max, 10**4),
 * weight[n] = deltad * dist_prior[j] (the ``b`` of the reference's logsumexp; entries
 * that are 0 drop out).  HOST pointers, copied.  Required before a call with
 * GWIN_MODE_MARG_DIST / GWIN_MODE_MARG_DIST_PHASE; n = 0 clears it.
 */
int gwin_plan_set_distance_marg(gwin_plan* plan, int64_t n, const double* dist, const double* weight);

/*
 * Calibration model == gwin/calibration.py:104-173 (CubicSpline) as the waveform generator
 * applies it to every detector's template ([UPSTREAM] pycbc FDomainDetFrameGenerator
 * ``recalib=``, handed over at base_data.py:225-230):
 *     h_det(f) *= (1 + dA(f)) (2 + i dphi(f)) / (2 - i dphi(f))
 * with dA, dphi = scipy UnivariateSpline(spline_points, values)(f) at its default k = 3,
 * s = n_points -- which for calibration-sized values (sum of squares <= n_points) is the
 * least-squares CUBIC POLYNOMIAL through the values, evaluated (and extrapolated) at every
 * bin.  spline_freqs: HOST [n_det][n_points] = CubicSpline.spline_points per detector
 * (plan detector order), 4 <= n_points <= GWIN_MAX_CALIB_POINTS; n_points = 0 clears.
 * Builds the values->coefficients maps and seven prefix tables per detector for the
 * closed-form <h|h> (|1 + dA|^2 is a sextic in f).
 */
#define GWIN_MAX_CALIB_POINTS 32
int gwin_plan_set_calibration(gwin_plan* plan, int n_points, const double* spline_freqs);

/*
 * One batched evaluation of GaussianNoise._loglr / MarginalizedGaussianNoise._loglr
 * for W parameter points (what pool.map(CallModel, points) computes one point at a
 * time in the reference).  All pointers are DEVICE pointers on the plan's device.
 *
 *   params  [W][GWIN_NPARAM] row-major:
 *           mass1, mass2 [Msun], spin1z, spin2z, distance [Mpc], inclination,
 *           coa_phase, ra, dec, polarization [rad], tc [GPS s],
 *           lambda1, lambda2 (dimensionless tidal deformabilities; 0 for black holes)
 *   skip    [W] or NULL; nonzero = prior was -inf, do not evaluate
 *           (base.py:556-568 / base_data.py:128-141 short-circuit): loglr=-inf
 *   loglr   [W]
 *   hd      [W][n_det][2] complex <d|h> per detector ({det}_cplx_loglr + hh/2, or
 *           {det}_matchedfilter_snrsq); may be NULL
 *   hh      [W][n_det] <h|h> per detector ({det}_optimal_snrsq); may be NULL
 *   stream  cudaStream_t (0 = legacy default stream).  Asynchronous.
 *
 * Calls on one plan share its workspace; a call on another stream than the previous call waits for
 * it on the device (cudaStreamWaitEvent), so callers need no ordering of their own.
 * Points with non-finite / non-physical parameters return loglr = -inf, never NaN.
 */
int gwin_loglr_batch(gwin_plan* plan, int mode, int64_t W,
                     const double* params, const uint8_t* skip,
                     double* loglr, double* hd, double* hh, void* stream);

/* Same with an explicit parameter layout: element (walker w, parameter i) is
 * params[w*stride_walker + i*stride_param].  (13, 1) is the row layout above; (1, W) is
 * parameter-major [13][W], which a host that holds one array per parameter (as gwin's
 * samplers do) can fill without a transpose and which the device reads coalesced. */
int gwin_loglr_batch_strided(gwin_plan* plan, int mode, int64_t W,
                             const double* params, int64_t stride_walker, int64_t stride_param,
                             const uint8_t* skip,
                             double* loglr, double* hd, double* hh, void* stream);

/* Same, with HOST buffers: pinned staging, H2D, kernels, D2H, synchronise. */
int gwin_loglr_batch_host(gwin_plan* plan, int mode, int64_t W,
                          const double* params, const uint8_t* skip,
                          double* loglr, double* hd, double* hh);

/* Host buffers with layout; params must be a dense [W][13] or [13][W] block. */
int gwin_loglr_batch_host_strided(gwin_plan* plan, int mode, int64_t W,
                                  const double* params, int64_t stride_walker, int64_t stride_param,
                                  const uint8_t* skip,
                                  double* loglr, double* hd, double* hh);

/*
 * Host entry with a column map, for callers that hold the SAMPLED parameters only (what a sampler hands
 * to pool.map: one row of model.sampling_params values per walker, models/__init__.py:101-110).
 *   points  HOST, dense [W][ndim] (stride_walker = ndim, stride_col = 1) or [ndim][W] (1, W)
 *   map     [GWIN_NPARAM]: for every kernel parameter the column of `points` that holds it, or -1
 *   fixed   [GWIN_NPARAM]: the value of every parameter whose map entry is -1 (static parameters)
 * The rows of gwin_loglr_batch are assembled on the device.  Buffers that are page-locked
 * (gwin_host_alloc, cudaHostAlloc, a pinned torch tensor) are copied from / into directly; pageable
 * ones go through the plan's staging buffers.  hd / hh may be NULL (then they are not copied back).
 */
int gwin_loglr_batch_points_host(gwin_plan* plan, int mode, int64_t W,
                                 const double* points, int ndim, int64_t stride_walker, int64_t stride_col,
                                 const int* map, const double* fixed, const uint8_t* skip,
                                 double* loglr, double* hd, double* hh);
/*
 * The same call in two halves, for callers with more than one batch to evaluate -- what the reference
 * gets from a process pool working on several tasks at once (pool.map of CallModel over the points,
 * gwin/option_utils.py:171-185; the separate prior / likelihood maps of a parallel-tempered ensemble,
 * gwin/sampler/emcee.py:230-250): _submit queues the
 * host-to-device copy, the kernels and the copy back and returns at once with a ticket; _wait blocks until
 * that call's results are in loglr / hd / hh.  Two calls may be in flight per plan: the copies of one
 * overlap the kernels of the other (separate copy streams, separate staging).  points / skip may be reused
 * as soon as _submit returns unless points is page-locked (then it is read until the copy has run: keep it
 * unchanged until _wait); loglr / hd / hh must stay valid until _wait.  Tickets are 0 and 1.
 */
int gwin_loglr_batch_points_submit(gwin_plan* plan, int mode, int64_t W,
                                   const double* points, int ndim, int64_t stride_walker, int64_t stride_col,
                                   const int* map, const double* fixed, const uint8_t* skip,
                                   double* loglr, double* hd, double* hh, int* ticket);
int gwin_loglr_batch_points_wait(gwin_plan* plan, int ticket);
/* Page-locked host memory for the buffers of the host entry points. */
int gwin_host_alloc(void** ptr, size_t bytes);
void gwin_host_free(void* ptr);

/*
 * FDomainDetFrameGenerator.generate (the call at gaussian_noise.py:322) for ONE
 * parameter point: writes the detector-frame waveforms h_det(f_k), un-whitened,
 * to a host buffer [n_det][n_f][2]; bins at and beyond the waveform's own length
 * are zero.  *length receives that length (len(h) in gaussian_noise.py:328).
 */
int gwin_generate_host(gwin_plan* plan, const double* params, double* h_out,
                       int64_t* length);

/*
 * The same calls with per-walker calibration values (what Recalibrate.map_to_adjust,
 * calibration.py:50-76, collects from the ``recalib_*`` parameters):
 *   calib   value v of walker w at calib[w * cstride_walker + v * cstride_value],
 *           v = (2 * det + which) * n_points + i, which = 0 for recalib_amplitude_<ifo>_<i>,
 *           1 for recalib_phase_<ifo>_<i>.  NULL = no calibration (the plain calls above).
 * gwin_loglr_batch_calib takes DEVICE pointers, the _host variants HOST pointers (dense
 * [W][n_det * 2 * n_points] or its transpose).  Non-finite values give loglr = -inf.
 */
int gwin_loglr_batch_calib(gwin_plan* plan, int mode, int64_t W,
                           const double* params, int64_t stride_walker, int64_t stride_param,
                           const double* calib, int64_t cstride_walker, int64_t cstride_value,
                           const uint8_t* skip,
                           double* loglr, double* hd, double* hh, void* stream);
int gwin_loglr_batch_calib_host(gwin_plan* plan, int mode, int64_t W,
                                const double* params, int64_t stride_walker, int64_t stride_param,
                                const double* calib, int64_t cstride_walker, int64_t cstride_value,
                                const uint8_t* skip,
                                double* loglr, double* hd, double* hh);
int gwin_generate_calib_host(gwin_plan* plan, const double* params, const double* calib,
                             double* h_out, int64_t* length);

/* Kernel launches issued by the last gwin_loglr_batch* call on this plan. */
int gwin_plan_last_launches(const gwin_plan* plan);

/* Device-side timing of the dominant (loglr) kernel with CUDA events recorded on the
 * launching stream.  enable != 0 switches it on; every call returns the mean duration
 * [ms] and count of the loglr-kernel launches since the previous call, then resets. */
int gwin_plan_kernel_timing(gwin_plan* plan, int enable, double* mean_ms, int* n_launches);
/* The same bracket split by launch (means over the launches the last gwin_plan_kernel_timing call reported):
 * [0] recurrence (or direct) kernel over its share of the band, [1] table kernel over the schedule's
 * sub-blocks, [2] table kernel over the tails' fine sub-blocks, [3] recurrence kernel over what is left. */
int gwin_plan_kernel_timing_detail(const gwin_plan* plan, double* mean_ms4);
/* The schedule of the moment tables: start bin, length, reference chirps and finest tail sub-block (0: none)
 * of up to `cap` sub-blocks; *n receives their number. */
int gwin_plan_table_blocks(const gwin_plan* plan, int64_t cap, int32_t* start, int32_t* m, int32_t* n_refs,
                           int32_t* fine_min, int64_t* n);

/* ------------------------------------------------------------------------------------------
 * Device-resident ensemble step (replaces, on the device, what emcee does on the host between
 * two pool.map calls -- gwin/sampler/emcee.py:74-76, 286-289; semantics of emcee 2.x, SURVEY.md
 * appendix B).  All buffers are DEVICE pointers owned by the caller; all calls are asynchronous on
 * `stream`.  Random numbers are Philox4x32-10(seed; iteration, stage, index): every rank of a
 * multi-GPU run regenerates identical numbers, so only log-likelihoods need to be exchanged.
 * ------------------------------------------------------------------------------------------ */
#define GWIN_ENS_MAXDIM 64

#define GWIN_PRIOR_UNIFORM        0   /* pycbc.distributions.Uniform                          */
#define GWIN_PRIOR_UNIFORM_ANGLE  1   /* UniformAngle: cyclic on [lo, hi)                      */
#define GWIN_PRIOR_SIN_ANGLE      2   /* SinAngle: p ~ sin x, reflected into [0, pi]           */
#define GWIN_PRIOR_COS_ANGLE      3   /* CosAngle: p ~ cos x, reflected into [-pi/2, pi/2]     */
#define GWIN_PRIOR_UNIFORM_RADIUS 4   /* UniformRadius: p ~ x^2                                */
#define GWIN_PRIOR_GAUSSIAN       5   /* Gaussian: normal (mu, sigma) truncated to [lo, hi]    */
#define GWIN_PRIOR_NONE           6   /* no prior of its own (a sampling parameter whose prior is on derived ones) */

#define GWIN_TARGET_MCHIRP 100        /* sampling dim is mchirp ...                            */
#define GWIN_TARGET_Q      101        /* ... / q = m1/m2: mapped to mass1, mass2 on the device */
#define GWIN_TARGET_CALIB0 200        /* + v: calibration value v of gwin_loglr_batch_calib (the recalib_* parameters) */

typedef struct gwin_prior_spec {
    int32_t ndim;                       /* sampling dimensions                                   */
    int32_t type[GWIN_ENS_MAXDIM];      /* GWIN_PRIOR_*                                          */
    int32_t target[GWIN_ENS_MAXDIM];    /* waveform-parameter column 0..12, or GWIN_TARGET_*     */
    double lo[GWIN_ENS_MAXDIM], hi[GWIN_ENS_MAXDIM];
    double mu[GWIN_ENS_MAXDIM], sigma[GWIN_ENS_MAXDIM];    /* GWIN_PRIOR_GAUSSIAN                     */
    double fixed[GWIN_NPARAM];          /* static values of the parameters that are not sampled  */
    /* Sampling transform (mass1, mass2) -> (mchirp, q) (gwin/models/base.py:109-220 with
     * pycbc.transforms.Mass1Mass2ToMchirpQ): the walkers move in mchirp, q (dimensions with target
     * GWIN_TARGET_MCHIRP / _Q and type GWIN_PRIOR_NONE), the prior is on mass1, mass2 (derived_*[0], [1]:
     * type, bounds, Gaussian parameters) and log |d(mass1, mass2)/d(mchirp, q)| is added.  0 = none. */
    int32_t derived_masses;
    int32_t derived_type[2];
    double derived_lo[2], derived_hi[2], derived_mu[2], derived_sigma[2];
} gwin_prior_spec;

/* Stretch-move proposals for one half (0/1) of every temperature.
 *   p [ntemps][nwalkers][ndim] -> q [ntemps][nwalkers/2][ndim], logfac [ntemps][nwalkers/2]
 *   move 0: EnsembleSampler (z = ((a-1)u+1)^2/a, logfac = (ndim-1) ln z), halves are contiguous
 *   move 1: PTSampler       (ln z ~ U(-ln a, ln a),  logfac = ndim ln z), halves are interleaved */
int gwin_ens_propose(uint64_t seed, uint32_t iter, const uint32_t* iter_dev,
                     int half, int ntemps, int nwalkers, int ndim,
                     double a, int move, int interleaved,
                     const double* p, double* q, double* logfac, void* stream);

/* Boundary conditions, log prior and the likelihood kernel's parameter rows for W points:
 * q [W][ndim] -> logp [W], skip [W] (1 where logp = -inf), params [13][W] (parameter-major, feed to
 * gwin_loglr_batch_strided with strides (1, W)).  Replaces _transform_params + logprior
 * (gwin/models/base.py:543-554, 602-621) for the supported distributions. */
int gwin_ens_prior(const gwin_prior_spec* spec, int64_t W, const double* q,
                   double* logp, uint8_t* skip, double* params, void* stream);
/* The same for a calibrated model: sampling dimensions with target GWIN_TARGET_CALIB0 + v also write
 * calib[v * W + w] (value-major [n_calib][W]: feed to gwin_loglr_batch_calib with strides (1, W)).  Values no
 * sampling dimension writes stay as the caller put them (static recalib_* parameters: fill them once). */
int gwin_ens_prior_calib(const gwin_prior_spec* spec, int64_t W, const double* q,
                         double* logp, uint8_t* skip, double* params,
                         double* calib, int n_calib, void* stream);

/* Accept/reject: ln u < logfac + (beta_k q_logl + q_logp) - lnprob; updates p, logl, lnprob
 * ([ntemps][nwalkers]...) and the per-walker acceptance counters. */
int gwin_ens_accept(uint64_t seed, uint32_t iter, const uint32_t* iter_dev,
                    int half, int ntemps, int nwalkers, int ndim,
                    int interleaved, const double* betas, const double* q, const double* logfac,
                    const double* q_logl, const double* q_logp,
                    double* p, double* logl, double* lnprob, int* naccept, void* stream);

/* Swap attempt between temperature `level` and `level-1` for the walker pairs
 * (iperm[j], i1perm[j]) (emcee PTSampler._temperature_swaps). */
int gwin_ens_swap(uint64_t seed, uint32_t iter, const uint32_t* iter_dev, int level, int nwalkers, int ndim,
                  const double* betas, const int* iperm, const int* i1perm,
                  double* p, double* logl, double* lnprob, int* nswap_accepted, void* stream);

/* All swap levels ntemps-1 .. 1 of one iteration in a single launch; perms is
 * [ntemps-1][2][nwalkers] (entry ntemps-1-level holds iperm, i1perm of that level); with a
 * device counter the kernel reads the block perms + *iter_dev * perm_stride. */
int gwin_ens_swap_all(uint64_t seed, uint32_t iter, const uint32_t* iter_dev,
                      int ntemps, int nwalkers, int ndim,
                      const double* betas, const int* perms, int64_t perm_stride,
                      double* p, double* logl, double* lnprob, int* nswap_accepted, void* stream);

/* Sampler iteration as ONE CUDA graph: every gwin_ens_* call above takes, besides the scalar
 * iteration `iter`, an optional DEVICE counter iter_dev (NULL = 0); the Philox key uses
 * iter + *iter_dev.  With the counter the argument lists of one iteration no longer change, so
 * the whole iteration (proposals, priors, likelihood launches, all-gather, accept, swaps,
 * history) can be captured once and replayed; gwin_ens_tick advances the counter at its end.
 * gwin_ens_record writes row it + *it_dev of the histories chain [W][niter][ndim],
 * lnp_h / lnl_h [W][niter] (what emcee's sample() appends per iteration). */
int gwin_ens_record(int64_t W, int ndim, int64_t niter, uint32_t it, const uint32_t* it_dev,
                    const double* p, const double* lnprob, const double* logl,
                    double* chain, double* lnp_h, double* lnl_h, void* stream);
int gwin_ens_tick(uint32_t* counter, void* stream);

/* FP64 issue-rate micro-benchmark on the plan's device: runs a dependent-chain-free
 * DFMA loop and returns achieved FP64 FMA lane-ops per second. */
int gwin_fp64_peak(int device, double* dfma_per_sec);

/* Diagnostic variants of the probe (FP64 instructions x 32 lanes per second):
 * 0 = gwin_fp64_peak; 1 = DFMAs whose three source registers are all private to the
 * instruction (no operand-reuse cache help); 2 = the hot loop's pattern, independent
 * complex multiplications (2 DMUL + 2 DFMA each). */
int gwin_fp64_probe(int device, int variant, double* ops_per_sec);

const char* gwin_last_error(void);
const char* gwin_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GWIN_B200_H */
