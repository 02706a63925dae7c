// gwin_b200 -- batched TaylorF2 log-likelihood-ratio on B200 (sm_100a).
//
// Replaces, for a whole ensemble of walkers at once, the per-walker chain the
// reference runs on the CPU:
//   waveform_generator.generate(**params)      gaussian_noise.py:322
//   h[kmin:kmax] *= weight                     gaussian_noise.py:337
//   <d|h> = data.inner(h)                      gaussian_noise.py:339
//   <h|h> = h.inner(h).real                    gaussian_noise.py:340
//   loglr combine                              gaussian_noise.py:341-350
//   marginalised-phase combine                 marginalized_gaussian_noise.py:456-511
//   distance(-phase) marginalisation           marginalized_gaussian_noise.py:431-448
//   calibration of the template                calibration.py:142-173
// No intermediate waveform array is ever materialised: one thread owns one walker and the
// frequency axis is cut into chunks so that the grid fills all 148 SMs.
//
// Kernels per batch (long segments, the path the benchmark takes):
//   sort_key -> bucket_order (walkers ordered by chirp coefficient, tails by last bin) -> walker_setup ->
//   tail_plan -> loglr_table<FINE=0> over the sub-blocks of the schedule -> loglr_table<FINE=1> over the
//   tails (fine sub-blocks found in closed form + the last bins) -> loglr_recur over whatever is left ->
//   combine -> slow_compact / loglr_direct / slow_finish for walkers outside the tables' coverage
//   (-> marg_dist).
// Short segments and small batches: loglr_recur over the whole band.  loglr_direct is an independent
// second implementation (and the slow path).  Calibrated batches take the CALIB variants of all three.
//
// The table kernel replaces the per-bin work by interpolation in per-plan moment
// tables of the data relative to a grid of reference chirps: 1299 FP64 operations per (walker,
// sub-block of up to 8192 bins); its taps x moments product runs on the FP64 tensor cores
// (mma.m8n8k4.f64: same pipe and peak as DFMA on B200, a fifth of the instructions).  The recurrence
// kernel is bound by the register-file bandwidth that feeds the FP64 pipe; its tables
// are L2-resident, staged into shared memory per warp by TMA bulk copies and read with warp-uniform
// 128-bit loads.  There is no dense contraction anywhere else on this path.
#include <cuda_runtime.h>
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/gwin_b200.h"

// ---------------------------------------------------------------------------
// constants (LAL values; provenance in DESIGN.md)
// ---------------------------------------------------------------------------
#define G_PI      3.141592653589793238462643383279502884
#define G_TWOPI   6.283185307179586476925286766559005768
#define G_GAMMA   0.577215664901532860606512090082402431
#define G_C_SI    299792458.0
#define G_MTSUN   4.925491025543575903411922162094833998e-6
#define G_MRSUN   1.476625061404649406193430731479084713e3
#define G_PC_SI   3.085677581491367278913937957796471611e16

#define WARP 32
#ifndef RECUR_UNROLL
#define RECUR_UNROLL 4        // bins per trip of the recurrence hot loop
#endif
static constexpr int kRecurUnroll = RECUR_UNROLL;
#ifndef CTA_THREADS
#define CTA_THREADS 128
#endif
#define RESYNC 64            // direct kernel: detector phasor re-sync interval
#ifndef MAX_BLOCK
#define MAX_BLOCK 256        // recurrence kernel: longest polynomial block
#endif
#define MIN_BLOCK 8

// per-walker coefficient rows (SoA: wc[row * wpad + walker])
enum {
    C_A0 = 0, C_A2, C_A3, C_A4, C_A5, C_B5, C_A6, C_B6, C_A7, C_A10, C_A12,   // phase [turns]
    C_DET0                                                        // then 4 per detector
};
enum { D_TAUHI = 0, D_TAULO, D_ARE, D_AIM, D_ECOS, D_ESIN,               // E = e^{+2 pi i tau}
       D_CA0, D_CA1, D_CA2, D_CA3, D_CP0, D_CP1, D_CP2, D_CP3,           // calibration cubics (amplitude, phase) in z
       D_NROW };
#define NCOEF(nd) (C_DET0 + D_NROW * (nd))

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}
// shared with gwin_ensemble.cu
int gwin_fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}
// GWIN_DEBUG=1: progress of the (rare, slow) plan-level steps on stderr
static bool dbg_on() { static const bool on = getenv("GWIN_DEBUG") && atoi(getenv("GWIN_DEBUG")) != 0; return on; }
#define DBG(...) do { if (dbg_on()) { fprintf(stderr, "[gwin] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(GWIN_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)

struct PlanDev {
    int n_det;
    int n_f;
    int kmin, kmax;          // likelihood band [kmin, kmax)
    int istart;              // first non-zero waveform bin
    double delta_f;
    double f_lower_wf;       // waveform generator's f_lower
    double epoch;
    double gps_utc;          // GPS-UTC leap seconds at the epoch
    double resp[GWIN_MAX_DET][9];
    double loc[GWIN_MAX_DET][3];
    const double4* basis;    // [n_f] {f^(1/3), f^(-5/3), ln f^(1/3), 0}
    const double2* T;        // [n_f][n_det] norm*conj(d)*f^(-7/6)/S
    const double*  C;        // [n_det][n_f+1] prefix sums of norm*f^(-7/3)/S
    // calibration (gwin_plan_set_calibration); n_cal = 0: none
    int n_cal;               // spline points per detector
    double cal_zs[GWIN_MAX_DET], cal_z0[GWIN_MAX_DET];   // z = k * zs + z0 maps the spline range to [-1, 1]
    const double* calM;      // [n_det][4][n_cal] values -> least-squares cubic coefficients in z
    const double* Ccal;      // [n_det][7][n_f+1] prefix sums of norm*f^(-7/3)/S * z^j
};

// Per distinct (block length, degree): node offsets and the linear map from the exact
// phase samples to the polynomial's forward differences.  Built on the host in extended
// precision; read with warp-uniform loads.
struct BlockMat {
    double m[16];   // D_r = sum_i m[4r+i] (f_{i+1} - f_0):  D1(0), D2(1), D3(2), D4
    int node[4];    // offsets of nodes 1..deg from the block start (node 0 is 0, node deg is B-1)
    int deg;        // 4 = quartic, 3 = cubic, 0 = evaluate the bins directly
    int pad[3];
};

// Table kernel: sub-block moments of the data looked up by a type-2 NUFFT
#define TAB_DEG 8            // degree of the Taylor polynomial of e^{-2 pi i R(j)}
#define TAB_NT 9             // Taylor orders of the phase kept per sub-block: s^0 .. s^8, s = j / (m/2)
#define TAB_NTP 10           // ... padded to an even count (128-bit shared-memory reads)
#define TAB_NTH 13           // orders the host-side bound of the schedule looks at
#define TAB_W 13             // taps of the exponential-of-semicircle interpolation kernel
#define TAB_BETA (2.30 * TAB_W)
#ifndef TAB_MIN
#define TAB_MIN 128          // shortest / longest sub-block [bins]; lengths are 2^k and 3 2^(k-1)
#endif
#ifndef TAB_MAX
#define TAB_MAX 8192
#endif
#define TAB_NM TAB_DEG       // moments stored per grid point: n = 0, 2, 3, .. TAB_DEG (e_1 = 0: nu carries the slope)
// truncation of e^{-2 pi i R}: the schedule is built so that the majorant of the dropped terms stays
// below TAB_EST_DESIGN for every covered walker; a walker whose own estimate exceeds TAB_EST_MAX (or
// whose |2 pi R| exceeds TAB_RHO_MAX) leaves the table path for the slow path (direct kernel)
#define TAB_EST_DESIGN 2.0e-9
#define TAB_EST_MAX 1.6e-8
#define TAB_RHO_DESIGN 0.05
#define TAB_RHO_MAX 0.10
#define TAB_KEY_BITS 20      // walkers are ordered by their Newtonian chirp coefficient, quantised over the covered range
#include "tab_poly.inc"      // TAB_WEIGHTS(v, w): the interpolation kernel's per-tap polynomials, literal coefficients

// One sub-block of the table schedule.  Its K reference chirps sit at a_lo + (r + 1/2) da in the
// Newtonian coefficient a0 (turns Hz^(5/3)); tables of reference r and detector d start at
// off + ((r n_det + d) (2 m + TAB_W)) TAB_NM double2.
struct TabBlk {
    int start, m;
    int K, pair0;            // pair0: index of (this sub-block, reference 0) in the per-pair arrays
    double a_lo, da, inv_da;
    double eps;              // (m/2) / centre bin
    long long off;
    // tails: a sub-block in which walkers may end carries a tree of fine sub-blocks (its halves, their
    // halves, ... down to fine_min bins), entries [fine_first, fine_first + fine_n) of the block array;
    // a fine sub-block knows its parent's start
    int fine_first, fine_n;
    int fine_min, parent_start;
};

// What the tables are built for (gwin_plan_cover_*): the range of the Newtonian coefficient and, for
// every other phase coefficient, the largest deviation of a walker from the equal-mass, non-spinning,
// point-particle reference chirp of the same a0.
struct Cover {
    bool valid;
    double a_lo, a_hi;
    double dev[16];
    double ke_lo, ke_hi;     // range of the walkers' last bins: where the fine sub-blocks of the tails are built
};

struct gwin_plan {
    int device;
    PlanDev d;
    double norm;
    double lognl_det[GWIN_MAX_DET];
    double lognl;
    // tuning
    int kernel;
    double phase_tol;
    double mchirp_min;
    bool auto_mchirp;        // host API: derive mchirp_min from each batch
    // recurrence schedule
    std::vector<int> blk_start;      // host copy, nblk+1 entries
    std::vector<int> blk_mat;        // nblk entries: index into mats
    std::vector<BlockMat> mats;
    int* d_blk_start;
    int* d_blk_mat;
    BlockMat* d_mats;
    int n_blk;
    bool sched_dirty;
    // workspace
    int64_t w_cap;
    int chunk_cap;
    double* d_wc;
    int *d_ks, *d_ke, *d_key_in, *d_key_out, *d_perm_in, *d_perm_out;
    double2* d_partial;
    size_t partial_cap;
    void* d_cub; size_t cub_bytes;
    // calibration: host copy of the per-bin <h|h> weights norm f^(-7/3)/S (for the z^j prefix tables)
    std::vector<double> w73;
    double *d_calM, *d_Ccal;
    double *h_calib, *d_calib; size_t calib_cap;
    // table kernel: sub-block schedule above k_split, reference chirps, moment tables
    Cover cover;             // what the current tables cover
    Cover want;              // what the next build should cover (gwin_plan_cover_*, automatic re-planning)
    bool cover_manual;       // coverage was declared by the caller: no automatic re-planning
    std::vector<TabBlk> tblk;
    int n_tblk, n_fine, n_pairs, k_split, n_blk_low, k_tab_end;      // n_tblk sub-blocks of the schedule, then n_fine fine ones
    size_t tab_bytes;
    bool tab_dirty;
    bool tab_worth;          // AUTO uses the tables only where they pay (long sub-blocks, enough of them)
    TabBlk* d_tblk;
    double *d_cmat, *d_tref, *d_invhat;
    double2* d_tab;
    double2* d_partialT; size_t partialT_cap;
    // walker order bookkeeping of the table path
    int *d_ke_fast, *d_ke_low, *d_ks_tail, *d_key2_in, *d_key2_out, *d_idx2_in, *d_perm_e, *d_grp_b;
    int *d_key3_in, *d_perm_t, *d_col_t, *d_ks_fine, *d_ks_rec, *d_tail_parent, *d_grp_p;
    double2* d_partialF; size_t partialF_cap;
    int tab_split;           // sub-blocks from here on have at most two reference chirps: taken in time order
    // slow path: walkers the fast kernels cannot take (outside the schedule / the tables' coverage)
    unsigned char* d_slow;
    int *d_slow_list, *d_slow_n;         // d_slow_n[0]: this batch, [1]: since the last query
    double2* d_partialS; size_t partialS_cap;
    unsigned long long* d_stats;         // cover_stats_kernel output
    bool replan;                         // re-measure the coverage on the next batch
    // ordering of calls that arrive on different streams (they share the workspace)
    cudaEvent_t ev_done; bool ev_done_set; cudaStream_t last_stream;
    // distance marginalisation grid (1/D_j, log w_j) and the per-walker (x, opt) hand-over
    int n_marg;
    double *d_marg_u, *d_marg_logw, *d_marg_opt;
    // staging for the host API
    double *h_params, *h_out; uint8_t* h_skip;
    double *d_params, *d_out, *d_points; uint8_t* d_skip;
    int64_t stage_cap;
    cudaStream_t stream;
    // asynchronous host entry (gwin_loglr_batch_points_submit / _wait): two calls in flight, each with its own
    // staging; copies in and out run on their own streams so that they overlap the other call's kernels
    struct AsyncSlot {
        double *d_points, *d_params, *d_out, *h_points, *h_out; uint8_t *d_skip, *h_skip;
        int64_t cap;
        bool busy, table;
        int64_t W; double *loglr, *hd, *hh; bool direct_out;
        cudaEvent_t ev_in, ev_k, ev_out;
    } aslot[2];
    cudaStream_t stream_in, stream_out;
    bool async_ready;
    int last_launches;
    // optional timing of the dominant kernel on the launching stream
    bool timing;
    cudaEvent_t ev0, ev1;
    cudaEvent_t evd[3];          // between the launches of the bracket: after the main recurrence / direct launch, the table launches, the fine tails
    double detail_ms_sum[4], detail_last[4];
    double kernel_ms_sum;
    int kernel_ms_n;
    bool ev_pending;
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double round_magic(double x) {
    // round-to-nearest-integer for |x| < 2^51 on the FP64 add pipe
    const double M = 6755399441055744.0;
    return __dadd_rn(__dadd_rn(x, M), -M);
}
__device__ __forceinline__ double frac1(double x) { return x - round_magic(x); }

// sin / cos of 2 pi y for y in turns, |y| < 2^51: exact reduction to an octant
// (y - rint(y), then quarter turns), fdlibm's kernel polynomials on |x| <= pi/4 (error < 1 ulp),
// quadrant fix-up by selects.  ~24 FP64 instructions, no branches, no slow path -- the library
// sincospi() costs ~55 and the recurrence kernel calls this 5 times per block.
__device__ __forceinline__ void sincos_turns(double y, double& sn, double& cs) {
    const double r = y - round_magic(y);                  // [-1/2, 1/2]
    const double q = round_magic(4.0 * r);                // -2 .. 2 quarter turns
    const double x = fma(-0.25, q, r) * G_TWOPI;          // [-pi/4, pi/4]
    const double t = x * x;
    double ps = fma(t, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(t, ps, 2.75573137070700676789e-06);
    ps = fma(t, ps, -1.98412698298579493134e-04);
    ps = fma(t, ps, 8.33333333332248946124e-03);
    ps = fma(t, ps, -1.66666666666666324348e-01);
    const double s0 = fma(x * t, ps, x);
    double pc = fma(t, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(t, pc, -2.75573143513906633035e-07);
    pc = fma(t, pc, 2.48015872894767294178e-05);
    pc = fma(t, pc, -1.38888888888741095749e-03);
    pc = fma(t, pc, 4.16666666666666019037e-02);
    const double c0 = fma(t * t, pc, fma(-0.5, t, 1.0));
    const int qi = (int)q & 3;
    const double a = (qi & 1) ? s0 : c0, b = (qi & 1) ? c0 : s0;
    cs = (qi == 1 || qi == 2) ? -a : a;
    sn = (qi >= 2) ? -b : b;
}
// e^{-2 pi i y} for y in turns: returns (cos, -sin)
__device__ __forceinline__ double2 phasor_neg(double y) {
    double s, c;
    sincos_turns(y, s, c);
    return make_double2(c, -s);
}
__device__ __forceinline__ double2 phasor_pos(double y) {
    double s, c;
    sincos_turns(y, s, c);
    return make_double2(c, s);
}
// e^{+2 pi i y}: Taylor series when the angle is small (the higher forward differences of
// the chirp phase are ~1e-3 .. 1e-9 turns), full sincospi otherwise.
__device__ __forceinline__ double2 phasor_pos_small(double y) {
    if (fabs(y) < 0.0125) {                      // |theta| < 0.0785 rad: next term < 3e-17
        const double th = G_TWOPI * y, t2 = th * th;
        const double c = fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 40320.0, -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
        const double s = th * fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 362880.0, -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
        return make_double2(c, s);
    }
    return phasor_pos(y);
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a*b + c
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// One Horner step of the recurrence hot loop: acc <- acc g + T ; g <- g r.
// (Emitting these eight instructions as inline PTX in an order where each shares a source
// register with its predecessor does not survive ptxas, which re-schedules them; the
// modelled operand-fetch cost stays at ~82 % of the DFMA rate either way, tools/sass_cost.py.)
// FP64 tensor-core MMA, D (8x8) += A (8x4, row-major) B (4x8, column-major): lane l holds A[l/4][l%4], B[l%4][l/4],
// D[l/4][2 (l%4)] and D[l/4][2 (l%4) + 1].  Same pipe and peak rate as DFMA on B200 (tools/fp64_probe.py) at one
// instruction and four operand registers per 8 multiply-adds per lane.
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void horner_step(double2& acc, double2& g, const double2 T, const double2 r) {
    acc = cfma(acc, g, T);
    g = cmul(g, r);
}
__device__ __forceinline__ void phasor_step(double2& r, const double2 s) { r = cmul(r, s); }

__device__ __forceinline__ double4 ld_basis(const double4* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// Calibration factor of one detector (gwin/calibration.py:167-170):
//   (1 + dA) (2 + i dphi) / (2 - i dphi),  dA, dphi cubic polynomials in z
struct CalCoef { double a[4], p[4]; };
__device__ __forceinline__ double2 calib_factor(const CalCoef& c, double z) {
    const double A = fma(fma(fma(c.a[3], z, c.a[2]), z, c.a[1]), z, c.a[0]);
    const double P = fma(fma(fma(c.p[3], z, c.p[2]), z, c.p[1]), z, c.p[0]);
    const double P2 = P * P;
    const double sc = (1.0 + A) / (4.0 + P2);
    return make_double2(sc * (4.0 - P2), sc * (4.0 * P));
}
// log of the factor, split: re = log(1 + dA), im = 2 atan(dphi / 2) [rad]; amp = 1 + dA
__device__ __forceinline__ double2 calib_log(const CalCoef& c, double z, double& amp) {
    const double A = fma(fma(fma(c.a[3], z, c.a[2]), z, c.a[1]), z, c.a[0]);
    const double P = fma(fma(fma(c.p[3], z, c.p[2]), z, c.p[1]), z, c.p[0]);
    amp = 1.0 + A;
    return make_double2(log1p(A), 2.0 * atan(0.5 * P));
}
__device__ __forceinline__ void load_cal_coef(CalCoef& c, const double* row /* detector's first row + walker */, int wpad) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c.a[i] = row[(size_t)(D_CA0 + i) * wpad];
        c.p[i] = row[(size_t)(D_CP0 + i) * wpad];
    }
}

// TaylorF2 phase in turns at one bin, basis = {x, x^-5, ln x}
struct PhaseCoef { double a0, a2, a3, a4, a5, b5, a6, b6, a7, a10, a12; };
#define NPH 11                 // phase-coefficient rows
__device__ __forceinline__ double phase_turns(const PhaseCoef& c, double4 b) {
    const double x = b.x, xm5 = b.y, lx = b.z;
    double lo = fma(fma(c.a4, x, c.a3), x, c.a2);
    lo = fma(lo, x * x, c.a0) * xm5;
    const double l5 = fma(c.b5, lx, c.a5);
    const double l6 = fma(c.b6, lx, c.a6);
    const double x2 = x * x;
    const double tid = fma(c.a12, x2, c.a10) * (x2 * x);        // tidal: a10 x^5 + a12 x^7 = x^2 (x^3 (...))
    const double hi = fma(fma(c.a7 + tid, x, l6), x, l5);
    return lo + hi;
}

// TaylorF2 phase coefficients of one walker (lalsimulation XLALSimInspiralPNPhasing_F2, 3.5PN with
// aligned spins, plus the 5PN/6PN tidal terms), folded with powers of (pi M)^(1/3) so that
//   Psi(f) [turns] = a0 x^-5 + a2 x^-3 + a3 x^-2 + a4 x^-1 + a5 + b5 ln x + a6 x + b6 x ln x + a7 x^2
//                    + a10 x^5 + a12 x^7,   x = f^(1/3)
// out[] in the order of the C_A* rows.  Returns (pi M)^(1/3).
__host__ __device__ inline double pn_phase_coefs(double m1, double m2, double chi1, double chi2, double phic,
                                                 double lam1, double lam2, double* out) {
    const double mtot = m1 + m2;
    const double dm = (m1 - m2) / (m1 + m2);
    const double eta = m1 * m2 / mtot / mtot;
    const double m1M = m1 / mtot, m2M = m2 / mtot;
    const double pfaN = 3.0 / (128.0 * eta);
    double v2 = 5.0 * (743.0 / 84.0 + 11.0 * eta) / 9.0;
    double v3 = -16.0 * G_PI;
    double v4 = 5.0 * (3058.673 / 7.056 + 5429.0 / 7.0 * eta + 617.0 * eta * eta) / 72.0;
    double v5 = 5.0 / 9.0 * (7729.0 / 84.0 - 13.0 * eta) * G_PI;
    double vl5 = 5.0 / 3.0 * (7729.0 / 84.0 - 13.0 * eta) * G_PI;
    double v6 = (11583.231236531 / 4.694215680 - 640.0 / 3.0 * G_PI * G_PI - 6848.0 / 21.0 * G_GAMMA)
              + eta * (-15737.765635 / 3.048192 + 2255.0 / 12.0 * G_PI * G_PI)
              + eta * eta * 76055.0 / 1728.0 - eta * eta * eta * 127825.0 / 1296.0;
    v6 += (-6848.0 / 21.0) * 1.3862943611198906188;   // log(4)
    const double vl6 = -6848.0 / 21.0;
    double v7 = G_PI * (77096675.0 / 254016.0 + 378515.0 / 1512.0 * eta - 74045.0 / 756.0 * eta * eta);
    {
        const double chi1sq = chi1 * chi1, chi2sq = chi2 * chi2;
        const double SL = m1M * m1M * chi1 + m2M * m2M * chi2;
        const double dSigmaL = dm * (m2M * chi2 - m1M * chi1);
        double pn_sigma = eta * (721.0 / 48.0 * chi1 * chi2 - 247.0 / 48.0 * chi1 * chi2);
        pn_sigma += (720.0 - 1.0) / 96.0 * m1M * m1M * chi1 * chi1;
        pn_sigma += (720.0 - 1.0) / 96.0 * m2M * m2M * chi2 * chi2;
        pn_sigma -= (240.0 - 7.0) / 96.0 * m1M * m1M * chi1sq;
        pn_sigma -= (240.0 - 7.0) / 96.0 * m2M * m2M * chi2sq;
        double pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1 * chi2;
        pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m1M - 120.0 * m1M * m1M)
                   + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1sq;
        pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m2M - 120.0 * m2M * m2M)
                   + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2sq;
        const double pn_gamma = (554345.0 / 1134.0 + 110.0 * eta / 9.0) * SL
                              + (13915.0 / 84.0 - 10.0 * eta / 3.0) * dSigmaL;
        v7 += (-8980424995.0 / 762048.0 + 6586595.0 * eta / 756.0 - 305.0 * eta * eta / 36.0) * SL
            - (170978035.0 / 48384.0 - 2876425.0 * eta / 672.0 - 4735.0 * eta * eta / 144.0) * dSigmaL;
        v6 += G_PI * (3760.0 * SL + 1490.0 * dSigmaL) / 3.0 + pn_ss3;
        v5 += -1.0 * pn_gamma;
        vl5 += -3.0 * pn_gamma;
        v4 += -10.0 * pn_sigma;
        v3 += 188.0 * SL / 3.0 + 25.0 * dSigmaL;
    }
    // fold v = (pi M f)^(1/3) = m3 * x,  x = f^(1/3), into coefficients [turns]
    const double m_sec = mtot * G_MTSUN;
    const double piM = G_PI * m_sec;
    const double m3 = cbrt(piM);
    const double lm3 = log(m3);
    const double s = pfaN / G_TWOPI;
    const double im3 = 1.0 / m3;
    const double im3_2 = im3 * im3, im3_3 = im3_2 * im3, im3_5 = im3_3 * im3_2;
    out[C_A0] = s * im3_5;
    out[C_A2] = s * v2 * im3_3;
    out[C_A3] = s * v3 * im3_2;
    out[C_A4] = s * v4 * im3;
    // constant term also absorbs -2 phi_c - pi/4
    out[C_A5] = s * (v5 + vl5 * lm3) - (2.0 * phic + 0.25 * G_PI) / G_TWOPI;
    out[C_B5] = s * vl5;
    out[C_A6] = s * (v6 + vl6 * lm3) * m3;
    out[C_B6] = s * vl6 * m3;
    out[C_A7] = s * v7 * m3 * m3;
    {   // tidal terms (lalsimulation TaylorF2Phasing_10PN/12PNTidalCoeff): pfaN (pft10 v^10 + pft12 v^12) / v^5
        const double m1M4 = m1M * m1M * m1M * m1M, m2M4 = m2M * m2M * m2M * m2M;
        const double pft10 = lam1 * (-288.0 + 264.0 * m1M) * m1M4 + lam2 * (-288.0 + 264.0 * m2M) * m2M4;
        const double pft12 = lam1 * (-15895.0 / 28.0 + 4595.0 / 28.0 * m1M + 5715.0 / 14.0 * m1M * m1M
                                     - 325.0 / 7.0 * m1M * m1M * m1M) * m1M4
                           + lam2 * (-15895.0 / 28.0 + 4595.0 / 28.0 * m2M + 5715.0 / 14.0 * m2M * m2M
                                     - 325.0 / 7.0 * m2M * m2M * m2M) * m2M4;
        const double m3_2 = m3 * m3, m3_5 = m3_2 * m3_2 * m3;
        out[C_A10] = s * pft10 * m3_5;
        out[C_A12] = s * pft12 * m3_5 * m3_2;
    }
    return m3;
}
// The reference chirp of Newtonian coefficient a0: equal masses, no spins, no tides.
__host__ __device__ inline void ref_phase_coefs(double a0, double* out) {
    // a0 = 3 / (128 eta 2 pi) (pi M)^(-5/3) with eta = 1/4
    const double piM = pow(a0 * (64.0 * G_PI / 3.0), -0.6);
    const double m = 0.5 * piM / (G_PI * G_MTSUN);
    pn_phase_coefs(m, m, 0.0, 0.0, 0.0, 0.0, 0.0, out);
}
// a0 from the masses alone (one cbrt): (3 / 256 pi) (pi Mc)^(-5/3), Mc^5 = (m1 m2)^3 / (m1 + m2)
__host__ __device__ inline double newtonian_coef(double m1, double m2) {
    const double pm = m1 * m2;
    const double mc5 = pm * pm * pm / (m1 + m2);
    const double pim = G_PI * G_MTSUN;
    return 3.0 / (256.0 * G_PI) / (cbrt(mc5) * pim * cbrt(pim * pim));
}

__device__ __forceinline__ void load_phase_coef(PhaseCoef& c, const double* wc, int wpad, int w) {
    c.a0 = wc[C_A0 * wpad + w]; c.a2 = wc[C_A2 * wpad + w]; c.a3 = wc[C_A3 * wpad + w];
    c.a4 = wc[C_A4 * wpad + w]; c.a5 = wc[C_A5 * wpad + w]; c.b5 = wc[C_B5 * wpad + w];
    c.a6 = wc[C_A6 * wpad + w]; c.b6 = wc[C_B6 * wpad + w]; c.a7 = wc[C_A7 * wpad + w];
    c.a10 = wc[C_A10 * wpad + w]; c.a12 = wc[C_A12 * wpad + w];
}

// ---------------------------------------------------------------------------
// kernel 0: sort key = last bin of each walker's waveform (depends on total mass only)
// ---------------------------------------------------------------------------
// one walker's parameter row, gathered from either layout (walker-major or parameter-major)
__device__ __forceinline__ void load_params(double (&p)[GWIN_NPARAM], const double* __restrict__ params,
                                            int64_t w, int64_t sw, int64_t si) {
#pragma unroll
    for (int i = 0; i < GWIN_NPARAM; ++i) p[i] = params[w * sw + i * si];
}
__device__ __forceinline__ bool params_ok(const double* p) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < GWIN_NPARAM; ++i) ok = ok && isfinite(p[i]);
    return ok && p[0] > 0.0 && p[1] > 0.0 && p[4] > 0.0;
}
__device__ __forceinline__ double f_isco(double mtot) {
    const double piM = G_PI * (mtot * G_MTSUN);
    const double vISCO = 1.0 / sqrt(6.0);
    return vISCO * vISCO * vISCO / piM;
}
// lalsimulation refuses f_max <= fStart; pycbc turns that into NoWaveformError and
// the reference into loglr = -inf (gaussian_noise.py:293-302, 321-324)
__device__ __forceinline__ bool has_waveform(const double* p, double f_lower_wf) {
    return f_isco(p[0] + p[1]) > f_lower_wf;
}
__device__ __forceinline__ int waveform_len(double mtot, double delta_f) {
    // XLALSimInspiralTaylorF2: n = (size_t)(fISCO/deltaF + 1), fISCO = 6^-3/2/(pi M)
    const double piM = G_PI * (mtot * G_MTSUN);
    const double vISCO = 1.0 / sqrt(6.0);
    const double fISCO = vISCO * vISCO * vISCO / piM;
    const double n = fISCO / delta_f + 1.0;
    return n > 2.0e9 ? 2000000000 : (int)n;
}

// Host entry with a column map: the caller's points[w * sw + c * si] (ndim columns: the sampled
// parameters, in the caller's order) become the kernels' parameter-major rows; parameters that are
// not sampled take their fixed value.
struct PointMap { int col[GWIN_NPARAM]; double fixed[GWIN_NPARAM]; };
__global__ void expand_points_kernel(int W, const double* __restrict__ pts, int64_t sw, int64_t si, PointMap M,
                                     double* __restrict__ params) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
#pragma unroll
    for (int i = 0; i < GWIN_NPARAM; ++i)
        params[(size_t)i * W + w] = M.col[i] >= 0 ? pts[w * sw + M.col[i] * si] : M.fixed[i];
}

// Sort keys.  key: the order the walkers are processed in -- key_a0 = false: their last bin (lanes of a
// warp finish together); true (table path): their Newtonian chirp coefficient quantised over the
// covered range (lanes of a warp share a reference chirp and sit in neighbouring table rows).
// key_e (table path): the last bin again, the order in which the tails are taken.  key_t (table path):
// the chirp's local time, the order in which the long sub-blocks at high frequency are taken (lanes
// of a warp read the same table rows).  Walkers without a waveform sort to one end.
__global__ void sort_key_kernel(PlanDev P, int W, const double* __restrict__ params,
                                int64_t sw, int64_t si,
                                const uint8_t* __restrict__ skip, int* __restrict__ key,
                                int* __restrict__ idx, bool key_a0, double a_lo, double a_scale,
                                int* __restrict__ key_e, int* __restrict__ key_t, double tfac,
                                double e_lo, double e_scale) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double p[GWIN_NPARAM];
    load_params(p, params, w, sw, si);
    const bool ok = !(skip && skip[w]) && params_ok(p) && has_waveform(p, P.f_lower_wf);
    int k = key_a0 ? (1 << TAB_KEY_BITS) - 1 : 0, ke = 0, kt = (1 << 22) - 1;
    if (ok) {
        // last bin, quantised to 12 bits over the range it can take (bucket order, see bucket_order_kernel)
        ke = min(waveform_len(p[0] + p[1], P.delta_f), P.kmax);
        ke = 1 + (int)fmin(fmax(((double)ke - e_lo) * e_scale, 0.0), 4094.0);
        if (key_a0) {
            const double a0 = newtonian_coef(p[0], p[1]);
            const double q = (a0 - a_lo) * a_scale;
            k = (int)fmin(fmax(q, 0.0), (double)((1 << TAB_KEY_BITS) - 2));
            // local time of the chirp at the frequency where the time-ordered sub-blocks begin, as a
            // fraction of the segment: t_c - (5/3) a0 f^(-8/3)   (Theta = a0 f^(-5/3) turns to leading order)
            const double tl = (p[10] - P.epoch - a0 * tfac) * P.delta_f;
            kt = (int)((tl - floor(tl)) * (double)((1 << 22) - 2));
        } else {
            k = ke;
        }
    }
    key[w] = k;
    idx[w] = w;
    if (key_e) key_e[w] = ke;
    if (key_t) key_t[w] = kt;
}

// Batches of up to 16384 walkers: one CTA sorts (key, index) pairs in shared memory (stable LSD radix
// sort, cub::BlockRadixSort); CTA x sorts key set x -- the table path's three orders in one launch
// instead of three device-wide sorts of five launches each.
struct SortJob { const int* key; int bits; int* out; };
struct SortJobs { SortJob j[3]; };
template <int ITEMS>
__global__ void __launch_bounds__(1024)
block_sort_kernel(int W, SortJobs jobs) {
    using Sort = cub::BlockRadixSort<int, 1024, ITEMS, int>;
    extern __shared__ __align__(16) unsigned char s_sort[];
    typename Sort::TempStorage& temp = *reinterpret_cast<typename Sort::TempStorage*>(s_sort);
    const int* key = jobs.j[blockIdx.x].key;
    int* out = jobs.j[blockIdx.x].out;
    const int bits = jobs.j[blockIdx.x].bits;
    int k[ITEMS], v[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int w = threadIdx.x * ITEMS + i;
        k[i] = w < W ? key[w] : 0x7fffffff;
        v[i] = w;
    }
    Sort(temp).Sort(k, v, 0, bits + 1);                 // + 1: the padding keys sort last
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int w = threadIdx.x * ITEMS + i;
        if (w < W) out[w] = v[i];
    }
}

// Cheaper than a sort, and all the kernels need: walkers ordered by the top 12 bits of their key (4096
// buckets; within a bucket in whatever order the atomics fall -- no result depends on a walker's
// neighbours).  One CTA per order: shared-memory histogram, scan, scatter.
#define ORDER_BUCKETS 4096
template <int ITEMS>
__global__ void __launch_bounds__(1024)
bucket_order_kernel(int W, SortJobs jobs) {
    using Scan = cub::BlockScan<int, 1024>;
    __shared__ int s_hist[ORDER_BUCKETS];
    __shared__ typename Scan::TempStorage s_scan;
    const int* key = jobs.j[blockIdx.x].key;
    int* out = jobs.j[blockIdx.x].out;
    const int shift = max(jobs.j[blockIdx.x].bits - 12, 0);
    for (int i = threadIdx.x; i < ORDER_BUCKETS; i += 1024) s_hist[i] = 0;
    __syncthreads();
    int bkt[ITEMS], pos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int w = threadIdx.x + i * 1024;
        if (w < W) {
            bkt[i] = min(key[w] >> shift, ORDER_BUCKETS - 1);
            pos[i] = atomicAdd(&s_hist[bkt[i]], 1);
        }
    }
    __syncthreads();
    int cnt[ORDER_BUCKETS / 1024], off[ORDER_BUCKETS / 1024];
#pragma unroll
    for (int i = 0; i < ORDER_BUCKETS / 1024; ++i) cnt[i] = s_hist[threadIdx.x * (ORDER_BUCKETS / 1024) + i];
    Scan(s_scan).ExclusiveSum(cnt, off);
#pragma unroll
    for (int i = 0; i < ORDER_BUCKETS / 1024; ++i) s_hist[threadIdx.x * (ORDER_BUCKETS / 1024) + i] = off[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int w = threadIdx.x + i * 1024;
        if (w < W) out[s_hist[bkt[i]] + pos[i]] = w;
    }
}

// What walker_setup needs to split a walker's band between the kernels.
struct FastArgs {
    double a_sched_max;      // largest Newtonian coefficient (smallest chirp mass) the recurrence schedule covers
    double cov_lo, cov_hi;   // Newtonian coefficients the moment tables cover (table path)
    int table;               // table path: recurrence below k_split, tables above, tails by a second recurrence pass
    int k_split, k_tab_end, n_tblk;
    const TabBlk* blks;
    int *ke_fast, *ke_low, *ks_tail, *inv;      // inv[w]: position of walker w in the processing order
    int *ks_rec;                                // tails: start of what is left to the recurrence kernel
    int *ks_fine, *tail_parent;                 // tails: end of the stretch taken by fine sub-blocks; the sub-block the walker ends in (-1: no fine sub-blocks)
    unsigned char* slow;
};

// ---------------------------------------------------------------------------
// kernel 1: per-walker coefficients.  Thread t handles walker perm[t] and writes
// its rows at sorted position t.
// ---------------------------------------------------------------------------
__global__ void walker_setup_kernel(PlanDev P, int W, int wpad,
                                    const double* __restrict__ params, int64_t sw, int64_t si,
                                    const uint8_t* __restrict__ skip,
                                    const int* __restrict__ perm,
                                    double* __restrict__ wc, int* __restrict__ ks_arr,
                                    int* __restrict__ ke_arr,
                                    double* __restrict__ hh_sorted /*[n_det][wpad]*/,
                                    const double* __restrict__ calib, int64_t csw, int64_t csi,
                                    FastArgs F) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W) return;
    const int w = perm ? perm[t] : t;
    double p[GWIN_NPARAM];
    load_params(p, params, w, sw, si);
    const int nd = P.n_det;
    bool ok = !(skip && skip[w]) && params_ok(p) && has_waveform(p, P.f_lower_wf);
    if (calib && ok)
        for (int v = 0; v < nd * 2 * P.n_cal; ++v) ok = ok && isfinite(calib[w * csw + v * csi]);
    if (!ok) {
        // no bins; combine() turns this into -inf via ks = -1
        ks_arr[t] = -1; ke_arr[t] = -1;
        if (F.ke_fast) { F.ke_fast[t] = -1; F.slow[t] = 0; }
        if (F.table) { F.ke_low[t] = -1; F.ks_tail[t] = -1; F.ks_fine[t] = -1; F.ks_rec[t] = -1; F.tail_parent[t] = -1; F.inv[w] = t; }
        for (int r = 0; r < NCOEF(nd); ++r) wc[r * wpad + t] = 0.0;
        for (int d = 0; d < nd; ++d) hh_sorted[d * wpad + t] = 0.0;
        return;
    }
    const double m1 = p[0], m2 = p[1], chi1 = p[2], chi2 = p[3], dist = p[4];
    const double incl = p[5], phic = p[6], ra = p[7], dec = p[8], psi = p[9], tc = p[10];

    // --- XLALSimInspiralPNPhasing_F2 folded into 11 phase coefficients [turns] ---------
    const double mtot = m1 + m2;
    const double eta = m1 * m2 / mtot / mtot;
    double m3;
    {
        double pcf[NPH];
        m3 = pn_phase_coefs(m1, m2, chi1, chi2, phic, p[11], p[12], pcf);
#pragma unroll
        for (int i = 0; i < NPH; ++i) wc[(size_t)i * wpad + t] = pcf[i];
    }
    const double im3 = 1.0 / m3;
    const double im3_3 = im3 * im3 * im3;

    // --- band ------------------------------------------------------------------
    const int n = waveform_len(mtot, P.delta_f);
    const int ke = min(n, P.kmax);
    const int ks = max(P.kmin, P.istart);
    ks_arr[t] = ks;
    ke_arr[t] = ke;        // ke <= ks means "detectors contribute 0" (gaussian_noise.py:329-333)
    if (F.ke_fast) {
        // A walker the fast kernels are not built for -- lighter than the recurrence schedule's smallest
        // chirp mass, or outside the range the moment tables cover -- keeps an empty band there and is
        // evaluated by the direct kernel afterwards (slow path: always right, counted).
        const double a0 = wc[(size_t)C_A0 * wpad + t];
        const bool slow = a0 > F.a_sched_max || (F.table && (a0 < F.cov_lo || a0 > F.cov_hi));
        F.slow[t] = slow ? 1 : 0;
        const int kef = slow ? ks : ke;
        F.ke_fast[t] = kef;
        if (F.table) {
            F.ke_low[t] = min(kef, F.k_split);      // the recurrence kernel's share of the band
            // table sub-blocks are taken whole; what lies between the last whole one and the walker's
            // last bin is left to a second pass of the recurrence kernel over [ks_tail, ke)
            // Where the sub-block the walker ends in carries fine sub-blocks, they take the stretch up to
            // the last multiple of the finest length; the recurrence pass keeps [ks_fine, ke).
            int tail = kef, fine = kef, parent = -1;
            if (kef > F.k_split) {
                if (kef >= F.k_tab_end) {
                    tail = fine = F.k_tab_end;
                } else {
                    int lo = 0, hi = F.n_tblk - 1;          // the sub-block holding bin kef
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (F.blks[mid].start <= kef) lo = mid; else hi = mid - 1; }
                    tail = fine = F.blks[lo].start;
                    const int fmin = F.blks[lo].fine_min;
                    if (F.blks[lo].fine_n > 0) {
                        fine = tail + (kef - tail) / fmin * fmin;
                        parent = lo;
                    }
                }
            }
            F.ks_tail[t] = tail;
            F.ks_fine[t] = fine;
            F.ks_rec[t] = parent >= 0 ? kef : tail;    // with fine sub-blocks the bins after them are done there too
            F.tail_parent[t] = parent;
            F.inv[w] = t;
        }
    }

    // --- amplitude: amp0 * sqrt(5/(32 eta)) * (pi M)^(-7/6) -------------------------
    const double r = dist * 1e6 * G_PC_SI;
    const double amp0 = -4.0 * m1 * m2 / r * G_MRSUN * G_MTSUN * sqrt(G_PI / 12.0);
    const double ampfac = amp0 * sqrt(5.0 / (32.0 * eta)) * (im3_3 / sqrt(m3));
    const double cfac = cos(incl);
    const double pfac = 0.5 * (1.0 + cfac * cfac);

    // --- XLALGreenwichMeanSiderealTime(tc), bit-faithful to LAL's arithmetic: the
    // Julian day is formed from INTEGER UTC seconds in a double (ulp 4.7e-10 d = 40 us),
    // the nanoseconds enter separately; no FMA contraction.
    double gmst;
    {
        const double sec_f = floor(tc);
        double nsd = rint((tc - sec_f) * 1e9);
        long long sec = (long long)sec_f;
        if (nsd >= 1e9) { sec += 1; nsd -= 1e9; }
        const long long utc = sec - (long long)P.gps_utc;
        long long days = utc / 86400;
        long long sod = utc - days * 86400;
        if (sod < 0) { sod += 86400; days -= 1; }
        // XLALConvertCivilTimeToJD: jd(int) + sec/86400. - 0.5 ; GPS 0 = 1980-01-06 -> 2444245
        const double jd = __dadd_rn(__dadd_rn((double)(2444245LL + days), __ddiv_rn((double)sod, 86400.0)), -0.5);
        const double t_hi = __ddiv_rn(__dadd_rn(jd, -2451545.0), 36525.0);
        const double t_lo = __ddiv_rn(nsd, 1e9 * 36525.0 * 86400.0);
        const double tj = __dadd_rn(t_hi, t_lo);
        double sid = __dadd_rn(__dmul_rn(-6.2e-6, tj), 0.093104);
        sid = __dadd_rn(__dmul_rn(__dmul_rn(sid, tj), tj), 67310.54841);
        sid = __dadd_rn(sid, __dmul_rn(8640184.812866, t_lo));
        sid = __dadd_rn(sid, __dmul_rn(3155760000.0, t_lo));
        sid = __dadd_rn(sid, __dmul_rn(8640184.812866, t_hi));
        sid = __dadd_rn(sid, __dmul_rn(3155760000.0, t_hi));
        gmst = __ddiv_rn(__dmul_rn(sid, G_PI), 43200.0);
    }
    const double gha = gmst - ra;
    double singha, cosgha, sindec, cosdec, sinpsi, cospsi;
    sincos(gha, &singha, &cosgha);
    sincos(dec, &sindec, &cosdec);
    sincos(psi, &sinpsi, &cospsi);
    const double X[3] = { -cospsi * singha - sinpsi * cosgha * sindec,
                          -cospsi * cosgha + sinpsi * singha * sindec,
                           sinpsi * cosdec };
    const double Y[3] = {  sinpsi * singha - cospsi * cosgha * sindec,
                           sinpsi * cosgha + cospsi * singha * sindec,
                           cospsi * cosdec };
    const double eh[3] = { cosdec * cosgha, cosdec * -singha, sindec };

    for (int d = 0; d < nd; ++d) {
        double fp = 0.0, fc = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double dx = P.resp[d][3 * i] * X[0] + P.resp[d][3 * i + 1] * X[1] + P.resp[d][3 * i + 2] * X[2];
            const double dy = P.resp[d][3 * i] * Y[0] + P.resp[d][3 * i + 1] * Y[1] + P.resp[d][3 * i + 2] * Y[2];
            fp += X[i] * dx - Y[i] * dy;
            fc += X[i] * dy + Y[i] * dx;
        }
        // time_delay_from_earth_center: (0 - loc).ehat / c
        const double delay = (-P.loc[d][0] * eh[0] - P.loc[d][1] * eh[1] - P.loc[d][2] * eh[2]) / G_C_SI;
        // generator.py: tc_det = tc + delay; apply_fd_time_shift: dt = tc_det - epoch
        const double tcd = tc + delay;
        const double dt = tcd - P.epoch;
        double tau = dt * P.delta_f;                 // turns per bin
        tau = tau - rint(tau);
        const double tau_hi = rint(tau * 1073741824.0) / 1073741824.0;   // 2^30: k*tau_hi exact
        const double tau_lo = tau - tau_hi;
        // h_det = (F+ pfac - i Fx cfac) * ampfac * f^(-7/6) * e^{-i(...)}
        const double are = fp * pfac * ampfac;
        const double aim = -fc * cfac * ampfac;
        double* row = wc + (size_t)(C_DET0 + D_NROW * d) * wpad + t;
        row[D_TAUHI * wpad] = tau_hi;
        row[D_TAULO * wpad] = tau_lo;
        row[D_ARE * wpad] = are;
        row[D_AIM * wpad] = aim;
        double es, ec;
        sincospi(2.0 * tau, &es, &ec);
        row[D_ECOS * wpad] = ec;
        row[D_ESIN * wpad] = es;
        // <h|h> closed form: |A|^2 * sum_{ks<=k<ke} norm f^(-7/3)/S  (phase-free)
        double hh = 0.0;
        if (calib) {
            // least-squares cubics through the spline values (what the reference's
            // UnivariateSpline(points, values) with its default smoothing returns)
            const double* M = P.calM + (size_t)d * 4 * P.n_cal;
            double b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double ca = 0.0, cp = 0.0;
                for (int j = 0; j < P.n_cal; ++j) {
                    const double m = M[i * P.n_cal + j];
                    ca = fma(m, calib[w * csw + ((2 * d) * P.n_cal + j) * csi], ca);
                    cp = fma(m, calib[w * csw + ((2 * d + 1) * P.n_cal + j) * csi], cp);
                }
                row[(size_t)(D_CA0 + i) * wpad] = ca;
                row[(size_t)(D_CP0 + i) * wpad] = cp;
                b[i] = ca;
            }
            if (ke > ks) {
                // |1 + dA(z)|^2 is a sextic in z; |(2 + i dphi)/(2 - i dphi)| = 1
                b[0] += 1.0;
                const double q[7] = { b[0] * b[0], 2.0 * b[0] * b[1], fma(2.0 * b[0], b[2], b[1] * b[1]),
                                      2.0 * fma(b[0], b[3], b[1] * b[2]), fma(2.0 * b[1], b[3], b[2] * b[2]),
                                      2.0 * b[2] * b[3], b[3] * b[3] };
                const double* C = P.Ccal + (size_t)d * 7 * (P.n_f + 1);
                double sum = 0.0;
#pragma unroll
                for (int j = 0; j < 7; ++j) sum = fma(q[j], C[(size_t)j * (P.n_f + 1) + ke] - C[(size_t)j * (P.n_f + 1) + ks], sum);
                hh = (are * are + aim * aim) * sum;
            }
        } else if (ke > ks) {
            const double* C = P.C + (size_t)d * (P.n_f + 1);
            hh = (are * are + aim * aim) * (C[ke] - C[ks]);
        }
        hh_sorted[d * wpad + t] = hh;
    }
}

// ---------------------------------------------------------------------------
// kernel 2a: direct evaluation.  One thread = one walker; CTA = 128 walkers x
// one frequency chunk.  Per bin: FP64 phase polynomial, sincospi, and per
// detector a delay-phasor recurrence (exact re-sync every RESYNC bins), a
// rotation and a complex MAC against T.
// ---------------------------------------------------------------------------
template <int ND, bool CALIB>
__global__ void __launch_bounds__(CTA_THREADS)
loglr_direct_kernel(PlanDev P, int W, int wpad, int chunk_bins, int n_chunks,
                    const double* __restrict__ wc, const int* __restrict__ ks_arr,
                    const int* __restrict__ ke_arr, double2* __restrict__ partial,
                    const int* __restrict__ list, const int* __restrict__ n_list) {
    // list != NULL (slow path): the walkers to evaluate are columns list[0 .. *n_list) of the
    // coefficient rows; their partial sums go to slots 0 .. *n_list.  The grid's x extent is fixed
    // and strides over the list, so an empty list costs one wave of CTAs that exit at once.
    const int n = list ? *n_list : W;
    const int chunk = blockIdx.y;
    const int band0 = max(P.kmin, P.istart);
    const int k0 = band0 + chunk * chunk_bins;
    const int k1 = min(k0 + chunk_bins, P.kmax);
    for (int grp = blockIdx.x; grp * CTA_THREADS < n; grp += gridDim.x) {
    const int t = grp * CTA_THREADS + threadIdx.x;
    const bool live = t < n;
    const int col = live ? (list ? list[t] : t) : 0;
    const int ks = live ? ks_arr[col] : 0, ke = live ? ke_arr[col] : 0;
    const int lo = max(k0, ks), hi = min(k1, ke);
    const bool has = lo < hi;
    int wlo = has ? lo : 0x7fffffff, whi = has ? hi : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
        whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
    }
    double2 acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = make_double2(0.0, 0.0);

    if (wlo < whi) {
        const int tt = col;
        PhaseCoef pc;
        load_phase_coef(pc, wc, wpad, tt);
        double tau_hi[ND], tau_lo[ND];
        double2 step[ND], ph[ND];
        CalCoef cal[CALIB ? ND : 1];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double* row = wc + (size_t)(C_DET0 + D_NROW * d) * wpad + tt;
            tau_hi[d] = row[D_TAUHI * wpad];
            tau_lo[d] = row[D_TAULO * wpad];
            step[d] = phasor_neg(tau_hi[d] + tau_lo[d]);
            ph[d] = make_double2(1.0, 0.0);
            if (CALIB) load_cal_coef(cal[d], row, wpad);
        }
        const double4* __restrict__ basis = P.basis;
        const double2* __restrict__ T = P.T;
        for (int k = wlo; k < whi; ++k) {
            if (k == wlo || (k & (RESYNC - 1)) == 0) {
                const double kd = (double)k;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    ph[d] = phasor_neg(fma(kd, tau_lo[d], frac1(kd * tau_hi[d])));
            }
            const double4 b = ld_basis(basis + k);
            const double2 u = phasor_neg(phase_turns(pc, b));
            const bool on = (k >= lo) && (k < hi);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double2 Tk = __ldg(T + (size_t)k * ND + d);
                double2 wv = cmul(u, ph[d]);
                if (CALIB) wv = cmul(wv, calib_factor(cal[d], fma((double)k, P.cal_zs[d], P.cal_z0[d])));
                if (on) acc[d] = cfma(Tk, wv, acc[d]);
                ph[d] = cmul(ph[d], step[d]);
            }
        }
    }
    if (live && has) {          // chunks a walker does not reach are neither written nor read back
#pragma unroll
        for (int d = 0; d < ND; ++d)
            partial[((size_t)chunk * ND + d) * wpad + t] = acc[d];
    }
    }
}

// ---------------------------------------------------------------------------
// kernel 2b: block-anchored chirp-phasor recurrence.
//
// The band is cut (once per plan) into blocks of B_b <= MAX_BLOCK bins, short at
// low frequency where the chirp phase curves fastest.  Inside a block the
// detector phase Theta_d(j) = Psi(k0+j) + (k0+j) tau_d is replaced by the quartic
// through 5 exact samples (both block ends are nodes), whose error is bounded by
// the plan's phase_tol for every chirp mass >= mchirp_min.  The quartic's
// forward differences D1..D4 become unit phasors g, r, s, t with
//      acc <- acc * g + T[j] ;  g <- g * r ;  r <- r * s ;  s <- s * t
// (Horner form of sum_j T_j e^{2 pi i (Theta(B-1) - Theta(j))}), so that one bin
// costs 8 + 8*ND FP64 issue slots instead of a phase polynomial and a sincos.
// The block sum is then rotated by the exact anchor phasor e^{-2 pi i Theta_d(B-1)}.
// ---------------------------------------------------------------------------
// Per-thread walker coefficients live in shared memory (one column per thread, no
// sharing, hence no barrier): they are only needed in the per-block set-up, and keeping
// them out of the register file buys a third/fourth resident CTA per SM.
#define RC_ROWS(nd) (NPH + 4 * (nd))
#define RC_ROWS_CAL(nd) (RC_ROWS(nd) + 8 * (nd))       // + calibration cubics: a0..a3, p0..p3 per detector
template <int ND>
__device__ __forceinline__ void load_block_cal(const double* sc, int d, CalCoef& c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c.a[i] = sc[(RC_ROWS(ND) + 8 * d + i) * CTA_THREADS];
        c.p[i] = sc[(RC_ROWS(ND) + 8 * d + 4 + i) * CTA_THREADS];
    }
}
template <int ND>
__device__ __forceinline__ double2 load_E(const double* sc, int d) {       // e^{+2 pi i tau_d}
    return make_double2(sc[(NPH + 2 * ND + 2 * d) * CTA_THREADS], sc[(NPH + 1 + 2 * ND + 2 * d) * CTA_THREADS]);
}
template <int ND>
__device__ __forceinline__ void load_block_coef(const double* sc, PhaseCoef& pc,
                                                double (&tau_hi)[ND], double (&tau_lo)[ND]) {
    asm volatile("" ::: "memory");      // re-read per block: do not hoist into registers
    pc.a0 = sc[0 * CTA_THREADS]; pc.a2 = sc[1 * CTA_THREADS]; pc.a3 = sc[2 * CTA_THREADS];
    pc.a4 = sc[3 * CTA_THREADS]; pc.a5 = sc[4 * CTA_THREADS]; pc.b5 = sc[5 * CTA_THREADS];
    pc.a6 = sc[6 * CTA_THREADS]; pc.b6 = sc[7 * CTA_THREADS]; pc.a7 = sc[8 * CTA_THREADS];
    pc.a10 = sc[9 * CTA_THREADS]; pc.a12 = sc[10 * CTA_THREADS];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        tau_hi[d] = sc[(NPH + 2 * d) * CTA_THREADS];
        tau_lo[d] = sc[(NPH + 1 + 2 * d) * CTA_THREADS];
    }
}

// ---------------------------------------------------------------------------
// Per-warp table stream: the rows T[k][0..ND) the warp is about to consume are staged in
// shared memory by the TMA engine (cp.async.bulk global->shared, completion on an mbarrier),
// one 64-bin tile ahead of the arithmetic, double buffered.  The hot loop then reads
// warp-uniform LDS.128 (broadcast, fixed ~30-cycle latency) instead of LDG, so no L1/L2
// latency is ever exposed and no registers are spent on load pipelining.
// (ncu r1c: with LDG the compiler, at 96 registers, placed one load 8 instructions ahead of
// its use -- 31 % of hot-loop stall samples were long_scoreboard.)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}

template <int ND>
struct TileStream {
    static constexpr int TILE = ND <= 3 ? 128 : 64;    // bins per tile
    double2* buf;          // this warp's [2][TILE * ND] staging buffers
    uint32_t bar;          // shared-memory address of this warp's 2 mbarriers
    const double2* T;      // global table
    int ka, kz;            // stream covers bins [ka, kz)
    int cur;               // tile currently readable (-1: none yet)
    int lane;

    __device__ __forceinline__ void issue(int t) const {           // lane 0 only
        const int k0 = ka + t * TILE;
        if (k0 >= kz) return;
        const uint32_t bytes = (uint32_t)(min(TILE, kz - k0) * ND * (int)sizeof(double2));
        const uint32_t dst = smem_u32(buf + (size_t)(t & 1) * TILE * ND);
        const uint32_t b = bar + 8u * (uint32_t)(t & 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(T + (size_t)k0 * ND), "r"(bytes), "r"(b) : "memory");
    }
    __device__ __forceinline__ void open(double2* buf_, uint64_t* bars, const double2* T_, int ka_, int kz_, int lane_) {
        buf = buf_; bar = smem_u32(bars); T = T_; ka = ka_; kz = kz_; cur = -1; lane = lane_;
        __syncwarp();                              // a previous stream of this warp is fully closed
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar + 8u) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            issue(0);
        }
        __syncwarp();
    }
    // make the tile holding bin k readable; bins are consumed in increasing order
    __device__ __forceinline__ void ensure(int k) {
        const int t = (k - ka) / TILE;
        if (t == cur) return;
        __syncwarp();                              // every lane is done with tile t-1 ...
        if (lane == 0) issue(t + 1);               // ... so its buffer can take tile t+1
        mbar_wait(bar + 8u * (uint32_t)(t & 1), (uint32_t)((t >> 1) & 1));
        cur = t;
    }
    __device__ __forceinline__ const double2* row(int k) const {
        const int r = k - ka;
        return buf + (size_t)((r / TILE) & 1) * TILE * ND + (size_t)(r % TILE) * ND;
    }
    __device__ __forceinline__ int tile_end(int k) const {         // first bin of the next tile
        return ka + ((k - ka) / TILE + 1) * TILE;
    }
    // no bulk copy may be in flight when the CTA retires; the barriers are invalidated so that the
    // warp may open another stream on them (mbarrier.init on a live barrier is undefined)
    __device__ __forceinline__ void close() {
        const int t = cur + 1;
        if (ka + t * TILE < kz) mbar_wait(bar + 8u * (uint32_t)(t & 1), (uint32_t)((t >> 1) & 1));
        __syncwarp();
        if (lane == 0) {
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar + 8u) : "memory");
        }
        __syncwarp();
    }
};

template <int ND, bool CALIB>
__device__ __forceinline__ void direct_block(const PlanDev& P, const double* sc, const double4* __restrict__ basis,
                                             TileStream<ND>& S, int kb, int B, int lo, int hi,
                                             double2 (&acc_out)[ND]) {
    // too curved / too short for a polynomial block: evaluate these bins directly
    PhaseCoef pc;
    double tau_hi[ND], tau_lo[ND];
    load_block_coef<ND>(sc, pc, tau_hi, tau_lo);
    for (int k = kb; k < kb + B; ++k) {
        S.ensure(k);
        if (k < lo || k >= hi) continue;
        const double2* sp = S.row(k);
        const double psi = phase_turns(pc, ld_basis(basis + k));
        const double kd = (double)k;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double2 u = phasor_neg(psi + fma(kd, tau_lo[d], frac1(kd * tau_hi[d])));
            if (CALIB) {
                CalCoef cc;
                load_block_cal<ND>(sc, d, cc);
                u = cmul(u, calib_factor(cc, fma(kd, P.cal_zs[d], P.cal_z0[d])));
            }
            acc_out[d] = cfma(sp[d], u, acc_out[d]);
        }
    }
}

// Calibrated polynomial block.  The factor c_d(k) = exp(L_d(k)), L_d = log(1 + dA) + 2 i atan(dphi/2),
// is smooth on the scale of a block, so L_d is sampled at the block's nodes and joins the block
// polynomial: its imaginary part as a per-detector phase, its real part as a per-detector log
// amplitude.  g, r, s, t become per-detector and non-unit; one bin costs 16 ND (quartic) or
// 12 ND (cubic) FP64 issue slots.  Returns false -- having done nothing -- when some lane's
// amplitude factor 1 + dA is not safely positive at a node (log undefined / not smooth): the
// caller then evaluates the block bin by bin, which has no such restriction.
template <int ND, bool MASKED, int DEG>
__device__ __forceinline__ bool recur_block_cal(
        const PlanDev& P, const double* sc, const double4* __restrict__ basis, TileStream<ND>& S,
        const BlockMat* __restrict__ M, int kb, int B, int lo, int hi, double2 (&acc_out)[ND]) {
    PhaseCoef pc;
    double tau_hi[ND], tau_lo[ND];
    load_block_coef<ND>(sc, pc, tau_hi, tau_lo);
    const double f0 = phase_turns(pc, ld_basis(basis + kb));
    double fn[DEG];
#pragma unroll
    for (int i = 0; i < DEG; ++i) fn[i] = phase_turns(pc, ld_basis(basis + kb + M->node[i]));
    const double fend = fn[DEG - 1];
    double D[4];
#pragma unroll
    for (int r = 0; r < DEG; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < DEG; ++i) acc = fma(M->m[4 * r + i], fn[i] - f0, acc);
        D[r] = acc;
    }
    // per-detector differences of the log-calibration
    double Dth[ND][4], Dam[ND][4], Lend_re[ND], Lend_im[ND];
    bool bad = false;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        CalCoef cc;
        load_block_cal<ND>(sc, d, cc);
        double amp;
        const double2 L0 = calib_log(cc, fma((double)kb, P.cal_zs[d], P.cal_z0[d]), amp);
        bad = bad || !(amp > 0.25);
        double2 Ln[DEG];
#pragma unroll
        for (int i = 0; i < DEG; ++i) {
            Ln[i] = calib_log(cc, fma((double)(kb + M->node[i]), P.cal_zs[d], P.cal_z0[d]), amp);
            bad = bad || !(amp > 0.25);
        }
#pragma unroll
        for (int r = 0; r < DEG; ++r) {
            double are = 0.0, aim = 0.0;
#pragma unroll
            for (int i = 0; i < DEG; ++i) {
                are = fma(M->m[4 * r + i], Ln[i].x - L0.x, are);
                aim = fma(M->m[4 * r + i], Ln[i].y - L0.y, aim);
            }
            Dam[d][r] = are;
            Dth[d][r] = fma(aim, -1.0 / G_TWOPI, D[r]);      // theta' = theta - Im L / 2 pi  [turns]
        }
        Lend_re[d] = Ln[DEG - 1].x;
        Lend_im[d] = Ln[DEG - 1].y;
    }
    if (__any_sync(0xffffffffu, bad)) return false;

    double2 g[ND], rq[ND], sq[ND], tq[ND], acc[ND];
    const double kend = (double)(kb + B - 1);
    S.ensure(kb);
    {
        const double2* sp = S.row(kb);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            // Horner factor of bin j: exp(2 pi i dTheta'_j - dRe L_j)
            double2 G = cmul(phasor_pos(Dth[d][0]), load_E<ND>(sc, d));
            const double e1 = exp(-Dam[d][0]);
            g[d] = make_double2(G.x * e1, G.y * e1);
            double2 q = phasor_pos_small(Dth[d][1]);
            const double e2 = exp(-Dam[d][1]);
            rq[d] = make_double2(q.x * e2, q.y * e2);
            q = phasor_pos_small(Dth[d][2]);
            const double e3 = exp(-Dam[d][2]);
            sq[d] = make_double2(q.x * e3, q.y * e3);
            if (DEG == 4) {
                q = phasor_pos_small(Dth[d][3]);
                const double e4 = exp(-Dam[d][3]);
                tq[d] = make_double2(q.x * e4, q.y * e4);
            }
            double2 T0 = sp[d];
            if (MASKED && !(kb >= lo && kb < hi)) T0 = make_double2(0.0, 0.0);
            acc[d] = T0;
        }
    }
    int k = kb + 1;
    const int kn = kb + B;
    while (k < kn) {
        S.ensure(k);
        const int ke = min(kn, S.tile_end(k));
        const double2* sp = S.row(k);
#pragma unroll kRecurUnroll
        for (; k < ke; ++k, sp += ND) {
            const bool on = !MASKED || ((k >= lo) && (k < hi));
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                double2 Tk = sp[d];
                if (MASKED && !on) Tk = make_double2(0.0, 0.0);
                acc[d] = cfma(acc[d], g[d], Tk);
                g[d] = cmul(g[d], rq[d]);
                rq[d] = cmul(rq[d], sq[d]);
                if (DEG == 4) sq[d] = cmul(sq[d], tq[d]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const double anchor = fend + fma(kend, tau_lo[d], frac1(kend * tau_hi[d])) - Lend_im[d] * (1.0 / G_TWOPI);
        const double2 a = phasor_neg(anchor);
        const double e = exp(Lend_re[d]);
        acc_out[d] = cfma(acc[d], make_double2(a.x * e, a.y * e), acc_out[d]);
    }
    return true;
}

template <int ND, bool MASKED, int DEG>
__device__ __forceinline__ void recur_block(
        const double* sc, const double4* __restrict__ basis, TileStream<ND>& S,
        const BlockMat* __restrict__ M, int kb, int B, int lo, int hi, double2 (&acc_out)[ND]) {
    PhaseCoef pc;
    double tau_hi[ND], tau_lo[ND];
    load_block_coef<ND>(sc, pc, tau_hi, tau_lo);
    // exact phase at the DEG+1 nodes (both block ends are nodes)
    const double f0 = phase_turns(pc, ld_basis(basis + kb));
    double fn[DEG];
#pragma unroll
    for (int i = 0; i < DEG; ++i) fn[i] = phase_turns(pc, ld_basis(basis + kb + M->node[i]));
    const double fend = fn[DEG - 1];
    // forward differences of the interpolating polynomial: D1(0), D2(1), D3(2), D4
    double D[4];
#pragma unroll
    for (int r = 0; r < DEG; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < DEG; ++i) acc = fma(M->m[4 * r + i], fn[i] - f0, acc);
        D[r] = acc;
    }
    double2 tq = make_double2(1.0, 0.0);
    if (DEG == 4) tq = phasor_pos_small(D[3]);
    double2 sq = phasor_pos_small(D[2]);
    double2 rq = phasor_pos_small(D[1]);
    const double2 G = phasor_pos(D[0]);
    double2 g[ND], acc[ND];
    const double kend = (double)(kb + B - 1);
    S.ensure(kb);
    {
        const double2* sp = S.row(kb);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            g[d] = cmul(G, load_E<ND>(sc, d));          // e^{2 pi i (D1(0) + tau_d)}
            double2 T0 = sp[d];
            if (MASKED && !(kb >= lo && kb < hi)) T0 = make_double2(0.0, 0.0);
            acc[d] = T0;
        }
    }
    // Horner over the remaining bins, one shared-memory tile segment at a time
    int k = kb + 1;
    const int kn = kb + B;
    while (k < kn) {
        S.ensure(k);
        const int ke = min(kn, S.tile_end(k));
        const double2* sp = S.row(k);
#pragma unroll kRecurUnroll
        for (; k < ke; ++k, sp += ND) {
            const bool on = !MASKED || ((k >= lo) && (k < hi));
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                double2 Tk = sp[d];
                if (MASKED && !on) Tk = make_double2(0.0, 0.0);
                horner_step(acc[d], g[d], Tk, rq);
            }
            phasor_step(rq, sq);
            if (DEG == 4) phasor_step(sq, tq);
        }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const double anchor = fend + fma(kend, tau_lo[d], frac1(kend * tau_hi[d]));
        acc_out[d] = cfma(acc[d], phasor_neg(anchor), acc_out[d]);
    }
}

#ifndef RECUR_MIN_CTAS
#define RECUR_MIN_CTAS 3
#endif
template <int ND, bool CALIB>
__global__ void __launch_bounds__(CTA_THREADS, CALIB ? 2 : RECUR_MIN_CTAS)
loglr_recur_kernel(PlanDev P, int W, int wpad, int blocks_per_chunk, int n_blk,
                   const int* __restrict__ blk_start, const int* __restrict__ blk_mat,
                   const BlockMat* __restrict__ mats,
                   const double* __restrict__ wc, const int* __restrict__ ks_arr,
                   const int* __restrict__ ke_arr, double2* __restrict__ partial,
                   const int* __restrict__ col_arr, const int* __restrict__ grp_b) {
    // Thread t works on column col_arr[t] of the per-walker arrays (NULL: t itself).
    // grp_b == NULL: the CTA takes chunk blockIdx.y of the schedule.  grp_b != NULL (the tails of the
    // table path, walkers ordered by their last bin): the 128 walkers of this CTA only need blocks
    // [grp_b[x], grp_b[gridDim.x + x]); the CTA takes chunks y, y + gridDim.y, ... of that range and
    // writes one partial sum per walker into slot y.
    const int t = blockIdx.x * CTA_THREADS + threadIdx.x;
    const bool live = t < W;
    const int col = live ? (col_arr ? col_arr[t] : t) : 0;
    const int ks = live ? ks_arr[col] : 0, ke = live ? ke_arr[col] : 0;
    const int cb0 = grp_b ? grp_b[blockIdx.x] : 0;
    const int cb1 = grp_b ? grp_b[gridDim.x + blockIdx.x] : n_blk;
    double2 acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = make_double2(0.0, 0.0);

    // dynamic shared memory: [per-warp tile buffers][per-thread coefficient columns][mbarriers]
    constexpr int TILE = TileStream<ND>::TILE;
    constexpr int NWARP = CTA_THREADS / WARP;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    double2 (*s_tile)[2 * TILE * ND] = reinterpret_cast<double2 (*)[2 * TILE * ND]>(s_dyn);
    double* s_coef = reinterpret_cast<double*>(s_dyn + sizeof(double2) * NWARP * 2 * TILE * ND);
    uint64_t (*s_bar)[2] = reinterpret_cast<uint64_t (*)[2]>(s_coef + (CALIB ? RC_ROWS_CAL(ND) : RC_ROWS(ND)) * CTA_THREADS);
    bool coef_staged = false;
    // tails: chunks are cut on the absolute block grid (chunk c = blocks [c per, (c+1) per), slot c mod
    // gridDim.y), so a walker's tail is summed in the same order whoever its neighbours are
    int chunk = blockIdx.y;
    if (grp_b) {
        const int c0 = cb0 / blocks_per_chunk;
        chunk = c0 + ((int)blockIdx.y - c0 % (int)gridDim.y + (int)gridDim.y) % (int)gridDim.y;
    }
    for (;; chunk += gridDim.y) {
    const int b0 = chunk * blocks_per_chunk;
    if (b0 >= cb1) break;
    const int b1 = min(b0 + blocks_per_chunk, n_blk);
    const int k0 = blk_start[b0], k1 = blk_start[b1];
    const int lo = max(k0, ks), hi = min(k1, ke);
    const bool has = lo < hi;
    int wlo = has ? lo : 0x7fffffff, whi = has ? hi : 0;      // union over the warp
    int ilo = has ? lo : 0x7fffffff, ihi = has ? hi : 0;      // intersection over the warp
    const unsigned any_idle = __ballot_sync(0xffffffffu, !has);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
        whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
        ilo = max(ilo, __shfl_xor_sync(0xffffffffu, ilo, o));
        ihi = min(ihi, __shfl_xor_sync(0xffffffffu, ihi, o));
    }
    if (any_idle) { ilo = 0x7fffffff; ihi = 0; }
    if (wlo < whi) {
        double* sc = s_coef + threadIdx.x;
        if (!coef_staged) {
            coef_staged = true;
#pragma unroll
            for (int r = 0; r < NPH; ++r) sc[r * CTA_THREADS] = wc[(size_t)r * wpad + col];
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double* row = wc + (size_t)(C_DET0 + D_NROW * d) * wpad + col;
                sc[(NPH + 2 * d) * CTA_THREADS] = row[D_TAUHI * wpad];
                sc[(NPH + 1 + 2 * d) * CTA_THREADS] = row[D_TAULO * wpad];
                sc[(NPH + 2 * ND + 2 * d) * CTA_THREADS] = row[D_ECOS * wpad];
                sc[(NPH + 1 + 2 * ND + 2 * d) * CTA_THREADS] = row[D_ESIN * wpad];
                if (CALIB) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        sc[(RC_ROWS(ND) + 8 * d + i) * CTA_THREADS] = row[(size_t)(D_CA0 + i) * wpad];
                }
            }
        }
        const int lane = threadIdx.x & (WARP - 1), warp = threadIdx.x / WARP;
        TileStream<ND> S;
        bool opened = false;
        for (int b = b0; b < b1; ++b) {
            const int kb = blk_start[b], kn = blk_start[b + 1];
            if (kn <= wlo || kb >= whi) continue;       // nobody in this warp needs it
            if (!opened) {                              // stream starts at the first needed block
                S.open(s_tile[warp], s_bar[warp], P.T, kb, k1, lane);
                opened = true;
            }
#ifndef NO_L2_PREFETCH
            // Pull the NEXT block's table rows (and its basis nodes) into L2 now, one 128-byte
            // line per lane: every line of the table is first touched by a whole wave-front of
            // CTAs at once, and an un-prefetched first touch parks all of them on DRAM.
            if (b + 1 < b1 && kn < whi) {
                const int kn2 = blk_start[b + 2];
                const char* base = reinterpret_cast<const char*>(P.T + (size_t)kn * ND);
                const int bytes = (kn2 - kn) * ND * (int)sizeof(double2);
                for (int off = lane * 128; off < bytes; off += WARP * 128)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(base + off));
                if (lane < 5) {
                    const BlockMat* M2 = mats + blk_mat[b + 1];
                    const int node = lane == 0 ? 0 : M2->node[lane - 1];
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(P.basis + kn + node));
                }
            }
#endif
            const BlockMat* M = mats + blk_mat[b];
            const int deg = M->deg;
            const bool full = kb >= ilo && kn <= ihi;       // every lane covers the whole block
            if (CALIB) {
                bool done = false;
                if (deg == 4) done = recur_block_cal<ND, true, 4>(P, sc, P.basis, S, M, kb, kn - kb, lo, hi, acc);
                else if (deg == 3) done = recur_block_cal<ND, true, 3>(P, sc, P.basis, S, M, kb, kn - kb, lo, hi, acc);
                if (!done) direct_block<ND, true>(P, sc, P.basis, S, kb, kn - kb, lo, hi, acc);
            } else if (deg == 4) {
                if (full) recur_block<ND, false, 4>(sc, P.basis, S, M, kb, kn - kb, lo, hi, acc);
                else      recur_block<ND, true, 4>(sc, P.basis, S, M, kb, kn - kb, lo, hi, acc);
            } else if (deg == 3) {
                if (full) recur_block<ND, false, 3>(sc, P.basis, S, M, kb, kn - kb, lo, hi, acc);
                else      recur_block<ND, true, 3>(sc, P.basis, S, M, kb, kn - kb, lo, hi, acc);
            } else {
                direct_block<ND, false>(P, sc, P.basis, S, kb, kn - kb, lo, hi, acc);
            }
        }
        if (opened) S.close();
    }
    if (!grp_b) {
        if (live && has) {      // chunks a walker does not reach are neither written nor read back
#pragma unroll
            for (int d = 0; d < ND; ++d)
                partial[((size_t)chunk * ND + d) * wpad + col] = acc[d];
        }
        break;
    }
    }
    if (grp_b && live) {
#pragma unroll
        for (int d = 0; d < ND; ++d)
            partial[((size_t)blockIdx.y * ND + d) * wpad + col] = acc[d];
    }
}


// ---------------------------------------------------------------------------
// kernel 2c: table-driven <d|h> above k_split.
//
// A sub-block of m bins centred on kc = s + m/2, s_j = j / h, h = m/2.  The TaylorF2 phase is a
// combination of 11 fixed functions of f (powers of f^(1/3) and two logarithms), so its Taylor
// coefficients about kc are a per-sub-block 11 x TAB_NT matrix (host, long double) applied to the
// walker's phase coefficients:  Theta_d(kc + j) = sum_n t_n s_j^n  (+ the detector's linear term).
// Per sub-block there are K reference chirps (equal masses, no spin, a0 on a uniform grid over the
// covered range); the tables of reference r hold T e^{-2 pi i N_r(j)}, N_r the reference's exact phase
// minus its own constant and linear Taylor terms.  Relative to its nearest reference a walker has
//     Theta_d(kc + j) - N_r(j) = th0 + nu_d j + R(j),  R = sum_{n=2..8} (t_n - t_n^ref) s^n,
// |2 pi R| <= a few 1e-2, and
//     sum_j T[kc+j] e^{-2 pi i Theta_d} = e^{-2 pi i th0} sum_{n <= TAB_DEG} e_n M_n(nu_d),
//     M_n(nu) = sum_j s_j^n T[kc+j] e^{-2 pi i N_r(j)} e^{-2 pi i nu j},
// e_n the Taylor coefficients of e^{-2 pi i R} (recurrence n e_n = sum_k k u_k e_{n-k}).  M_n -- a
// trigonometric polynomial of nu with period 1 -- is tabulated on the grid nu = g/(2m) (tab_dft_kernel:
// the DFT of the kernel-deconvolved sequence) and evaluated per walker by TAB_W-tap interpolation
// with the "exponential of semicircle" kernel (type-2 NUFFT, Barnett et al. 2019).  Cost per
// (walker, sub-block, detector): O(TAB_W TAB_DEG), whatever m.  The dropped terms e_9, e_10 are
// formed as well: a walker whose truncation estimate is too large (it lies outside what the
// schedule was built for) is handed to the slow path instead of getting a wrong number.
// Walkers are ordered by a0, so the lanes of a warp share the reference and sit within a few grid
// points of each other: the warp stages the union of their table rows in shared memory.
// ---------------------------------------------------------------------------
__global__ void tab_tref_kernel(int n_pairs, const int* __restrict__ pair_blk, const TabBlk* __restrict__ blks,
                                const double* __restrict__ cmat, double* __restrict__ tref,
                                double* __restrict__ refcoef) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    const int b = pair_blk[pair];
    const TabBlk B = blks[b];
    const int r = pair - B.pair0;
    double c[NPH];
    ref_phase_coefs(B.a_lo + ((double)r + 0.5) * B.da, c);
    const double* cm = cmat + (size_t)b * NPH * TAB_NTP;
    for (int n = 0; n < TAB_NTP; ++n) {
        double acc = 0.0;
        for (int i = 0; i < NPH; ++i) acc = fma(c[i], cm[i * TAB_NTP + n], acc);
        tref[(size_t)pair * TAB_NTP + n] = acc;
    }
    for (int i = 0; i < NPH; ++i) refcoef[(size_t)pair * NPH + i] = c[i];
}

// seg[j] = T[s + j] e^{-2 pi i N_r(j)} / phi_hat(j): the sequence whose moments' DFT is the table
__global__ void __launch_bounds__(256)
tab_seg_kernel(PlanDev P, const int* __restrict__ pair_blk, const TabBlk* __restrict__ blks,
               const long long* __restrict__ pair_seg, const int* __restrict__ blk_hat,
               const double* __restrict__ invhat, const double* __restrict__ tref,
               const double* __restrict__ refcoef, double2* __restrict__ seg, int pair_base) {
    const int pair = pair_base + blockIdx.y, d = blockIdx.z;
    const int b = pair_blk[pair];
    const int s = blks[b].start, m = blks[b].m;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    PhaseCoef pc;
    load_phase_coef(pc, refcoef + (size_t)pair * NPH, 1, 0);
    const int k = s + j, jc = j - m / 2;
    const double th = phase_turns(pc, ld_basis(P.basis + k));
    const double nref = (th - tref[(size_t)pair * TAB_NTP]) - tref[(size_t)pair * TAB_NTP + 1] * ((double)jc / (0.5 * (double)m));
    const double2 Tk = __ldg(P.T + (size_t)k * P.n_det + d);
    double2 u = cmul(Tk, phasor_neg(nref));
    const double w = invhat[blk_hat[b] + j];
    u.x *= w; u.y *= w;
    seg[pair_seg[pair] + (size_t)d * m + j] = u;
}

// tab[g][n] = sum_j seg[j] s_j^n e^{-2 pi i j_c g / L}, j_c = j - m/2, L = 2m; rows carry TAB_W
// wrap-around points so that taps never wrap.  One thread per grid point; the sequence goes through
// shared memory in tiles; the twiddle factor advances by rotation, re-anchored every 32 terms.
__global__ void __launch_bounds__(256)
tab_dft_kernel(const int* __restrict__ pair_blk, const TabBlk* __restrict__ blks,
               const long long* __restrict__ pair_seg, const double2* __restrict__ seg,
               double2* __restrict__ tab, int n_det, int pair_base) {
    const int pair = pair_base + blockIdx.y, d = blockIdx.z;
    const int b = pair_blk[pair];
    const TabBlk B = blks[b];
    const int m = B.m, L = 2 * m;
    if ((int)(blockIdx.x * blockDim.x) >= L + TAB_W) return;          // uniform per CTA
    const int gp = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = gp < L + TAB_W;
    const int g = gp >= L ? gp - L : gp;
    const double2* sq = seg + pair_seg[pair] + (size_t)d * m;
    const double ih2 = 2.0 / (double)m, invL = 1.0 / (double)L;
    __shared__ double2 s_seg[256];
    double2 acc[TAB_DEG + 1];
#pragma unroll
    for (int n = 0; n <= TAB_DEG; ++n) acc[n] = make_double2(0.0, 0.0);
    double sn, cs;
    sincospi(-2.0 * (double)g * invL, &sn, &cs);
    const double2 rot = make_double2(cs, sn);                   // e^{-2 pi i g / L}
    for (int j0 = 0; j0 < m; j0 += 256) {
        __syncthreads();
        s_seg[threadIdx.x] = j0 + (int)threadIdx.x < m ? sq[j0 + threadIdx.x] : make_double2(0.0, 0.0);
        __syncthreads();
        if (!live) continue;
        const int nj = min(256, m - j0);
        double2 w = make_double2(1.0, 0.0);
        for (int jj = 0; jj < nj; ++jj) {
            const int jc = j0 + jj - m / 2;
            if ((jj & 31) == 0) {
                int idx = (int)(((long long)jc * g) % L);      // (jc g) mod L
                if (idx < 0) idx += L;
                sincospi(-2.0 * (double)idx * invL, &sn, &cs);
                w = make_double2(cs, sn);
            }
            double2 u = cmul(s_seg[jj], w);
            w = cmul(w, rot);
            const double sc = (double)jc * ih2;
#pragma unroll
            for (int n = 0; n <= TAB_DEG; ++n) {
                acc[n].x += u.x; acc[n].y += u.y;
                u.x *= sc; u.y *= sc;
            }
        }
    }
    if (!live) return;
    // layout [r][d][g][TAB_NM]: the moments of one grid point are contiguous (one cache line per tap)
    double2* out = tab + B.off + (((size_t)(pair - B.pair0) * n_det + d) * (L + TAB_W) + gp) * TAB_NM;
    out[0] = acc[0];
#pragma unroll
    for (int n = 2; n <= TAB_DEG; ++n) out[n - 1] = acc[n];
}

// What a batch needs the tables to cover: range of the Newtonian coefficient a0 and, per other phase
// coefficient, the largest deviation from the reference chirp of the same a0.
// out[0] = min a0, [1] = max a0, [2 + i] = max |a_i - a_i^ref| (bit patterns of non-negative doubles
// order like the numbers), [2 + NPH] = number of walkers with a waveform, [14] / [15] = min / max last bin.
__global__ void cover_stats_kernel(PlanDev P, int W, const double* __restrict__ params, int64_t sw, int64_t si,
                                   const uint8_t* __restrict__ skip, unsigned long long* __restrict__ out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2 + NPH];
    v[0] = INFINITY; v[1] = 0.0;
#pragma unroll
    for (int i = 0; i < NPH; ++i) v[2 + i] = 0.0;
    bool ok = false;
    double ke_lo = INFINITY, ke_hi = 0.0;
    if (w < W) {
        double p[GWIN_NPARAM];
        load_params(p, params, w, sw, si);
        ok = !(skip && skip[w]) && params_ok(p) && has_waveform(p, P.f_lower_wf);
        if (ok) {
            ke_lo = ke_hi = (double)min(waveform_len(p[0] + p[1], P.delta_f), P.kmax);
            double a[NPH], ar[NPH];
            pn_phase_coefs(p[0], p[1], p[2], p[3], 0.0, p[11], p[12], a);
            ref_phase_coefs(a[0], ar);
            v[0] = v[1] = a[0];
#pragma unroll
            for (int i = 1; i < NPH; ++i) v[2 + i] = fabs(a[i] - ar[i]);
            v[2 + C_A5] = 0.0;                  // the constant term carries no curvature
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        v[0] = fmin(v[0], __shfl_xor_sync(0xffffffffu, v[0], o));
        ke_lo = fmin(ke_lo, __shfl_xor_sync(0xffffffffu, ke_lo, o));
        ke_hi = fmax(ke_hi, __shfl_xor_sync(0xffffffffu, ke_hi, o));
#pragma unroll
        for (int i = 1; i < 2 + NPH; ++i) v[i] = fmax(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
    }
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, ok));
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicMin(out, (unsigned long long)__double_as_longlong(v[0]));
        for (int i = 1; i < 2 + NPH; ++i) atomicMax(out + i, (unsigned long long)__double_as_longlong(v[i]));
        atomicAdd(out + 2 + NPH, (unsigned long long)cnt);
        atomicMin(out + 14, (unsigned long long)__double_as_longlong(ke_lo));
        atomicMax(out + 15, (unsigned long long)__double_as_longlong(ke_hi));
    }
}

#ifndef TABLE_MIN_CTAS
#define TABLE_MIN_CTAS 2     // 255 registers: a third resident CTA (168) spills in the tap loop and loses 60 %
#endif
#ifndef TAB_SPAN
#define TAB_SPAN 48
#endif
#define TAB_KPRE 6           // steps (of 4 window rows) whose B fragments are fetched ahead of the weights
// TAB_SPAN: widest window of table rows (TAB_W + the spread of the lanes' first rows) the tensor-core product takes
#define TAB_WDP 36           // pitch [doubles] of the dense weight matrix [window row][lane]: conflict-free A fragments
#ifndef TAB_CHUNK_MAX
#define TAB_CHUNK_MAX 8
#endif
// TAB_CHUNK_MAX: sub-blocks per CTA at most (their Taylor matrices are staged in shared memory)
#define TAB_CROWS(nd) (NPH + 2 * (nd))      // per-thread coefficient column: phase coefficients, tau_hi/lo per detector
// per lane (i.e. x 32 per warp) [doubles]: its column of the dense weight matrix [TAB_SPAN][TAB_WDP] and its row of
// the moments handed back by the tensor cores [32][2 TAB_NM]
#define TAB_WPITCH ((TAB_SPAN * TAB_WDP + WARP - 1) / WARP + 2 * TAB_NM)
// Thread t works on column col_arr[t] of the per-walker arrays (NULL: t itself): the sub-blocks with
// several reference chirps are taken with the walkers ordered by a0 (lanes share a reference), the
// long sub-blocks above them with the walkers ordered by their local time (lanes share table rows:
// one shared-memory wavefront per load instead of one per distinct row).  The launch covers sub-blocks
// [blk_lo, blk_hi); CTA y takes blocks_per_chunk of them and writes partial-sum slot blk_lo / blocks_per_chunk + y.
// FINE: the tails.  Thread t is the t-th walker in the order of the last bins; between the last whole
// sub-block and its last bin a walker takes the fine sub-blocks (halves, quarters, ... of the sub-block
// parent_arr[col] it ends in) of the binary decomposition of that stretch -- at most one per level, found
// in closed form -- and evaluates the few bins left after the finest one directly.  CTA y takes levels
// y, y + gridDim.y, ... and writes partial-sum slot y.
#ifndef TABLE_FINE_CTAS
#define TABLE_FINE_CTAS TABLE_MIN_CTAS
#endif
template <int ND, bool FINE, bool CALIB>
__global__ void __launch_bounds__(CTA_THREADS, FINE ? TABLE_FINE_CTAS : TABLE_MIN_CTAS)
loglr_table_kernel(PlanDev P, int W, int wpad, int blocks_per_chunk, int blk_lo, int blk_hi,
                   const TabBlk* __restrict__ blks, const double* __restrict__ cmat,
                   const double* __restrict__ tref, const double2* __restrict__ tab,
                   const double* __restrict__ wc, const int* __restrict__ ke_arr,
                   double2* __restrict__ partial, unsigned char* __restrict__ slow,
                   const int* __restrict__ col_arr,
                   const int* __restrict__ ks_tail_arr, const int* __restrict__ ks_fine_arr,
                   const int* __restrict__ parent_arr) {
    const int t = blockIdx.x * CTA_THREADS + threadIdx.x;
    const int b0 = FINE ? 0 : blk_lo + blockIdx.y * blocks_per_chunk;
    const int b1 = FINE ? 0 : min(b0 + blocks_per_chunk, blk_hi);
    const int chunk = FINE ? (int)blockIdx.y : b0 / blocks_per_chunk;
    const bool live = t < W;
    const int col = live ? (col_arr ? col_arr[t] : t) : 0;
    const int ke = live ? ke_arr[col] : 0;
    // FINE: the stretch [tail_s, tail_s + tail_len) goes to fine sub-blocks
    const int tail_s = (FINE && live) ? ks_tail_arr[col] : 0;
    const int tail_len = (FINE && live) ? ks_fine_arr[col] - tail_s : 0;
    // shared memory: [chunk's Taylor matrices][per-thread coefficient columns][per warp: the dense weight
    // matrix [TAB_SPAN][TAB_WDP] of the tensor-core product and the 4 KB buffer its result comes back through]
    extern __shared__ __align__(16) unsigned char s_tab[];
    double* s_c = reinterpret_cast<double*>(s_tab);
    double* s_coef = s_c + (size_t)blocks_per_chunk * NPH * TAB_NTP;
    double* s_wt = s_coef + TAB_CROWS(ND) * CTA_THREADS + (size_t)(threadIdx.x / WARP) * WARP * TAB_WPITCH;
    if (!FINE)
        for (int i = threadIdx.x; i < (b1 - b0) * NPH * TAB_NTP; i += CTA_THREADS)
            s_c[i] = cmat[(size_t)b0 * NPH * TAB_NTP + i];
    // dense weight matrix [window row][lane] of the tensor-core product: zero outside the lanes' windows
    for (int j = 0; j < TAB_SPAN; ++j) s_wt[j * TAB_WDP + (threadIdx.x & (WARP - 1))] = 0.0;
    int off_prev = -1;
    // per-thread coefficients live in shared memory (one column per thread, no sharing): they are needed
    // once per sub-block, and keeping them out of the register file keeps the tap loop from spilling
    double* sa = s_coef + threadIdx.x;
#define TCOEF(i) sa[(i) * CTA_THREADS]
#pragma unroll
    for (int i = 0; i < NPH; ++i) TCOEF(i) = wc[(size_t)i * wpad + col];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const double* row = wc + (size_t)(C_DET0 + D_NROW * d) * wpad + col;
        TCOEF(NPH + 2 * d) = row[D_TAUHI * wpad];
        TCOEF(NPH + 2 * d + 1) = row[D_TAULO * wpad];
    }
    __syncthreads();
    const int k0 = (FINE || b0 >= b1) ? 0 : blks[b0].start;
    const bool has = FINE ? live : (b0 < b1 && k0 < ke);
    int whi = has ? ke : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
    double2 acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = make_double2(0.0, 0.0);
    const int lane = threadIdx.x & (WARP - 1);
    bool dead = false;                                  // this walker left the table path
    // one sub-block for the lanes with `act` set (a warp-wide operation: every lane calls it)
    auto do_block = [&](const int b, const TabBlk& B, bool act) {
            const int s = B.start, m = B.m, L = 2 * m;
            const unsigned amask = __ballot_sync(0xffffffffu, act);
            if (!amask) return;
            // ---- reference chirp: the nearest one on the sub-block's grid (a function of the walker alone)
            asm volatile("" ::: "memory");               // re-read the coefficient column per sub-block: do not hoist
            int r = (int)((TCOEF(0) - B.a_lo) * B.inv_da);
            r = max(0, min(r, B.K - 1));
            // ---- Taylor coefficients of the phase about the centre: t_n = sum_i a_i c[i][n]
            const double* sc = FINE ? cmat + (size_t)b * NPH * TAB_NTP : s_c + (size_t)(b - b0) * NPH * TAB_NTP;
            double tq[TAB_NT];
#pragma unroll
            for (int n = 0; n < TAB_NT; ++n) tq[n] = 0.0;
#pragma unroll
            for (int i = 0; i < NPH; ++i) {
                const double ai = TCOEF(i);
#pragma unroll
                for (int n = 0; n < TAB_NT; ++n) tq[n] = fma(ai, sc[i * TAB_NTP + n], tq[n]);
            }
            // ---- e_n: Taylor coefficients of exp(U), U = i sum_{k=2..8} w_k s^k, w_k = -2 pi (t_k - t_k^ref):
            //   n e_n = sum_k k (i w_k) e_{n-k};  e_0 = 1, e_1 = 0 (nu carries the slope)
            const double* tr = tref + ((size_t)B.pair0 + r) * TAB_NTP;
            double kw[TAB_DEG + 1];
            double rho = 0.0;
#pragma unroll
            for (int k = 2; k <= TAB_DEG; ++k) {
                const double dt = tq[k] - __ldg(tr + k);
                kw[k] = -G_TWOPI * (double)k * dt;
                rho += fabs(dt);
            }
            double2 en[TAB_DEG + 3];
            en[0] = make_double2(1.0, 0.0);
            en[1] = make_double2(0.0, 0.0);
            if (!CALIB) {
#pragma unroll
                for (int n = 2; n <= TAB_DEG + 2; ++n) {
                    double re = 0.0, im = 0.0;
#pragma unroll
                    for (int k = 2; k <= TAB_DEG; ++k) {
                        if (k <= n) {
                            re = fma(-kw[k], en[n - k].y, re);
                            im = fma(kw[k], en[n - k].x, im);
                        }
                    }
                    en[n] = make_double2(re * (1.0 / n), im * (1.0 / n));
                }
            }
            if (CALIB) {
                // (the series is per detector: see the detector loop; here only the phase's own bound)
                if (act && G_TWOPI * rho > TAB_RHO_MAX) { slow[col] = 1; dead = true; act = false; }
            } else {
                // what the degree-TAB_DEG polynomial leaves out at the sub-block's edge: e_9, e_10, and
                // the phase's own ninth/tenth-order terms (a power law's Taylor coefficients fall off by
                // less than 1.2 eps per order)
                const double u8 = fabs(kw[TAB_DEG]) * (1.0 / TAB_DEG);
                const double est = fabs(en[TAB_DEG + 1].x) + fabs(en[TAB_DEG + 1].y)
                                 + fabs(en[TAB_DEG + 2].x) + fabs(en[TAB_DEG + 2].y)
                                 + 1.2 * B.eps * u8 * (1.0 + 1.2 * B.eps);
                if (act && (est > TAB_EST_MAX || G_TWOPI * rho > TAB_RHO_MAX)) {
                    slow[col] = 1;                      // outside what the tables were built for: slow path
                    dead = true;
                    act = false;
                }
            }
            const double h = 0.5 * (double)m;
            const double kc = (double)(s + m / 2);
            const double2* row0 = tab + B.off;
            // The lanes of a warp are neighbours in a0: nearly always they share the reference, and their
            // 13 taps x 128 B live within a few grid points of each other (the walkers' local times differ
            // by less than the grid spacing times a few).  The lanes of one reference are taken together:
            // the union of their rows is one window of the reference's table, and taps x moments over that
            // window is one small matrix product for the whole warp (below).  A warp that straddles a
            // boundary of the reference grid makes a second pass for the other reference.
            double xs[ND];
            int gs[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double nu = tq[1] / h + (TCOEF(NPH + 2 * d) + TCOEF(NPH + 2 * d + 1));    // turns per bin
                xs[d] = (nu - floor(nu)) * (double)L;                     // [0, L)
                const int g0 = (int)ceil(xs[d] - 0.5 * TAB_W);            // -6 .. L - 6
                gs[d] = g0 < 0 ? g0 + L : g0;
                xs[d] = 2.0 * (xs[d] - (double)g0) - (double)(TAB_W - 1); // v: the Horner variable
            }
            unsigned todo = __ballot_sync(0xffffffffu, act);
            while (todo) {
                const int rL = __shfl_sync(0xffffffffu, r, __ffs(todo) - 1);
                const bool mine = act && r == rL;
                todo &= ~__ballot_sync(0xffffffffu, mine);
                int lo[ND], span_d[ND];
                bool staged[ND];
                __syncwarp();                           // the previous pass's rows have been read
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    int mn = __reduce_min_sync(0xffffffffu, mine ? gs[d] : 0x7fffffff);
                    const int mx = __reduce_max_sync(0xffffffffu, mine ? gs[d] : -1);
                    // the window starts on a multiple of 4 grid points: which of a lane's 13 taps share a step of
                    // the tensor-core product then depends on the lane alone (its first row mod 4), not on its
                    // neighbours -- a walker's result does not depend on the batch it is evaluated in
                    mn &= ~3;
                    lo[d] = mn;
                    const int span = mx - mn + TAB_W;
                    span_d[d] = span;
                    staged[d] = span <= TAB_SPAN;
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    double2 Mn[TAB_NM];
#pragma unroll
                    for (int n = 0; n < TAB_NM; ++n) Mn[n] = make_double2(0.0, 0.0);
                    const double v = xs[d];
                    // kernel weights: per pair of mirror taps an even and an odd Horner chain in v^2,
                    // v = 2 (x - g0) - (W - 1); coefficients as literal constants (no loads)
                    // B fragments of all steps, straight from the table (rows shared by the warp's lanes and by
                    // the neighbouring warps: L1 / L2 hits): in flight while the weights are evaluated.  Rows past
                    // the window's end are table rows too (the allocation is padded) and meet zero weights.
                    double bf[TAB_KPRE][2];
                    const double* pbg = reinterpret_cast<const double*>(row0 + (((size_t)rL * ND + d) * (L + TAB_W) + (staged[d] ? lo[d] : 0)) * TAB_NM)
                                      + (lane & 3) * (2 * TAB_NM) + (lane >> 2);
                    if (staged[d]) {
                        const int nks_ = (span_d[d] + 3) >> 2;
#pragma unroll
                        for (int ks = 0; ks < TAB_KPRE; ++ks) {
                            const int kr = min(ks, nks_ - 1) * 4 * (2 * TAB_NM);
                            bf[ks][0] = __ldg(pbg + kr); bf[ks][1] = __ldg(pbg + kr + 8);
                        }
                    }
                    double2 cfac = make_double2(1.0, 0.0);
                    if (CALIB) {
                        // Calibrated template (gwin/calibration.py:142-173): h_d is multiplied by c_d(f) = exp(L_d),
                        // L_d = log(1 + dA_d) + 2 i atan(dphi_d / 2), dA_d and dphi_d cubics in z (walker_setup).  On the
                        // sub-block z = z_c + eps s: c_d = c_d(z_c) exp(sum_{n>=1} L_n s^n); the L_n (Taylor series of
                        // log(1 + r) and of atan about the centre, by their differential equations -- no transcendental
                        // function) join the phase's u_n = -2 pi i (t_n - t_n^ref), now with a first-order term, and the
                        // series e_n of exp(sum u_n s^n) becomes this detector's own.
                        CalCoef cc;
                        load_cal_coef(cc, wc + (size_t)(C_DET0 + D_NROW * d) * wpad + col, wpad);
                        const double zc = fma(kc, P.cal_zs[d], P.cal_z0[d]);
                        const double ez = h * P.cal_zs[d];
                        const double A0 = fma(fma(fma(cc.a[3], zc, cc.a[2]), zc, cc.a[1]), zc, cc.a[0]);
                        const double P0 = fma(fma(fma(cc.p[3], zc, cc.p[2]), zc, cc.p[1]), zc, cc.p[0]);
                        const double ia = 1.0 / (1.0 + A0);
                        // r = (dA(z) - dA(z_c)) / (1 + dA(z_c)) and x = dphi(z) / 2, as cubics in s
                        double rr[4], xx[4];
                        rr[1] = ez * fma(fma(3.0 * cc.a[3], zc, 2.0 * cc.a[2]), zc, cc.a[1]) * ia;
                        rr[2] = ez * ez * fma(3.0 * cc.a[3], zc, cc.a[2]) * ia;
                        rr[3] = ez * ez * ez * cc.a[3] * ia;
                        xx[0] = 0.5 * P0;
                        xx[1] = 0.5 * ez * fma(fma(3.0 * cc.p[3], zc, 2.0 * cc.p[2]), zc, cc.p[1]);
                        xx[2] = 0.5 * ez * ez * fma(3.0 * cc.p[3], zc, cc.p[2]);
                        xx[3] = 0.5 * ez * ez * ez * cc.p[3];
                        // f = log(1 + r): (1 + r) f' = r'  ->  n f_n = n r_n - sum_{k<n} (n - k) r_k f_{n-k}
                        // g = atan(x):    q g' = x', q = 1 + x^2  ->  n q_0 g_n = n x_n - sum_{0<k<n} (n - k) q_k g_{n-k}
                        double qq[6];
                        qq[0] = fma(xx[0], xx[0], 1.0);
                        qq[1] = 2.0 * xx[0] * xx[1];
                        qq[2] = fma(xx[1], xx[1], 2.0 * xx[0] * xx[2]);
                        qq[3] = 2.0 * fma(xx[0], xx[3], xx[1] * xx[2]);
                        qq[4] = fma(xx[2], xx[2], 2.0 * xx[1] * xx[3]);
                        qq[5] = 2.0 * xx[2] * xx[3];
                        const double iq = 1.0 / qq[0];
                        double ff[7], gg[7];
#pragma unroll
                        for (int n = 1; n <= 6; ++n) {
                            double sf = 0.0, sg = 0.0;
#pragma unroll
                            for (int k = 1; k < n; ++k) {
                                if (k <= 3) sf = fma((double)(n - k) * rr[k], ff[n - k], sf);
                                if (k <= 5) sg = fma((double)(n - k) * qq[k], gg[n - k], sg);
                            }
                            ff[n] = (n <= 3 ? rr[n] : 0.0) - sf * (1.0 / n);
                            gg[n] = ((n <= 3 ? xx[n] : 0.0) - sg * (1.0 / n)) * iq;
                        }
                        // k u_k, k = 1 .. 8 (orders past 5 of L_d: below the truncation bound checked here)
                        double2 ku[TAB_DEG + 1];
#pragma unroll
                        for (int k = 1; k <= TAB_DEG; ++k)
                            ku[k] = make_double2(k <= 5 ? (double)k * ff[k] : 0.0,
                                                 (k <= 5 ? 2.0 * (double)k * gg[k] : 0.0) + (k >= 2 ? kw[k] : 0.0));
#pragma unroll
                        for (int n = 1; n <= TAB_DEG + 2; ++n) {
                            double2 a2 = make_double2(0.0, 0.0);
#pragma unroll
                            for (int k = 1; k <= TAB_DEG; ++k)
                                if (k <= n) a2 = cfma(ku[k], en[n - k], a2);
                            en[n] = make_double2(a2.x * (1.0 / n), a2.y * (1.0 / n));
                        }
                        const double u8 = fabs(kw[TAB_DEG]) * (1.0 / TAB_DEG);
                        const double est = fabs(en[TAB_DEG + 1].x) + fabs(en[TAB_DEG + 1].y)
                                         + fabs(en[TAB_DEG + 2].x) + fabs(en[TAB_DEG + 2].y)
                                         + 1.2 * B.eps * u8 * (1.0 + 1.2 * B.eps)
                                         + fabs(ff[6]) + 2.0 * fabs(gg[6]);
                        // (a detector whose amplitude factor comes near zero, or whose series does not settle: the
                        //  walker goes down the slow path -- the direct kernel takes any values)
                        if (mine && (!(1.0 + A0 > 0.25) || !(est <= TAB_EST_MAX))) { slow[col] = 1; dead = true; }
                        const double P2 = P0 * P0, scf = (1.0 + A0) / (4.0 + P2);
                        cfac = make_double2(scf * (4.0 - P2), scf * (4.0 * P0));
                    }
                    double wt[TAB_W];
                    TAB_WEIGHTS(v, wt);
                    if (staged[d]) {
                        // M[lane][c] = sum_j Wd[j][lane] R[j][c] over the window rows j (span = 13 + the spread of
                        // the lanes' first rows, padded to a multiple of 4; Wd holds a lane's 13 weights at its
                        // own rows and zeros elsewhere; R = the table rows, 16 reals each): a 32 x K x 16 product
                        // on the FP64 tensor cores -- 4 x 2 accumulator tiles of mma.m8n8k4, K / 4 steps.  The
                        // pipe does the same multiply-adds as a DFMA form (one row at a time, broadcast loads) at
                        // its full rate, from 6 loads per step instead of 9 per row and 8 instructions per step
                        // instead of 16 per row.
                        const int nks = (span_d[d] + 3) >> 2;
                        double* wd = s_wt;                               // [TAB_SPAN][TAB_WDP]
                        double* md = s_wt + TAB_SPAN * TAB_WDP;          // [32][2 TAB_NM], 16-byte units swizzled
                        // A lane's column holds its 13 weights at the rows of its window and zeros everywhere else:
                        // only the rows of the previous window that the new one does not cover are cleared.
                        // (The previous detector's A fragments were all loaded before its lanes passed the
                        //  __syncwarp ahead of their moment reads.)
                        const int off = mine ? gs[d] - lo[d] : -2 * TAB_W;
                        if (off_prev >= 0) {
                            const int nz = min(abs(off - off_prev), TAB_W);
                            const int z0 = off > off_prev ? off_prev : off_prev + TAB_W - nz;
                            for (int z = 0; z < nz; ++z) wd[(z0 + z) * TAB_WDP + lane] = 0.0;
                        }
                        if (mine) {
#pragma unroll
                            for (int i = 0; i < TAB_W; ++i) wd[(off + i) * TAB_WDP + lane] = wt[i];
                        }
                        off_prev = mine ? off : -1;
                        __syncwarp();
                        const int l4 = lane & 3, g8 = lane >> 2;
                        const double* pa = wd + l4 * TAB_WDP + g8;
                        double cf[4][2][2];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) cf[mt][nt][0] = cf[mt][nt][1] = 0.0;
#pragma unroll
                        for (int ks = 0; ks < TAB_KPRE; ++ks) {
                            if (ks < nks) {
                                double fa[4];
#pragma unroll
                                for (int mt = 0; mt < 4; ++mt) fa[mt] = pa[ks * 4 * TAB_WDP + 8 * mt];
#pragma unroll
                                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) dmma884(cf[mt][nt], fa[mt], bf[ks][nt]);
                            }
                        }
                        // a window wider than 24 rows (lanes further apart in time): the steps beyond, fetched as they come
#pragma unroll 1
                        for (int ks = TAB_KPRE; ks < nks; ++ks) {
                            double fa[4], fb[2];
                            fb[0] = __ldg(pbg + ks * 4 * (2 * TAB_NM)); fb[1] = __ldg(pbg + ks * 4 * (2 * TAB_NM) + 8);
#pragma unroll
                            for (int mt = 0; mt < 4; ++mt) fa[mt] = pa[ks * 4 * TAB_WDP + 8 * mt];
#pragma unroll
                            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                                for (int nt = 0; nt < 2; ++nt) dmma884(cf[mt][nt], fa[mt], fb[nt]);
                        }
                        // accumulator tile (mt, nt): walker 8 mt + lane / 4, complex moment 4 nt + lane % 4.  Through
                        // shared memory to the walker's own lane; unit (n + f(walker)) mod 8 of the walker's 128-byte
                        // row, f = 4 (w & 1) + ((w >> 1) & 3): conflict-free 128-bit stores and loads.
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            const int wk = 8 * mt + g8;
                            const int fw = 4 * (wk & 1) + ((wk >> 1) & 3);
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt)
                                *reinterpret_cast<double2*>(md + wk * (2 * TAB_NM) + 2 * ((4 * nt + l4 + fw) & 7)) =
                                    make_double2(cf[mt][nt][0], cf[mt][nt][1]);
                        }
                        __syncwarp();
                        {
                            const int fw = 4 * (lane & 1) + ((lane >> 1) & 3);
#pragma unroll
                            for (int n = 0; n < TAB_NM; ++n)
                                Mn[n] = *reinterpret_cast<const double2*>(md + lane * (2 * TAB_NM) + 2 * ((n + fw) & 7));
                        }
                    } else if (mine) {                   // lanes too far apart in time: every lane its own rows, from L2
                        // rows are padded with TAB_W wrap-around points: the taps are contiguous
                        const double2* rd = row0 + (((size_t)rL * ND + d) * (L + TAB_W) + gs[d]) * TAB_NM;
#pragma unroll
                        for (int i = 0; i < TAB_W; ++i) {
                            const double2* e = rd + i * TAB_NM;
#pragma unroll
                            for (int n = 0; n < TAB_NM; ++n) {
                                const double2 q = __ldg(e + n);
                                Mn[n].x = fma(wt[i], q.x, Mn[n].x);
                                Mn[n].y = fma(wt[i], q.y, Mn[n].y);
                            }
                        }
                    }
                    // sum_n e_n M_n; Mn[] holds n = 0, 2, 3, .. TAB_DEG
                    double2 sum = Mn[0];
#pragma unroll
                    for (int n = 2; n <= TAB_DEG; ++n) sum = cfma(en[n], Mn[n - 1], sum);
                    if (CALIB) {
                        // the first moment is not tabulated: M_1 = (i / 2 pi h) dM_0/dnu = (2 i / pi) sum_i w_i' tab_0[g0 + i],
                        // w' the derivative of the interpolation weights with respect to the grid coordinate
                        double dw[TAB_W];
                        TAB_DWEIGHTS(v, dw);
                        double2 m1 = make_double2(0.0, 0.0);
                        if (mine) {
                            const double2* rd = row0 + (((size_t)rL * ND + d) * (L + TAB_W) + gs[d]) * TAB_NM;
#pragma unroll
                            for (int i = 0; i < TAB_W; ++i) {
                                const double2 q = __ldg(rd + i * TAB_NM);
                                m1.x = fma(dw[i], q.x, m1.x);
                                m1.y = fma(dw[i], q.y, m1.y);
                            }
                        }
                        const double c1 = 2.0 / G_PI;
                        sum = cfma(en[1], make_double2(-c1 * m1.y, c1 * m1.x), sum);
                        sum = cmul(sum, cfac);
                    }
                    const double th0 = tq[0] + fma(kc, TCOEF(NPH + 2 * d + 1), frac1(kc * TCOEF(NPH + 2 * d)));
                    if (mine) acc[d] = cfma(sum, phasor_neg(th0), acc[d]);
                }
            }
    };
    if (!FINE) {
        if (b0 < b1 && k0 < whi) {
            for (int b = b0; b < b1; ++b) {
                const TabBlk B = blks[b];
                if (B.start >= whi) break;
                do_block(b, B, !dead && ke >= B.start + B.m);       // a lane takes the sub-block whole
            }
        }
    } else {
        // binary decomposition of the stretch: at the level of sub-blocks of ml bins the walker takes the
        // one at offset (q - 1) ml when q = len / ml is odd
        const int par = live ? parent_arr[col] : -1;
        int pf = 0, pm = 0, pmin = 0x7fffffff;
        if (par >= 0) { pf = blks[par].fine_first; pm = blks[par].m; pmin = blks[par].fine_min; }
        // (the last slot keeps the bins after the finest sub-block; the others share the levels)
        const int lev_slots = max((int)gridDim.y - 1, 1);
        for (int lev = blockIdx.y; lev < 20 && (int)blockIdx.y < lev_slots; lev += lev_slots) {
            const int ml = pm >> (lev + 1);
            const bool deep = par >= 0 && ml >= pmin;
            if (!__any_sync(0xffffffffu, deep)) break;
            int myb = -1;
            if (deep) { const int q = tail_len / ml; if (q & 1) myb = pf + (2 << lev) - 2 + (q - 1); }
            unsigned todo = __ballot_sync(0xffffffffu, myb >= 0);
            while (todo) {
                const int b = __shfl_sync(0xffffffffu, myb, __ffs(todo) - 1);
                const bool same = myb == b;
                todo &= ~__ballot_sync(0xffffffffu, same);
                const TabBlk B = blks[b];
                do_block(b, B, same && !dead);
            }
        }
        if (blockIdx.y == gridDim.y - 1) {
            // the bins after the finest sub-block (fewer than its length): per-bin phase and sincos, every
            // lane its own bins
            const int r0 = tail_s + tail_len;
            const int nres = par >= 0 ? ke - r0 : 0;
            int nmax = nres;
#pragma unroll
            for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
            if (nmax > 0) {
                PhaseCoef pc;
                pc.a0 = TCOEF(C_A0); pc.a2 = TCOEF(C_A2); pc.a3 = TCOEF(C_A3); pc.a4 = TCOEF(C_A4); pc.a5 = TCOEF(C_A5);
                pc.b5 = TCOEF(C_B5); pc.a6 = TCOEF(C_A6); pc.b6 = TCOEF(C_B6); pc.a7 = TCOEF(C_A7);
                pc.a10 = TCOEF(C_A10); pc.a12 = TCOEF(C_A12);
                // (branch-free, so that the loads of several bins are in flight at once: every lane walks its
                //  own bins, nothing is coalesced, and the loop would otherwise wait out one L2 latency per bin)
                const int kpark = max(P.kmin, 1);
                CalCoef ccd[ND];
                if (CALIB) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) load_cal_coef(ccd[d], wc + (size_t)(C_DET0 + D_NROW * d) * wpad + col, wpad);
                }
#pragma unroll 4
                for (int i = 0; i < nmax; ++i) {
                    const bool on = i < nres && !dead;
                    const int k = on ? r0 + i : kpark;
                    const double kd = (double)k;
                    const double psi = phase_turns(pc, ld_basis(P.basis + k));
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        double2 u = phasor_neg(psi + fma(kd, TCOEF(NPH + 2 * d + 1), frac1(kd * TCOEF(NPH + 2 * d))));
                        if (CALIB) u = cmul(u, calib_factor(ccd[d], fma(kd, P.cal_zs[d], P.cal_z0[d])));
                        double2 Tk = __ldg(P.T + (size_t)k * ND + d);
                        if (!on) Tk = make_double2(0.0, 0.0);
                        acc[d] = cfma(Tk, u, acc[d]);
                    }
                }
            }
        }
    }
    if (live && has) {          // (FINE: every slot is written)
#pragma unroll
        for (int d = 0; d < ND; ++d)
            partial[((size_t)chunk * ND + d) * wpad + col] = acc[d];
    }
#undef TCOEF
}

// Tails of the table path.  Walkers are taken in the order of their last bin: perm_e[t] is the t-th
// walker in that order, inv[] its column in the processing order; col_e[t] keeps the composition.  The
// 128 walkers of a CTA then need only a short run of recurrence blocks: grp_b[g] .. grp_b[groups + g].
__global__ void __launch_bounds__(CTA_THREADS)
tail_plan_kernel(int W, int n_blk, const int* __restrict__ blk_start, const int* __restrict__ perm_e,
                 const int* __restrict__ inv, int* __restrict__ col_e,
                 const int* __restrict__ ks_fine, const int* __restrict__ ke_fast, int* __restrict__ grp_b,
                 const int* __restrict__ perm_t, int* __restrict__ col_t,
                 const int* __restrict__ tail_parent, int* __restrict__ grp_p) {
    __shared__ int s_lo[CTA_THREADS / WARP], s_hi[CTA_THREADS / WARP], s_pl[CTA_THREADS / WARP], s_ph[CTA_THREADS / WARP];
    const int t = blockIdx.x * CTA_THREADS + threadIdx.x;
    int lo = 0x7fffffff, hi = 0, pl = 0x7fffffff, ph = -1;
    if (t < W) {
        const int col = inv[perm_e[t]];
        col_e[t] = col;
        if (perm_t) col_t[t] = inv[perm_t[t]];
        const int a = ks_fine[col], e = ke_fast[col];
        if (a >= 0 && a < e) { lo = a; hi = e; }
        const int par = tail_parent[col];
        if (par >= 0) { pl = par; ph = par; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        pl = min(pl, __shfl_xor_sync(0xffffffffu, pl, o));
        ph = max(ph, __shfl_xor_sync(0xffffffffu, ph, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x / WARP] = lo; s_hi[threadIdx.x / WARP] = hi;
        s_pl[threadIdx.x / WARP] = pl; s_ph[threadIdx.x / WARP] = ph;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int i = 1; i < CTA_THREADS / WARP; ++i) {
        lo = min(lo, s_lo[i]); hi = max(hi, s_hi[i]);
        pl = min(pl, s_pl[i]); ph = max(ph, s_ph[i]);
    }
    grp_p[blockIdx.x] = pl;                             // pl > ph: nobody in this group has fine sub-blocks
    grp_p[gridDim.x + blockIdx.x] = ph;
    int b_lo = 0, b_hi = 0;
    if (lo < hi) {
        int x = 0, y = n_blk - 1;                       // last block starting at or below lo
        while (x < y) { const int mid = (x + y + 1) >> 1; if (blk_start[mid] <= lo) x = mid; else y = mid - 1; }
        b_lo = x;
        x = b_lo; y = n_blk;                            // first boundary at or above hi
        while (x < y) { const int mid = (x + y) >> 1; if (blk_start[mid] >= hi) y = mid; else x = mid + 1; }
        b_hi = x;
    }
    grp_b[blockIdx.x] = b_lo;
    grp_b[gridDim.x + blockIdx.x] = b_hi;
}

// Slow path: the list of walkers (columns) the fast kernels did not take.
__global__ void slow_compact_kernel(int W, const unsigned char* __restrict__ slow, int* __restrict__ list,
                                    int* __restrict__ n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W || !slow[t]) return;
    list[atomicAdd(n, 1)] = t;
    atomicAdd(n + 1, 1);
}

// ---------------------------------------------------------------------------
// kernel 3: combine chunk partials, scale by the walker amplitude, form loglr,
// scatter back to the caller's walker order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double log_i0(double x) {
    // log I0(x); the reference's numpy.log(special.i0(x)) (marginalized_gaussian_noise.py:459)
    // overflows for x >~ 713 -- identical below, finite above.
    if (x < 400.0) return log(cyl_bessel_i0(x));
    const double y = 1.0 / (8.0 * x);
    // I0(x) ~ e^x/sqrt(2 pi x) * sum_n ((2n-1)!!)^2 / n! * y^n
    const double ser = 1.0 + y * (1.0 + y * (4.5 + y * (37.5 + y * (459.375 + y * (7441.875 + y * 150077.8125)))));
    return x - 0.5 * log(G_TWOPI * x) + log(ser);
}

// One walker's result from its per-detector sums: scale by the walker amplitude, form loglr, scatter
// to the caller's walker order.
__device__ __forceinline__ void finish_walker(int n_det, int wpad, int mode, int t, int w, bool invalid,
                                              const double2* __restrict__ sums, const double* __restrict__ wc,
                                              const double* __restrict__ hh_sorted,
                                              double* __restrict__ loglr, double* __restrict__ hd_out,
                                              double* __restrict__ hh_out, double* __restrict__ marg_opt) {
    double lr = 0.0, sre = 0.0, sim = 0.0, shh = 0.0;
    for (int d = 0; d < n_det; ++d) {
        const double re = sums[d].x, im = sums[d].y;
        const double* row = wc + (size_t)(C_DET0 + D_NROW * d) * wpad + t;
        const double are = row[D_ARE * wpad], aim = row[D_AIM * wpad];
        const double hre = are * re - aim * im;
        const double him = are * im + aim * re;
        const double hh = hh_sorted[d * wpad + t];
        if (hd_out) {
            hd_out[((size_t)w * n_det + d) * 2 + 0] = invalid ? 0.0 : hre;
            hd_out[((size_t)w * n_det + d) * 2 + 1] = invalid ? 0.0 : him;
        }
        if (hh_out) hh_out[(size_t)w * n_det + d] = invalid ? 0.0 : hh;
        lr += hre - 0.5 * hh;
        sre += hre; sim += him; shh += hh;
    }
    if (mode == GWIN_MODE_MARG_PHASE) lr = log_i0(hypot(sre, sim)) - 0.5 * shh;
    if (mode >= GWIN_MODE_MARG_DIST) {
        // hand (x, opt) to marg_dist_kernel: x travels in loglr[w]
        const double mf = hypot(sre, sim);
        lr = mode == GWIN_MODE_MARG_DIST ? mf : log_i0(mf);
        marg_opt[w] = shh;
    }
    if (invalid || isnan(lr)) lr = -INFINITY;
    loglr[w] = lr;
}

#define COMB_SEG 16     // chunk segments summed in parallel per walker (fixed order: deterministic)
__global__ void __launch_bounds__(32 * COMB_SEG)
combine_kernel(int n_det, int W, int wpad, int n_chunks, int mode,
               int per_chunk, int n_blk, const int* __restrict__ blk_start, int band0, int kmax,
               const int* __restrict__ perm, const double* __restrict__ wc,
               const int* __restrict__ ks_arr, const int* __restrict__ ke_arr,
               const double2* __restrict__ partial,
               const double* __restrict__ hh_sorted,
               double* __restrict__ loglr, double* __restrict__ hd_out,
               double* __restrict__ hh_out, double* __restrict__ marg_opt,
               int n_chunks2, int per_chunk2, const TabBlk* __restrict__ tblk,
               const double2* __restrict__ partial2, const int* __restrict__ ke2_arr,
               int n_chunks3, const double2* __restrict__ partial3,
               int n_chunks4, const double2* __restrict__ partial4) {
    __shared__ double2 s_part[COMB_SEG][GWIN_MAX_DET][32];
    const int lx = threadIdx.x, seg = threadIdx.y;
    const int t = blockIdx.x * 32 + lx;
    const bool live = t < W;
    // stage 1: thread (lx, seg) sums chunks seg, seg + COMB_SEG, ... of walker t (coalesced in t)
    // (only the chunks that overlap the walker's band [ks, ke) were written by the main kernel;
    //  blk_start != NULL: chunk c = blocks [c per_chunk, (c+1) per_chunk), else per_chunk bins)
    const int ks = live ? ks_arr[t] : 0, ke = live ? ke_arr[t] : 0;
    for (int d = 0; d < n_det; ++d) {
        double re = 0.0, im = 0.0;
        if (live)
            for (int c = seg; c < n_chunks; c += COMB_SEG) {
                int k0, k1;
                if (blk_start) { k0 = blk_start[c * per_chunk]; k1 = blk_start[min((c + 1) * per_chunk, n_blk)]; }
                else { k0 = band0 + c * per_chunk; k1 = min(k0 + per_chunk, kmax); }
                if (max(k0, ks) >= min(k1, ke)) continue;
                const double2 p = partial[((size_t)c * n_det + d) * wpad + t];
                re += p.x; im += p.y;
            }
        // table-kernel chunks (above k_split): written where the chunk starts below the walker's end
        if (live && n_chunks2 > 0) {
            const int ke2 = ke2_arr[t];
            for (int c = seg; c < n_chunks2; c += COMB_SEG) {
                if (tblk[c * per_chunk2].start >= ke2) continue;
                const double2 p = partial2[((size_t)c * n_det + d) * wpad + t];
                re += p.x; im += p.y;
            }
            // the tail between the last whole sub-block and the walker's last bin, done by the second
            // pass of the recurrence kernel: every slot is written
            for (int c = seg; c < n_chunks3; c += COMB_SEG) {
                const double2 p = partial3[((size_t)c * n_det + d) * wpad + t];
                re += p.x; im += p.y;
            }
            // ... and of the fine sub-blocks before it
            for (int c = seg; c < n_chunks4; c += COMB_SEG) {
                const double2 p = partial4[((size_t)c * n_det + d) * wpad + t];
                re += p.x; im += p.y;
            }
        }
        s_part[seg][d][lx] = make_double2(re, im);
    }
    __syncthreads();
    if (seg != 0 || !live) return;
    // stage 2: one thread per walker finishes
    double2 sums[GWIN_MAX_DET];
    for (int d = 0; d < n_det; ++d) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int q = 0; q < COMB_SEG; ++q) { re += s_part[q][d][lx].x; im += s_part[q][d][lx].y; }
        sums[d] = make_double2(re, im);
    }
    finish_walker(n_det, wpad, mode, t, perm ? perm[t] : t, ks_arr[t] < 0, sums, wc, hh_sorted,
                  loglr, hd_out, hh_out, marg_opt);
}

// Slow path: the direct kernel's chunk sums of the listed walkers, in chunk order; overwrites what
// combine_kernel wrote for them.
__global__ void slow_finish_kernel(int n_det, int wpad, int n_chunks, int mode, int per_chunk, int band0, int kmax,
                                   const int* __restrict__ perm, const double* __restrict__ wc,
                                   const int* __restrict__ ks_arr, const int* __restrict__ ke_arr,
                                   const double2* __restrict__ partial, const double* __restrict__ hh_sorted,
                                   const int* __restrict__ list, const int* __restrict__ n_list,
                                   double* __restrict__ loglr, double* __restrict__ hd_out,
                                   double* __restrict__ hh_out, double* __restrict__ marg_opt) {
    const int n = *n_list;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int t = list[i];
        const int ks = ks_arr[t], ke = ke_arr[t];
        double2 sums[GWIN_MAX_DET];
        for (int d = 0; d < n_det; ++d) {
            double re = 0.0, im = 0.0;
            for (int c = 0; c < n_chunks; ++c) {
                const int k0 = band0 + c * per_chunk, k1 = min(k0 + per_chunk, kmax);
                if (max(k0, ks) >= min(k1, ke)) continue;
                const double2 p = partial[((size_t)c * n_det + d) * wpad + i];
                re += p.x; im += p.y;
            }
            sums[d] = make_double2(re, im);
        }
        finish_walker(n_det, wpad, mode, t, perm ? perm[t] : t, false, sums, wc, hh_sorted,
                      loglr, hd_out, hh_out, marg_opt);
    }
}

// ---------------------------------------------------------------------------
// kernel 4 (distance marginalisation only): one warp per walker,
//   loglr = log sum_j w_j exp(x u_j - opt u_j^2 / 2),  u_j = 1 / D_j
// (marginalized_gaussian_noise.py:431-448; scipy.special.logsumexp(a, b=w)).  Two passes
// over the grid: the maximum of a_j + log w_j, then the sum of exponentials below it.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
marg_dist_kernel(int W, int n, const double* __restrict__ u, const double* __restrict__ logw,
                 const double* __restrict__ opt_arr, double* __restrict__ loglr) {
    const int w = blockIdx.x * (blockDim.x / WARP) + threadIdx.x / WARP;
    const int lane = threadIdx.x % WARP;
    if (w >= W) return;
    const double x = loglr[w];
    if (!(x > -INFINITY)) return;                  // no waveform / skipped: stays -inf
    const double mh = -0.5 * opt_arr[w];
    double m = -INFINITY;
    for (int j = lane; j < n; j += WARP) {
        const double uj = __ldg(u + j);
        m = fmax(m, fma(fma(mh, uj, x), uj, __ldg(logw + j)));
    }
#pragma unroll
    for (int o = WARP / 2; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    double s = 0.0;
    if (m > -INFINITY)
        for (int j = lane; j < n; j += WARP) {
            const double uj = __ldg(u + j);
            const double a = fma(fma(mh, uj, x), uj, __ldg(logw + j)) - m;
            if (a > -60.0) s += exp(a);            // below e^-60 of the largest term: < 1e-22 relative
        }
#pragma unroll
    for (int o = WARP / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) loglr[w] = m > -INFINITY ? m + log(s) : -INFINITY;
}

// ---------------------------------------------------------------------------
// waveform dump (FDomainDetFrameGenerator.generate for one point)
// ---------------------------------------------------------------------------
__global__ void generate_kernel(PlanDev P, int wpad, const double* __restrict__ wc,
                                const int* __restrict__ ke_arr, double2* __restrict__ out, bool calibrated) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.n_f) return;
    PhaseCoef pc;
    load_phase_coef(pc, wc, wpad, 0);
    // un-clipped waveform length is stored in ke_arr[1]
    const int n = ke_arr[1];
    for (int d = 0; d < P.n_det; ++d) {
        double2 h = make_double2(0.0, 0.0);
        if (k >= P.istart && k < n && k > 0) {
            const double* row = wc + (size_t)(C_DET0 + D_NROW * d) * wpad;
            const double kd = (double)k;
            const double ph = phase_turns(pc, ld_basis(P.basis + k)) +
                              fma(kd, row[D_TAULO * wpad], frac1(kd * row[D_TAUHI * wpad]));
            const double2 u = phasor_neg(ph);
            const double f = kd * P.delta_f;
            const double a = pow(f, -7.0 / 6.0);
            const double2 A = make_double2(row[D_ARE * wpad] * a, row[D_AIM * wpad] * a);
            h = cmul(A, u);
            if (calibrated) {
                CalCoef cc;
                load_cal_coef(cc, row, wpad);
                h = cmul(h, calib_factor(cc, fma(kd, P.cal_zs[d], P.cal_z0[d])));
            }
        }
        out[(size_t)d * P.n_f + k] = h;
    }
}
__global__ void generate_len_kernel(PlanDev P, const double* __restrict__ params, int* __restrict__ ke_arr) {
    ke_arr[1] = (params_ok(params) && has_waveform(params, P.f_lower_wf)) ? waveform_len(params[0] + params[1], P.delta_f) : 0;
}

// FP64 issue-rate probes.  variant 0: 8 independent DFMA chains sharing the multiplier and
// addend registers (operand-reuse friendly, the classical "peak" loop); variant 1: every DFMA
// reads three registers no neighbouring instruction reads (no reuse-cache help); variant 2:
// the hot loop's own pattern, 4 independent unit-phasor complex multiplications per step
// (2 DMUL + 2 DFMA each).
template <int VARIANT>
__global__ void dfma_probe_kernel(double* out, int iters) {
    const double t0 = threadIdx.x * 1e-9;
    if (VARIANT == 3 || VARIANT == 4) {
        // 3: eight independent FP64 tensor-core MMAs (m8n8k4) per iteration; 4: the same interleaved with
        // eight independent DFMAs (do the two pipes overlap?)
        double d[8][2], f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { d[i][0] = t0 + i; d[i][1] = t0 - i; f[i] = t0 + 0.5 * i; }
        const double a = 0.999999 + t0, b = 1e-9 + t0, m = 0.999999999, c = 1e-12;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                dmma884(d[q], a, b);
                if (VARIANT == 4) f[q] = fma(f[q], m, c);
            }
        }
        double r = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += d[i][0] + d[i][1] + f[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
        return;
    }
    if (VARIANT == 0) {
        double a0 = t0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
        const double m = 0.999999999, c = 1e-12;
        for (int i = 0; i < iters; ++i) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else if (VARIANT == 1) {
        double a[8], b[8], c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = t0 + i; b[i] = 0.999999999 - 1e-12 * i - t0; c[i] = 1e-12 * (i + 1) + t0; }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = fma(a[q], b[q], c[q]);
        }
        double r = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {
        double2 z[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { z[i] = make_double2(1.0 - t0, t0 + i * 1e-9); w[i] = make_double2(cos(1e-3 * (i + 1) + t0), sin(1e-3 * (i + 1) + t0)); }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) z[q] = cmul(z[q], w[q]);
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = z[0].x + z[1].y + z[2].x + z[3].y;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static const int64_t LEAPS_GPS[] = {
    46828800, 78364801, 109900802, 173059203, 252028804, 315187205, 346723206,
    394156807, 425692808, 457228809, 504489610, 551750411, 599184012, 820108813,
    914803214, 1025136015, 1119744016, 1167264017 };

// Linear map (f_1-f_0, ..., f_deg-f_0) -> forward differences of the interpolating polynomial
// through nodes 0 = x_0 < x_1 < ... < x_deg = B-1:  Newton divided differences -> monomial
// coefficients c1..c4 -> D1(0) = c1+c2+c3+c4, D2(1) = 2c2+6c3+14c4, D3(2) = 6c3+36c4, D4 = 24c4.
static BlockMat make_block_mat(int B, int deg) {
    BlockMat M;
    memset(&M, 0, sizeof M);
    M.deg = deg;
    if (deg == 0) return M;
    const int e = B - 1;
    int x[5] = {0, 0, 0, 0, 0};
    if (deg == 4) { const int q = (e + 4) / 8; x[1] = q; x[2] = e / 2; x[3] = e - q; x[4] = e; }   // ~Chebyshev-Lobatto
    else          { const int q = (e + 2) / 4; x[1] = q; x[2] = e - q; x[3] = e; }
    for (int i = 0; i < deg; ++i) M.node[i] = x[i + 1];
    for (int col = 1; col <= deg; ++col) {
        long double dd[5] = {0, 0, 0, 0, 0};
        dd[col] = 1.0L;                         // samples g_i = f_i - f_0 of the unit vector
        long double nw[5];                      // Newton coefficients
        nw[0] = dd[0];
        for (int lvl = 1; lvl <= deg; ++lvl) {
            for (int i = 0; i + lvl <= deg; ++i)
                dd[i] = (dd[i + 1] - dd[i]) / (long double)(x[i + lvl] - x[i]);
            nw[lvl] = dd[0];
        }
        // Newton -> monomial: P(j) = sum_l nw[l] prod_{i<l} (j - x_i)
        long double c[5] = {0, 0, 0, 0, 0}, basis[6] = {1, 0, 0, 0, 0, 0};
        for (int lvl = 0; lvl <= deg; ++lvl) {
            for (int q = 0; q <= lvl; ++q) c[q] += nw[lvl] * basis[q];
            // basis *= (j - x_lvl)
            long double nb[6] = {0, 0, 0, 0, 0, 0};
            for (int q = 0; q <= lvl; ++q) { nb[q + 1] += basis[q]; nb[q] -= basis[q] * (long double)x[lvl]; }
            for (int q = 0; q < 6; ++q) basis[q] = nb[q];
        }
        const long double c1 = c[1], c2 = c[2], c3 = c[3], c4 = deg == 4 ? c[4] : 0.0L;
        M.m[4 * 0 + col - 1] = (double)(c1 + c2 + c3 + c4);
        M.m[4 * 1 + col - 1] = (double)(2 * c2 + 6 * c3 + 14 * c4);
        M.m[4 * 2 + col - 1] = (double)(6 * c3 + 36 * c4);
        M.m[4 * 3 + col - 1] = (double)(24 * c4);
    }
    return M;
}

static void build_schedule(gwin_plan* p) {
    // Blocks over [band0, kmax).  Inside a block the chirp phase is replaced by the polynomial
    // through deg+1 exact samples at (rounded) Chebyshev-Lobatto nodes; with the leading-order
    // chirp Psi(f) = 3/128 (pi Mc f)^(-5/3) the interpolation error is bounded by
    //   quartic: |Psi^(5)|/5! * 0.005    (B-1)^5      cubic: |Psi^(4)|/4! * 0.015625 (B-1)^4
    // (x2 safety for the higher PN orders).  Each block takes whichever degree is cheaper per
    // bin: the cubic saves one complex multiply per bin but needs shorter blocks.
    const PlanDev& d = p->d;
    const int band0 = std::max(d.kmin, d.istart);
    p->blk_start.clear(); p->blk_mat.clear(); p->mats.clear();
    std::vector<int> mat_of((MAX_BLOCK + MIN_BLOCK + 8) * 2, -1);      // (block length, degree) -> mats index
    auto mat_index = [&](int B, int deg) {
        if (deg == 0) { B = 0; }
        int& slot = mat_of[(size_t)B * 2 + (deg == 3 ? 1 : 0)];
        if (slot < 0) { slot = (int)p->mats.size(); p->mats.push_back(make_block_mat(B, deg)); }
        return slot;
    };
    const double K4 = (5.0 / 3) * (8.0 / 3) * (11.0 / 3) * (14.0 / 3), K5 = K4 * (17.0 / 3);
    const double chirp = 2.0 * (3.0 / 128.0) * pow(G_PI * p->mchirp_min * G_MTSUN, -5.0 / 3.0);
    const double setup_cost = 450.0, bin4 = 80.0, bin3 = 71.0;      // FP64-pipe cycles (ncu r1b)
    int k = std::max(band0, 1);
    while (k < d.kmax) {
        const double f = k * d.delta_f;
        const double d5 = chirp * K5 * pow(d.delta_f, 5.0) * pow(f, -20.0 / 3.0);   // rad / bin^5
        const double d4 = chirp * K4 * pow(d.delta_f, 4.0) * pow(f, -17.0 / 3.0);   // rad / bin^4
        const double B4 = std::min<double>(pow(p->phase_tol * 120.0 / (0.005 * d5), 0.2) + 1.0, MAX_BLOCK);
        const double B3 = std::min<double>(pow(p->phase_tol * 24.0 / (0.015625 * d4), 0.25) + 1.0, MAX_BLOCK);
        int deg = 4, Bi = (int)B4;
        if (B3 >= MIN_BLOCK && bin3 + setup_cost / B3 < bin4 + setup_cost / B4) { deg = 3; Bi = (int)B3; }
        if (Bi >= MIN_BLOCK) {
            Bi &= ~3;
        } else {
            // Too curved for an 8-bin block.  A 5-bin quartic has every bin as a node and is exact;
            // 6 or 7 bins are taken when the one or two interpolated bins still meet the bound
            // (exact node polynomial instead of the Chebyshev estimate).
            deg = 4;
            for (Bi = 7; Bi > 5; --Bi) {
                const int e = Bi - 1, q = (e + 4) / 8;
                const int x[5] = {0, q, e / 2, e - q, e};
                double worst = 0.0;
                for (int j = 0; j <= e; ++j) {
                    double prod = 1.0;
                    for (int i = 0; i < 5; ++i) prod *= (double)(j - x[i]);
                    worst = std::max(worst, std::fabs(prod));
                }
                if (d5 / 120.0 * worst <= p->phase_tol) break;
            }
        }
        if (k + Bi > d.kmax) Bi = d.kmax - k;
        if (Bi < 5) deg = 0;                                        // band remainder: direct evaluation
        p->blk_start.push_back(k);
        p->blk_mat.push_back(mat_index(Bi, deg));
        k += Bi;
    }
    p->blk_start.push_back(d.kmax);
    p->n_blk = (int)p->blk_mat.size();
}

static int upload_schedule(gwin_plan* p) {
    build_schedule(p);
    cudaFree(p->d_blk_start); cudaFree(p->d_blk_mat); cudaFree(p->d_mats);
    p->d_blk_start = p->d_blk_mat = nullptr; p->d_mats = nullptr;
    CK(cudaMalloc(&p->d_blk_start, sizeof(int) * p->blk_start.size()));
    CK(cudaMalloc(&p->d_blk_mat, sizeof(int) * (p->blk_mat.size() + 1)));
    CK(cudaMalloc(&p->d_mats, sizeof(BlockMat) * p->mats.size()));
    CK(cudaMemcpy(p->d_blk_start, p->blk_start.data(), sizeof(int) * p->blk_start.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_blk_mat, p->blk_mat.data(), sizeof(int) * p->blk_mat.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_mats, p->mats.data(), sizeof(BlockMat) * p->mats.size(), cudaMemcpyHostToDevice));
    p->sched_dirty = false;
    p->tab_dirty = true;
    return GWIN_OK;
}


// ---- table kernel: coverage, schedule, moment tables ------------------------------------------
static const std::vector<int>& table_sizes() {            // 2^k and 3 2^(k-1) in [TAB_MIN, TAB_MAX]
    static const std::vector<int> sizes = [] {
        std::vector<int> v;
        for (int m = TAB_MIN; m <= TAB_MAX; m *= 2) {
            v.push_back(m);
            if (m + m / 2 <= TAB_MAX) v.push_back(m + m / 2);
        }
        std::sort(v.begin(), v.end());
        return v;
    }();
    return sizes;
}

static double es_kernel_host(long double x) {
    const long double z = 2.0L * x / TAB_W;
    if (fabsl(z) >= 1.0L) return 0.0;
    return (double)expl((long double)TAB_BETA * (sqrtl(1.0L - z * z) - 1.0L));
}

// 1 / phi_hat(j_c / L), j_c = j - m/2, phi_hat by midpoint quadrature of the interpolation kernel over
// its support (depends on m only: computed once per process)
static const std::vector<double>& inv_hat(int m) {
    static std::vector<double> cache[64];
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    const std::vector<int>& sizes = table_sizes();
    const size_t si = std::find(sizes.begin(), sizes.end(), m) - sizes.begin();
    std::vector<double>& hc = cache[si];
    if (hc.empty()) {
        const int nq = 2048, L = 2 * m;
        std::vector<long double> xs(nq), ph(nq);
        for (int q = 0; q < nq; ++q) {
            xs[q] = ((long double)q + 0.5L) / nq * TAB_W - 0.5L * TAB_W;
            ph[q] = es_kernel_host(xs[q]);
        }
        hc.resize(m);
        for (int j = 0; j < m; ++j) {
            const long double xi = (long double)(j - m / 2) / L;
            long double sum = 0.0L;
            for (int q = 0; q < nq; ++q) sum += ph[q] * cosl(2.0L * 3.14159265358979323846264338327950288L * xi * xs[q]);
            hc[j] = (double)(1.0L / (sum * TAB_W / nq));
        }
    }
    return hc;
}

// Taylor coefficients, in s = (k - kc) / h, of the 11 functions of f the TaylorF2 phase is made of
// (the order of the C_A* rows): x^p for p = -5, -3, -2, -1, 0, (ln x), 1, (x ln x), 2, 5, 7, x = f^(1/3).
static void taylor_matrix(double kc, double h, double delta_f, long double c[NPH][TAB_NTH]) {
    static const int pw[NPH] = {-5, -3, -2, -1, 0, 0, 1, 1, 2, 5, 7};
    static const int lg[NPH] = {0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0};
    const long double fc = (long double)kc * (long double)delta_f, eps = (long double)h / (long double)kc;
    long double ls[TAB_NTH];                               // ln x = ln(fc)/3 + ln(1 + eps s)/3
    ls[0] = logl(fc) / 3.0L;
    long double en = 1.0L;
    for (int n = 1; n < TAB_NTH; ++n) { en *= eps; ls[n] = ((n & 1) ? en : -en) / n / 3.0L; }
    for (int i = 0; i < NPH; ++i) {
        const long double q = (long double)pw[i] / 3.0L;
        long double bs[TAB_NTH];                           // fc^q (1 + eps s)^q
        bs[0] = powl(fc, q);
        for (int n = 0; n + 1 < TAB_NTH; ++n) bs[n + 1] = bs[n] * (q - n) / (n + 1) * eps;
        for (int n = 0; n < TAB_NTH; ++n) {
            if (!lg[i]) { c[i][n] = bs[n]; continue; }
            long double acc = 0.0L;
            for (int k = 0; k <= n; ++k) acc += bs[k] * ls[n - k];
            c[i][n] = acc;
        }
    }
}

// Majorant of what the degree-TAB_DEG polynomial of e^{-2 pi i R} drops, for a walker da_half away (in
// a0) from its reference chirp and dev[i] away in the other phase coefficients.
static bool table_admissible(const long double c[NPH][TAB_NTH], const double* dev, double da_half) {
    long double u[TAB_NTH], e[TAB_NTH];
    long double rho = 0.0L;
    for (int k = 2; k < TAB_NTH; ++k) {
        long double acc = (long double)da_half * fabsl(c[0][k]);
        for (int i = 1; i < NPH; ++i) acc += (long double)dev[i] * fabsl(c[i][k]);
        u[k] = 2.0L * 3.14159265358979323846264338327950288L * acc;
        if (k <= TAB_DEG) rho += u[k];
    }
    e[0] = 1.0L; e[1] = 0.0L;
    long double est = 0.0L;
    for (int n = 2; n < TAB_NTH; ++n) {
        long double acc = 0.0L;
        for (int k = 2; k <= n; ++k) acc += k * u[k] * e[n - k];
        e[n] = acc / n;
        if (n > TAB_DEG) est += e[n];
    }
    return est <= TAB_EST_DESIGN && rho <= TAB_RHO_DESIGN;
}

// References needed by a sub-block [k, k + m) for the coverage C (0: not admissible at any spacing)
static int table_refs_needed(const gwin_plan* p, const Cover& C, int k, int m, int kmax_refs) {
    long double c[NPH][TAB_NTH];
    taylor_matrix((double)(k + m / 2), 0.5 * m, p->d.delta_f, c);
    const double half = 0.5 * (C.a_hi - C.a_lo);
    if (table_admissible(c, C.dev, half)) return 1;
    if (!table_admissible(c, C.dev, half / kmax_refs)) return 0;
    double lo = half / kmax_refs, hi = half;            // lo admissible, hi not
    for (int it = 0; it < 14; ++it) {
        const double mid = std::sqrt(lo * hi);
        if (table_admissible(c, C.dev, mid)) lo = mid; else hi = mid;
    }
    return std::min(kmax_refs, (int)std::ceil(half / lo));
}

static void free_tables(gwin_plan* p) {
    cudaFree(p->d_tblk); cudaFree(p->d_cmat); cudaFree(p->d_tref); cudaFree(p->d_tab);
    p->d_tblk = nullptr; p->d_cmat = p->d_tref = nullptr; p->d_tab = nullptr;
    p->n_tblk = 0; p->n_fine = 0; p->n_pairs = 0; p->tab_bytes = 0;
    p->tblk.clear();
}

// Sub-blocks over [k_split, kmax) for the coverage `C`, longest first, with at most K0 (f/f0)^(-11/9)
// reference chirps per sub-block (the split of a memory budget that minimises the number of
// sub-blocks when the reference count grows like m^2 f^(-11/3)).  Returns the table bytes.
static double table_schedule(const gwin_plan* p, const Cover& C, double K0, std::vector<TabBlk>& out, int* ib_split) {
    const PlanDev& d = p->d;
    const std::vector<int>& sizes = table_sizes();
    const int band0 = std::max(std::max(d.kmin, d.istart), 1);
    const double f0 = band0 * d.delta_f;
    out.clear();
    static const int kmax_refs = 4096;
    auto pick = [&](int k, int* K_out) -> int {          // longest admissible size at bin k within the cap
        const double cap = std::max(1.0, K0 * std::pow(k * d.delta_f / f0, -11.0 / 9.0));
        for (int si = (int)sizes.size() - 1; si >= 0; --si) {
            const int m = sizes[si];
            if (k + m > d.kmax) continue;
            const int K = table_refs_needed(p, C, k, m, kmax_refs);
            if (K > 0 && K <= cap) { *K_out = K; return m; }
        }
        return 0;
    };
    // k_split: the first boundary of the recurrence schedule from which sub-blocks qualify
    int ib = 0, K = 0, m = 0;
    while (ib < p->n_blk && (m = pick(p->blk_start[ib], &K)) == 0) {
        // skip ahead by at least 64 bins
        const int k_next = p->blk_start[ib] + 64;
        while (ib < p->n_blk && p->blk_start[ib] < k_next) ++ib;
    }
    *ib_split = ib;
    if (ib >= p->n_blk) return 0.0;
    double bytes = 0.0;
    int pair0 = 0;
    long long off = 0;
    for (int k = p->blk_start[ib]; k < d.kmax;) {
        m = pick(k, &K);
        if (m == 0) break;                                // what is left below kmax goes to the tails
        TabBlk B;
        memset(&B, 0, sizeof B);
        B.start = k; B.m = m; B.K = K; B.pair0 = pair0;
        B.da = (C.a_hi - C.a_lo) / K;
        B.a_lo = C.a_lo;
        B.inv_da = 1.0 / B.da;
        B.eps = 0.5 * m / (double)(k + m / 2);
        B.off = off;
        out.push_back(B);
        pair0 += K;
        off += (long long)K * d.n_det * (2 * m + TAB_W) * TAB_NM;
        k += m;
    }
    bytes = (double)off * sizeof(double2);
    return bytes;
}

static int build_tables(gwin_plan* p) {
    PlanDev& d = p->d;
    CK(cudaDeviceSynchronize());                          // nothing in flight reads the old tables
    free_tables(p);
    p->k_split = d.kmax; p->n_blk_low = p->n_blk; p->k_tab_end = d.kmax;
    p->tab_dirty = false;
    p->tab_worth = false;
    p->cover = p->want;
    if (!p->cover.valid || !(p->cover.a_hi > p->cover.a_lo)) { p->cover.valid = false; return GWIN_OK; }
    static const double budget_gb = getenv("GWIN_TABLE_GB") ? atof(getenv("GWIN_TABLE_GB")) : 2.0;
    const double budget = budget_gb * 1073741824.0;
    // the largest reference-count scale that fits the budget
    std::vector<TabBlk> blks;
    int ib = 0;
    double K0 = 2048.0;
    double bytes = table_schedule(p, p->cover, K0, blks, &ib);
    if (bytes > budget) {
        double lo = 0.5, hi = K0;
        for (int it = 0; it < 10; ++it) {
            const double mid = std::sqrt(lo * hi);
            if (table_schedule(p, p->cover, mid, blks, &ib) > budget) hi = mid; else lo = mid;
        }
        K0 = lo;
        bytes = table_schedule(p, p->cover, K0, blks, &ib);
        if (bytes > budget) return GWIN_OK;               // stay on the recurrence kernel
    }
    const int n_main = (int)blks.size();
    if (n_main == 0) return GWIN_OK;
    {
        // Fine sub-blocks for the tails: every sub-block a covered walker may end in gets its halves,
        // their halves, ... down to >= TAB_MIN bins, appended parent after parent (so that the fine
        // sub-blocks of neighbouring parents are contiguous).
        static const bool fine_on = !(getenv("GWIN_TABLE_FINE") && atoi(getenv("GWIN_TABLE_FINE")) == 0);
        int pair0 = blks[n_main - 1].pair0 + blks[n_main - 1].K;
        long long off = blks[n_main - 1].off + (long long)blks[n_main - 1].K * d.n_det * (2 * blks[n_main - 1].m + TAB_W) * TAB_NM;
        for (int b = 0; fine_on && b < n_main; ++b) {
            const int ps = blks[b].start, pm = blks[b].m;
            if (ps + pm <= p->cover.ke_lo || ps > p->cover.ke_hi || pm / 2 < TAB_MIN) continue;
            if ((double)off * sizeof(double2) > 1.5 * budget) break;
            blks[b].fine_first = (int)blks.size();
            int m = pm / 2;
            for (; m >= TAB_MIN; m /= 2) {
                for (int k = ps; k + m <= ps + pm; k += m) {
                    TabBlk B;
                    memset(&B, 0, sizeof B);
                    B.start = k; B.m = m;
                    B.K = std::max(1, table_refs_needed(p, p->cover, k, m, 4096));
                    B.pair0 = pair0;
                    B.da = (p->cover.a_hi - p->cover.a_lo) / B.K;
                    B.a_lo = p->cover.a_lo;
                    B.inv_da = 1.0 / B.da;
                    B.eps = 0.5 * m / (double)(k + m / 2);
                    B.off = off;
                    B.parent_start = ps;
                    blks.push_back(B);
                    pair0 += B.K;
                    off += (long long)B.K * d.n_det * (2 * m + TAB_W) * TAB_NM;
                }
                blks[b].fine_min = m;
                if (m % 2 || m / 2 < TAB_MIN) break;
            }
            blks[b].fine_n = (int)blks.size() - blks[b].fine_first;
        }
        bytes = (double)off * sizeof(double2);
    }
    const int n = (int)blks.size();
    int n_pairs = 0;
    for (const TabBlk& B : blks) n_pairs += B.K;
    DBG("tables: a0 in [%.6g, %.6g], last bins in [%.0f, %.0f], %d sub-blocks from bin %d (m %d .. %d) + %d fine ones, "
        "%d references, %.1f MB, K0 %.1f", p->cover.a_lo, p->cover.a_hi, p->cover.ke_lo, p->cover.ke_hi, n_main, blks[0].start,
        blks[0].m, blks[n_main - 1].m, n - n_main, n_pairs, bytes / 1048576.0, K0);
    // per-sub-block Taylor matrices, per-pair bookkeeping of the build
    std::vector<double> cmat((size_t)n * NPH * TAB_NTP, 0.0);
    std::vector<int> pair_blk(n_pairs), blk_hat(n);
    std::vector<long long> pair_seg(n_pairs);
    std::vector<double> invhat;
    std::vector<int> hat_of(TAB_MAX + 1, -1);
    long long seg_total = 0;
    int m_max = 0;
    for (int b = 0; b < n; ++b) {
        const TabBlk& B = blks[b];
        long double c[NPH][TAB_NTH];
        taylor_matrix((double)(B.start + B.m / 2), 0.5 * B.m, d.delta_f, c);
        for (int i = 0; i < NPH; ++i)
            for (int q = 0; q < TAB_NT; ++q) cmat[((size_t)b * NPH + i) * TAB_NTP + q] = (double)c[i][q];
        if (hat_of[B.m] < 0) {
            hat_of[B.m] = (int)invhat.size();
            const std::vector<double>& hc = inv_hat(B.m);
            invhat.insert(invhat.end(), hc.begin(), hc.end());
        }
        blk_hat[b] = hat_of[B.m];
        for (int r = 0; r < B.K; ++r) {
            pair_blk[B.pair0 + r] = b;
            pair_seg[B.pair0 + r] = seg_total;
            seg_total += (long long)d.n_det * B.m;
        }
        m_max = std::max(m_max, B.m);
    }
    int *d_pair_blk = nullptr, *d_blk_hat = nullptr;
    long long* d_pair_seg = nullptr;
    double *d_invhat = nullptr, *d_refcoef = nullptr;
    double2* d_seg = nullptr;
    auto build = [&]() -> int {
        CK(cudaMalloc(&p->d_tblk, sizeof(TabBlk) * n));
        CK(cudaMalloc(&p->d_cmat, sizeof(double) * cmat.size()));
        CK(cudaMalloc(&p->d_tref, sizeof(double) * (size_t)n_pairs * TAB_NTP));
        CK(cudaMalloc(&d_pair_blk, sizeof(int) * n_pairs));
        CK(cudaMalloc(&d_blk_hat, sizeof(int) * n));
        CK(cudaMalloc(&d_pair_seg, sizeof(long long) * n_pairs));
        CK(cudaMalloc(&d_invhat, sizeof(double) * invhat.size()));
        CK(cudaMalloc(&d_refcoef, sizeof(double) * (size_t)n_pairs * NPH));
        if (cudaMalloc(&p->d_tab, (size_t)bytes + 1024) != cudaSuccess ||       // (+ the rows a padded window may touch)
            cudaMalloc(&d_seg, sizeof(double2) * (size_t)seg_total) != cudaSuccess) {
            cudaGetLastError();                           // not enough device memory: no tables
            return 1;
        }
        // (those rows meet zero weights in the tensor-core product: they must hold numbers, not whatever the
        //  allocation left there)
        CK(cudaMemset(reinterpret_cast<unsigned char*>(p->d_tab) + (size_t)bytes, 0, 1024));
        CK(cudaMemcpy(p->d_tblk, blks.data(), sizeof(TabBlk) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p->d_cmat, cmat.data(), sizeof(double) * cmat.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_pair_blk, pair_blk.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_blk_hat, blk_hat.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_pair_seg, pair_seg.data(), sizeof(long long) * n_pairs, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_invhat, invhat.data(), sizeof(double) * invhat.size(), cudaMemcpyHostToDevice));
        tab_tref_kernel<<<(n_pairs + 127) / 128, 128>>>(n_pairs, d_pair_blk, p->d_tblk, p->d_cmat, p->d_tref, d_refcoef);
        for (int base = 0; base < n_pairs; base += 32768) {
            const int np = std::min(32768, n_pairs - base);
            tab_seg_kernel<<<dim3((m_max + 255) / 256, np, d.n_det), 256>>>(d, d_pair_blk, p->d_tblk, d_pair_seg, d_blk_hat,
                                                                           d_invhat, p->d_tref, d_refcoef, d_seg, base);
            tab_dft_kernel<<<dim3((2 * m_max + TAB_W + 255) / 256, np, d.n_det), 256>>>(d_pair_blk, p->d_tblk, d_pair_seg,
                                                                                       d_seg, p->d_tab, d.n_det, base);
        }
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        DBG("tables: built");
        return GWIN_OK;
    };
    const int rc = build();
    cudaFree(d_pair_blk); cudaFree(d_blk_hat); cudaFree(d_pair_seg); cudaFree(d_invhat); cudaFree(d_refcoef); cudaFree(d_seg);
    if (rc != GWIN_OK) {
        free_tables(p);
        return rc > 0 ? GWIN_OK : rc;                     // no memory for the tables: the recurrence kernel keeps the band
    }
    p->tblk = blks;
    p->n_tblk = n_main; p->n_fine = n - n_main; p->n_pairs = n_pairs; p->tab_bytes = (size_t)bytes;
    p->tab_split = n_main;
    while (p->tab_split > 0 && blks[p->tab_split - 1].K <= 2) --p->tab_split;
    p->k_split = blks[0].start; p->n_blk_low = ib;
    p->k_tab_end = blks[n_main - 1].start + blks[n_main - 1].m;
    // a sub-block costs about as much as 70 bins of the recurrence kernel, and the table path adds
    // launches: short segments (8 s and below at 4 kHz) are better off without it
    const int64_t tab_bins = (int64_t)p->k_tab_end - p->k_split;
    p->tab_worth = tab_bins >= 16384 && tab_bins >= (int64_t)350 * n_main;
    return GWIN_OK;
}

// ---- coverage ---------------------------------------------------------------------------------
static void cover_merge(Cover& dst, const Cover& src) {
    if (!src.valid) return;
    if (!dst.valid) { dst = src; return; }
    dst.a_lo = std::min(dst.a_lo, src.a_lo);
    dst.a_hi = std::max(dst.a_hi, src.a_hi);
    dst.ke_lo = std::min(dst.ke_lo, src.ke_lo);
    dst.ke_hi = std::max(dst.ke_hi, src.ke_hi);
    for (int i = 0; i < NPH; ++i) dst.dev[i] = std::max(dst.dev[i], src.dev[i]);
}
static bool cover_contains(const Cover& big, const Cover& small) {
    if (!small.valid) return true;
    if (!big.valid) return false;
    if (small.a_lo < big.a_lo || small.a_hi > big.a_hi) return false;
    if (small.ke_lo < big.ke_lo || small.ke_hi > big.ke_hi) return false;
    for (int i = 0; i < NPH; ++i) if (small.dev[i] > big.dev[i]) return false;
    return true;
}
// margins around what a batch was seen to need: the next batches of a sampler wander
static void cover_widen(Cover& c) {
    const double pad = 0.05 * (c.a_hi - c.a_lo) + 2e-3 * c.a_hi;
    c.a_lo = std::max(c.a_lo - pad, 0.25 * c.a_lo);
    c.a_hi += pad;
    c.ke_lo = std::floor(0.97 * c.ke_lo) - 64.0;
    c.ke_hi = std::ceil(1.03 * c.ke_hi) + 64.0;
    for (int i = 0; i < NPH; ++i) c.dev[i] *= 1.5;
}

// Statistics of a batch of DEVICE parameter rows (synchronises `st`)
static int cover_from_device(gwin_plan* p, int64_t W, const double* params, int64_t sw, int64_t si,
                             const uint8_t* skip, cudaStream_t st, Cover* out) {
    out->valid = false;
    if (W <= 0) return GWIN_OK;
    if (!p->d_stats) CK(cudaMalloc(&p->d_stats, sizeof(unsigned long long) * 16));
    unsigned long long init[16];
    memset(init, 0, sizeof init);
    const double inf = INFINITY;
    memcpy(&init[0], &inf, sizeof(double));
    memcpy(&init[14], &inf, sizeof(double));
    CK(cudaMemcpyAsync(p->d_stats, init, sizeof init, cudaMemcpyHostToDevice, st));
    cover_stats_kernel<<<(int)((W + 127) / 128), 128, 0, st>>>(p->d, (int)W, params, sw, si, skip, p->d_stats);
    unsigned long long res[16];
    CK(cudaMemcpyAsync(res, p->d_stats, sizeof res, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (res[2 + NPH] == 0) return GWIN_OK;
    memcpy(&out->a_lo, &res[0], sizeof(double));
    memcpy(&out->a_hi, &res[1], sizeof(double));
    for (int i = 0; i < NPH; ++i) memcpy(&out->dev[i], &res[2 + i], sizeof(double));
    memcpy(&out->ke_lo, &res[14], sizeof(double));
    memcpy(&out->ke_hi, &res[15], sizeof(double));
    out->dev[0] = 0.0;
    out->valid = true;
    return GWIN_OK;
}

// the coverage changed: the recurrence schedule must reach the lightest covered system, the tables
// are rebuilt on the next call that uses them
static void cover_apply(gwin_plan* p) {
    if (!p->want.valid) return;
    if (!cover_contains(p->cover, p->want)) p->tab_dirty = true;
    if (p->auto_mchirp) {
        // a0 = 3 / (256 pi) (pi Mc)^(-5/3)  ->  the smallest chirp mass covered, a quarter octave down
        const double mc = std::pow(p->want.a_hi * (256.0 * G_PI / 3.0), -0.6) / (G_PI * G_MTSUN);
        const double q = std::exp2(std::floor(4.0 * std::log2(std::min(std::max(mc, 0.02), 1e4))) / 4.0);
        if (q < p->mchirp_min || q > 2.0 * p->mchirp_min) { p->mchirp_min = q; p->sched_dirty = true; }
    }
}

extern "C" const char* gwin_last_error(void) { return g_err; }
extern "C" const char* gwin_version(void) { return "gwin_b200 0.1 (sm_100a)"; }

extern "C" int gwin_plan_create(gwin_plan** out, int device, int n_det, int64_t n_f,
                                double delta_f, double epoch_gps,
                                double f_lower, double f_upper, double f_lower_wf,
                                const double* det_response, const double* det_location,
                                const double* data, const double* psd, double norm) {
    if (!out) return fail(GWIN_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_det < 1 || n_det > GWIN_MAX_DET) return fail(GWIN_EINVAL, "n_det must be 1..%d", GWIN_MAX_DET);
    if (n_f < 4 || n_f > (1 << 22)) return fail(GWIN_EINVAL, "n_f out of range");
    if (!(delta_f > 0.0)) return fail(GWIN_EINVAL, "delta_f must be > 0");
    if (!det_response || !det_location || !data) return fail(GWIN_EINVAL, "NULL input array");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(GWIN_ENODEVICE, "no CUDA device: gwin_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GWIN_EINVAL, "device %d not present", device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(GWIN_ENODEVICE, "device %s is sm_%d%d; this library is built for sm_100a only",
                                     prop.name, prop.major, prop.minor);

    // pycbc.filter.get_cutoff_indices(f_lower, f_upper, df, (n_f-1)*2)   gaussian_noise.py:247
    const int64_t N = (n_f - 1) * 2;
    int64_t kmin = f_lower ? (int64_t)(f_lower / delta_f) : 1;
    if (kmin < 0) return fail(GWIN_EINVAL, "low frequency cutoff must be >= 0");
    int64_t kmax = (int64_t)((N + 1) / 2.);
    if (f_upper > 0.0) kmax = std::min<int64_t>((int64_t)(f_upper / delta_f), kmax);
    if (kmax <= kmin) return fail(GWIN_EINVAL, "kmax (%lld) <= kmin (%lld)", (long long)kmax, (long long)kmin);
    if (!(norm > 0.0)) norm = 4.0 * delta_f;

    gwin_plan* p = new (std::nothrow) gwin_plan();
    if (!p) return fail(GWIN_ENOMEM, "out of host memory");
    p->device = device;
    p->norm = norm;
    p->kernel = GWIN_KERNEL_AUTO;
    p->phase_tol = 1e-10;
    p->mchirp_min = 0.5;
    p->auto_mchirp = true;
    p->sched_dirty = true;
    PlanDev& d = p->d;
    d.n_det = n_det; d.n_f = (int)n_f; d.kmin = (int)kmin; d.kmax = (int)kmax;
    d.istart = (int)std::ceil(f_lower_wf / delta_f);
    if (d.istart < 1) d.istart = 1;
    d.delta_f = delta_f; d.epoch = epoch_gps; d.f_lower_wf = f_lower_wf;
    int leaps = 0;
    for (int64_t t : LEAPS_GPS) if ((int64_t)std::floor(epoch_gps) >= t) ++leaps;
    d.gps_utc = leaps;
    for (int i = 0; i < n_det; ++i) {
        memcpy(d.resp[i], det_response + 9 * i, sizeof(double) * 9);
        memcpy(d.loc[i], det_location + 3 * i, sizeof(double) * 3);
    }

    // tables (built once per plan, on the host, in extended precision)
    std::vector<double> basis((size_t)n_f * 4), T((size_t)(n_f + 4) * n_det * 2, 0.0), C((size_t)n_det * (n_f + 1), 0.0);
    std::vector<long double> f76(n_f, 0.0L), f73(n_f, 0.0L);
    basis[0] = 1.0; basis[1] = 1.0; basis[2] = 0.0; basis[3] = 0.0;
    for (int64_t k = 1; k < n_f; ++k) {
        const long double f = (long double)k * (long double)delta_f;
        const long double x = cbrtl(f);
        const long double lx = logl(x);
        basis[4 * k + 0] = (double)x;
        basis[4 * k + 1] = (double)(1.0L / (x * x * x * x * x));
        basis[4 * k + 2] = (double)lx;
        basis[4 * k + 3] = 0.0;
        if (k >= kmin && k < kmax) {
            const long double x2 = x * x, x7 = x2 * x2 * x2 * x;
            f73[k] = 1.0L / x7;                 // f^(-7/3)
            f76[k] = 1.0L / sqrtl(x7);          // f^(-7/6)
        }
    }
    p->w73.assign((size_t)n_det * n_f, 0.0);
    for (int det = 0; det < n_det; ++det) {
        long double csum = 0.0L, nl = 0.0L;
        const double* dd = data + (size_t)det * n_f * 2;
        const double* S = psd ? psd + (size_t)det * n_f : nullptr;
        double* Cd = C.data() + (size_t)det * (n_f + 1);
        double* wd = p->w73.data() + (size_t)det * n_f;
        for (int64_t k = 0; k < n_f; ++k) {
            Cd[k] = (double)csum;
            wd[k] = 0.0;
            if (k >= kmin && k < kmax) {
                const long double w2 = S ? (long double)norm / (long double)S[k] : (long double)norm;
                const long double re = dd[2 * k], im = dd[2 * k + 1];
                nl += (re * re + im * im) * w2;
                csum += f73[k] * w2;
                wd[k] = (double)(f73[k] * w2);
                T[((size_t)k * n_det + det) * 2 + 0] = (double)(re * w2 * f76[k]);
                T[((size_t)k * n_det + det) * 2 + 1] = (double)(-im * w2 * f76[k]);
            }
        }
        Cd[n_f] = (double)csum;
        p->lognl_det[det] = (double)(-0.5L * nl);
    }
    p->lognl = 0.0;
    for (int det = 0; det < n_det; ++det) p->lognl += p->lognl_det[det];

    // device replicas; on any failure the partially built plan is torn down again
    auto upload = [&]() -> int {
        double4* db = nullptr; double2* dT = nullptr; double* dC = nullptr;
        CK(cudaMalloc(&db, basis.size() * sizeof(double)));
        d.basis = db;
        CK(cudaMalloc(&dT, T.size() * sizeof(double)));
        d.T = dT;
        CK(cudaMalloc(&dC, C.size() * sizeof(double)));
        d.C = dC;
        CK(cudaMemcpy(db, basis.data(), basis.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dT, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dC, C.data(), C.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        return GWIN_OK;
    };
    const int rc = upload();
    if (rc != GWIN_OK) { gwin_plan_destroy(p); return rc; }
    *out = p;
    return GWIN_OK;
}

extern "C" void gwin_plan_destroy(gwin_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree((void*)p->d.basis); cudaFree((void*)p->d.T); cudaFree((void*)p->d.C);
    cudaFree(p->d_blk_start); cudaFree(p->d_blk_mat); cudaFree(p->d_mats);
    cudaFree(p->d_wc); cudaFree(p->d_ks); cudaFree(p->d_ke);
    cudaFree(p->d_key_in); cudaFree(p->d_key_out); cudaFree(p->d_perm_in); cudaFree(p->d_perm_out);
    cudaFree(p->d_partial); cudaFree(p->d_cub);
    cudaFree(p->d_marg_u); cudaFree(p->d_marg_logw); cudaFree(p->d_marg_opt);
    cudaFree(p->d_calM); cudaFree(p->d_Ccal); cudaFree(p->d_calib); cudaFreeHost(p->h_calib);
    free_tables(p); cudaFree(p->d_partialT); cudaFree(p->d_partialS);
    cudaFree(p->d_ke_fast); cudaFree(p->d_ke_low); cudaFree(p->d_ks_tail);
    cudaFree(p->d_key2_in); cudaFree(p->d_key2_out); cudaFree(p->d_idx2_in); cudaFree(p->d_perm_e); cudaFree(p->d_grp_b);
    cudaFree(p->d_key3_in); cudaFree(p->d_perm_t); cudaFree(p->d_col_t);
    cudaFree(p->d_ks_fine); cudaFree(p->d_ks_rec); cudaFree(p->d_tail_parent); cudaFree(p->d_grp_p); cudaFree(p->d_partialF);
    cudaFree(p->d_slow); cudaFree(p->d_slow_list); cudaFree(p->d_slow_n); cudaFree(p->d_stats);
    if (p->ev_done_set) cudaEventDestroy(p->ev_done);
    cudaFree(p->d_params); cudaFree(p->d_out); cudaFree(p->d_skip); cudaFree(p->d_points);
    cudaFreeHost(p->h_params); cudaFreeHost(p->h_out); cudaFreeHost(p->h_skip);
    if (p->timing) { cudaEventDestroy(p->ev0); cudaEventDestroy(p->ev1); for (cudaEvent_t e : p->evd) cudaEventDestroy(e); }
    if (p->async_ready) {
        for (int i = 0; i < 2; ++i) {
            gwin_plan::AsyncSlot& a = p->aslot[i];
            cudaFree(a.d_points); cudaFree(a.d_params); cudaFree(a.d_out); cudaFree(a.d_skip);
            cudaFreeHost(a.h_points); cudaFreeHost(a.h_out); cudaFreeHost(a.h_skip);
            cudaEventDestroy(a.ev_in); cudaEventDestroy(a.ev_k); cudaEventDestroy(a.ev_out);
        }
        cudaStreamDestroy(p->stream_in); cudaStreamDestroy(p->stream_out);
    }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" int gwin_plan_cutoffs(const gwin_plan* p, int64_t* kmin, int64_t* kmax) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (kmin) *kmin = p->d.kmin;
    if (kmax) *kmax = p->d.kmax;
    return GWIN_OK;
}

extern "C" int gwin_plan_lognl(const gwin_plan* p, double* total, double* per_det) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (total) *total = p->lognl;
    if (per_det) for (int i = 0; i < p->d.n_det; ++i) per_det[i] = p->lognl_det[i];
    return GWIN_OK;
}

extern "C" int gwin_plan_set_calibration(gwin_plan* p, int n_points, const double* spline_freqs) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    PlanDev& d = p->d;
    cudaFree(p->d_calM); cudaFree(p->d_Ccal);
    p->d_calM = p->d_Ccal = nullptr; d.calM = d.Ccal = nullptr; d.n_cal = 0;
    if (n_points == 0) return GWIN_OK;
    if (n_points < 4 || n_points > GWIN_MAX_CALIB_POINTS)
        return fail(GWIN_EINVAL, "calibration needs 4..%d spline points, got %d", GWIN_MAX_CALIB_POINTS, n_points);
    if (!spline_freqs) return fail(GWIN_EINVAL, "spline_freqs is NULL");
    const int nd = d.n_det, n = n_points;
    const int64_t n_f = d.n_f;
    std::vector<double> M((size_t)nd * 4 * n);
    std::vector<double> C((size_t)nd * 7 * (n_f + 1));
    for (int det = 0; det < nd; ++det) {
        const double* sf = spline_freqs + (size_t)det * n;
        double fmin = sf[0], fmax = sf[0];
        for (int j = 0; j < n; ++j) {
            if (!std::isfinite(sf[j])) return fail(GWIN_EINVAL, "spline frequency %d of detector %d is not finite", j, det);
            fmin = std::min(fmin, sf[j]); fmax = std::max(fmax, sf[j]);
        }
        if (!(fmax > fmin)) return fail(GWIN_EINVAL, "spline frequencies of detector %d span no range", det);
        const double fm = 0.5 * (fmax + fmin), fh = 0.5 * (fmax - fmin);
        d.cal_zs[det] = d.delta_f / fh;
        d.cal_z0[det] = -fm / fh;
        // least-squares cubic: Householder QR of the n x 4 Vandermonde matrix in z, M = R^-1 Q^T
        std::vector<long double> A((size_t)n * 4), Qt((size_t)n * n, 0.0L);
        for (int j = 0; j < n; ++j) {
            const long double z = ((long double)sf[j] - (long double)fm) / (long double)fh;
            A[j * 4 + 0] = 1.0L; A[j * 4 + 1] = z; A[j * 4 + 2] = z * z; A[j * 4 + 3] = z * z * z;
            Qt[(size_t)j * n + j] = 1.0L;
        }
        for (int c = 0; c < 4; ++c) {
            long double nrm = 0.0L;
            for (int j = c; j < n; ++j) nrm += A[j * 4 + c] * A[j * 4 + c];
            nrm = sqrtl(nrm);
            if (!(nrm > 0.0L)) return fail(GWIN_EINVAL, "spline frequencies of detector %d are degenerate", det);
            const long double alpha = A[c * 4 + c] > 0 ? -nrm : nrm;
            std::vector<long double> v(n, 0.0L);
            for (int j = c; j < n; ++j) v[j] = A[j * 4 + c];
            v[c] -= alpha;
            long double vv = 0.0L;
            for (int j = c; j < n; ++j) vv += v[j] * v[j];
            if (!(vv > 0.0L)) continue;
            for (int cc = 0; cc < 4; ++cc) {              // A <- (I - 2 v v^T / v^T v) A
                long double dot = 0.0L;
                for (int j = c; j < n; ++j) dot += v[j] * A[j * 4 + cc];
                dot = 2.0L * dot / vv;
                for (int j = c; j < n; ++j) A[j * 4 + cc] -= dot * v[j];
            }
            for (int cc = 0; cc < n; ++cc) {              // Qt <- H Qt
                long double dot = 0.0L;
                for (int j = c; j < n; ++j) dot += v[j] * Qt[(size_t)j * n + cc];
                dot = 2.0L * dot / vv;
                for (int j = c; j < n; ++j) Qt[(size_t)j * n + cc] -= dot * v[j];
            }
        }
        for (int col = 0; col < n; ++col) {               // back substitution per right-hand side
            long double x[4];
            for (int i = 3; i >= 0; --i) {
                long double acc = Qt[(size_t)i * n + col];
                for (int q = i + 1; q < 4; ++q) acc -= A[i * 4 + q] * x[q];
                if (A[i * 4 + i] == 0.0L) return fail(GWIN_EINVAL, "spline frequencies of detector %d are degenerate", det);
                x[i] = acc / A[i * 4 + i];
            }
            for (int i = 0; i < 4; ++i) M[((size_t)det * 4 + i) * n + col] = (double)x[i];
        }
        // prefix sums of the <h|h> weights times z^j
        const double* wd = p->w73.data() + (size_t)det * n_f;
        long double sum[7] = {0, 0, 0, 0, 0, 0, 0};
        double* Cd = C.data() + (size_t)det * 7 * (n_f + 1);
        for (int64_t k = 0; k <= n_f; ++k) {
            for (int j = 0; j < 7; ++j) Cd[(size_t)j * (n_f + 1) + k] = (double)sum[j];
            if (k == n_f || wd[k] == 0.0) continue;
            const long double z = (long double)k * (long double)d.cal_zs[det] + (long double)d.cal_z0[det];
            long double zj = wd[k];
            for (int j = 0; j < 7; ++j) { sum[j] += zj; zj *= z; }
        }
    }
    CK(cudaMalloc(&p->d_calM, sizeof(double) * M.size()));
    CK(cudaMalloc(&p->d_Ccal, sizeof(double) * C.size()));
    CK(cudaMemcpy(p->d_calM, M.data(), sizeof(double) * M.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_Ccal, C.data(), sizeof(double) * C.size(), cudaMemcpyHostToDevice));
    d.calM = p->d_calM; d.Ccal = p->d_Ccal; d.n_cal = n;
    return GWIN_OK;
}

extern "C" int gwin_plan_set_distance_marg(gwin_plan* p, int64_t n, const double* dist, const double* weight) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (n < 0 || n > (1 << 24)) return fail(GWIN_EINVAL, "distance grid size out of range");
    if (n && (!dist || !weight)) return fail(GWIN_EINVAL, "dist/weight is NULL");
    std::vector<double> u((size_t)n), lw((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        if (!(dist[j] > 0.0) || !std::isfinite(dist[j])) return fail(GWIN_EINVAL, "dist[%lld] must be positive and finite", (long long)j);
        if (!(weight[j] >= 0.0) || !std::isfinite(weight[j])) return fail(GWIN_EINVAL, "weight[%lld] must be finite and >= 0", (long long)j);
        u[j] = 1.0 / dist[j];
        lw[j] = weight[j] > 0.0 ? std::log(weight[j]) : -INFINITY;
    }
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    cudaFree(p->d_marg_u); cudaFree(p->d_marg_logw);
    p->d_marg_u = p->d_marg_logw = nullptr; p->n_marg = 0;
    if (n == 0) return GWIN_OK;
    CK(cudaMalloc(&p->d_marg_u, sizeof(double) * n));
    CK(cudaMalloc(&p->d_marg_logw, sizeof(double) * n));
    CK(cudaMemcpy(p->d_marg_u, u.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_marg_logw, lw.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    p->n_marg = (int)n;
    return GWIN_OK;
}

extern "C" int gwin_plan_configure(gwin_plan* p, int kernel, double phase_tol, double mchirp_min) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (kernel < GWIN_KERNEL_AUTO || kernel > GWIN_KERNEL_TABLE) return fail(GWIN_EINVAL, "bad kernel id");
    p->kernel = kernel;
    if (phase_tol > 0.0 && phase_tol != p->phase_tol) { p->phase_tol = phase_tol; p->sched_dirty = true; }
    if (mchirp_min > 0.0) {
        p->auto_mchirp = false;
        if (mchirp_min != p->mchirp_min) { p->mchirp_min = mchirp_min; p->sched_dirty = true; }
    }
    return GWIN_OK;
}

static int cover_declare(gwin_plan* p, const Cover& seen) {
    if (!seen.valid) return fail(GWIN_EINVAL, "none of the points has a waveform");
    Cover c = seen;
    cover_widen(c);
    cover_merge(p->want, c);
    p->cover_manual = true;
    cover_apply(p);
    return GWIN_OK;
}

extern "C" int gwin_plan_cover_device(gwin_plan* p, int64_t W, const double* params,
                                      int64_t stride_walker, int64_t stride_param, void* stream) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    CK(cudaSetDevice(p->device));
    if (W <= 0 || !params) {                              // back to automatic coverage
        p->want.valid = false; p->cover_manual = false; p->tab_dirty = true;
        return GWIN_OK;
    }
    if (stride_walker < 1 || stride_param < 1) return fail(GWIN_EINVAL, "parameter strides must be >= 1");
    Cover seen;
    memset(&seen, 0, sizeof seen);
    int rc = cover_from_device(p, W, params, stride_walker, stride_param, nullptr, (cudaStream_t)stream, &seen);
    if (rc) return rc;
    return cover_declare(p, seen);
}

extern "C" int gwin_plan_cover_host(gwin_plan* p, int64_t W, const double* params,
                                    int64_t stride_walker, int64_t stride_param) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (W <= 0 || !params) return gwin_plan_cover_device(p, 0, nullptr, 0, 0, nullptr);
    if (!((stride_walker == GWIN_NPARAM && stride_param == 1) || (stride_walker == 1 && stride_param == W)))
        return fail(GWIN_EINVAL, "host parameters must be a dense [W][13] or [13][W] block");
    CK(cudaSetDevice(p->device));
    double* d = nullptr;
    CK(cudaMalloc(&d, sizeof(double) * GWIN_NPARAM * W));
    int rc = GWIN_OK;
    if (cudaMemcpy(d, params, sizeof(double) * GWIN_NPARAM * W, cudaMemcpyHostToDevice) != cudaSuccess)
        rc = fail(GWIN_ECUDA, "copy of the coverage points failed");
    if (!rc) rc = gwin_plan_cover_device(p, W, d, stride_walker, stride_param, p->stream);
    cudaFree(d);
    return rc;
}

extern "C" int gwin_plan_table_info(gwin_plan* p, int64_t* n_subblocks, int64_t* n_references,
                                    int64_t* bytes, int64_t* k_split) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (n_subblocks) *n_subblocks = p->n_tblk;
    if (n_references) *n_references = p->n_pairs;
    if (bytes) *bytes = (int64_t)p->tab_bytes;
    if (k_split) *k_split = p->n_tblk ? p->k_split : p->d.kmax;
    return GWIN_OK;
}

extern "C" int gwin_plan_slow_path(gwin_plan* p, int64_t* n_walkers) {
    if (!p || !n_walkers) return fail(GWIN_EINVAL, "NULL argument");
    *n_walkers = 0;
    if (!p->d_slow_n) return GWIN_OK;
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    int n = 0;
    CK(cudaMemcpy(&n, p->d_slow_n + 1, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemset(p->d_slow_n + 1, 0, sizeof(int)));
    *n_walkers = n;
    return GWIN_OK;
}

extern "C" int gwin_plan_last_launches(const gwin_plan* p) { return p ? p->last_launches : 0; }

static void timing_harvest(gwin_plan* p) {
    if (!p->timing || !p->ev_pending) return;
    float ms = 0.f;
    if (cudaEventSynchronize(p->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, p->ev0, p->ev1) == cudaSuccess) {
        p->kernel_ms_sum += ms; p->kernel_ms_n++;
        cudaEvent_t seq[5] = { p->ev0, p->evd[0], p->evd[1], p->evd[2], p->ev1 };
        for (int i = 0; i < 4; ++i)
            if (cudaEventElapsedTime(&ms, seq[i], seq[i + 1]) == cudaSuccess) p->detail_ms_sum[i] += ms;
    }
    p->ev_pending = false;
}

extern "C" int gwin_plan_kernel_timing(gwin_plan* p, int enable, double* mean_ms, int* n_launches) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    CK(cudaSetDevice(p->device));
    timing_harvest(p);
    if (mean_ms) *mean_ms = p->kernel_ms_n ? p->kernel_ms_sum / p->kernel_ms_n : 0.0;
    if (n_launches) *n_launches = p->kernel_ms_n;
    for (int i = 0; i < 4; ++i) p->detail_last[i] = p->kernel_ms_n ? p->detail_ms_sum[i] / p->kernel_ms_n : 0.0;
    if (enable && !p->timing) {
        CK(cudaEventCreate(&p->ev0)); CK(cudaEventCreate(&p->ev1));
        for (cudaEvent_t& e : p->evd) CK(cudaEventCreate(&e));
        p->timing = true;
    } else if (!enable && p->timing) {
        cudaEventDestroy(p->ev0); cudaEventDestroy(p->ev1);
        for (cudaEvent_t e : p->evd) cudaEventDestroy(e);
        p->timing = false;
    }
    p->kernel_ms_sum = 0.0; p->kernel_ms_n = 0;
    for (double& x : p->detail_ms_sum) x = 0.0;
    return GWIN_OK;
}

extern "C" int gwin_plan_kernel_timing_detail(const gwin_plan* p, double* mean_ms4) {
    if (!p || !mean_ms4) return fail(GWIN_EINVAL, "NULL argument");
    for (int i = 0; i < 4; ++i) mean_ms4[i] = p->detail_last[i];
    return GWIN_OK;
}

extern "C" int gwin_plan_table_blocks(const gwin_plan* p, int64_t cap, int32_t* start, int32_t* m, int32_t* n_refs,
                                      int32_t* fine_min, int64_t* n) {
    if (!p || !n) return fail(GWIN_EINVAL, "NULL argument");
    *n = p->n_tblk;
    for (int64_t i = 0; i < std::min<int64_t>(cap, p->n_tblk); ++i) {
        const TabBlk& B = p->tblk[i];
        if (start) start[i] = B.start;
        if (m) m[i] = B.m;
        if (n_refs) n_refs[i] = B.K;
        if (fine_min) fine_min[i] = B.fine_n > 0 ? B.fine_min : 0;
    }
    return GWIN_OK;
}

static int ensure_workspace(gwin_plan* p, int64_t W, int n_chunks) {
    const int nd = p->d.n_det;
    if (W > p->w_cap) {
        int64_t cap = std::max<int64_t>(W, 1024);
        cap = (cap + CTA_THREADS - 1) / CTA_THREADS * CTA_THREADS;
        CK(cudaDeviceSynchronize());
        int** ints[] = { &p->d_ks, &p->d_ke, &p->d_key_in, &p->d_key_out, &p->d_perm_in, &p->d_perm_out,
                         &p->d_ke_fast, &p->d_ke_low, &p->d_ks_tail, &p->d_key2_in, &p->d_key2_out,
                         &p->d_idx2_in, &p->d_perm_e, &p->d_slow_list, &p->d_key3_in, &p->d_perm_t, &p->d_col_t,
                         &p->d_ks_fine, &p->d_ks_rec, &p->d_tail_parent };
        for (int** q : ints) { cudaFree(*q); *q = nullptr; }
        cudaFree(p->d_wc); cudaFree(p->d_cub); cudaFree(p->d_marg_opt); cudaFree(p->d_slow); cudaFree(p->d_grp_b); cudaFree(p->d_grp_p);
        p->d_wc = nullptr; p->d_cub = nullptr; p->d_marg_opt = nullptr; p->d_slow = nullptr; p->d_grp_b = p->d_grp_p = nullptr;
        p->w_cap = 0;
        // coefficient rows + n_det rows of hh
        CK(cudaMalloc(&p->d_wc, sizeof(double) * (NCOEF(nd) + nd) * cap));
        for (int** q : ints) CK(cudaMalloc(q, sizeof(int) * cap));
        CK(cudaMalloc(&p->d_marg_opt, sizeof(double) * cap));
        CK(cudaMalloc(&p->d_slow, cap));
        CK(cudaMalloc(&p->d_grp_b, sizeof(int) * 2 * (cap / CTA_THREADS + 1)));
        CK(cudaMalloc(&p->d_grp_p, sizeof(int) * 2 * (cap / CTA_THREADS + 1)));
        if (!p->d_slow_n) {
            CK(cudaMalloc(&p->d_slow_n, sizeof(int) * 2));
            CK(cudaMemset(p->d_slow_n, 0, sizeof(int) * 2));
        }
        p->partialT_cap = 0;
        p->partialS_cap = 0;
        p->partialF_cap = 0;
        p->cub_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, p->cub_bytes, p->d_key_in, p->d_key_out,
                                           p->d_perm_in, p->d_perm_out, (int)cap, 0, 23));
        CK(cudaMalloc(&p->d_cub, p->cub_bytes));
        p->w_cap = cap;
        p->partial_cap = 0;
    }
    const size_t need = (size_t)n_chunks * nd * p->w_cap;
    if (need > p->partial_cap) {
        CK(cudaDeviceSynchronize());
        cudaFree(p->d_partial); p->d_partial = nullptr; p->partial_cap = 0;
        CK(cudaMalloc(&p->d_partial, sizeof(double2) * need));
        p->partial_cap = need;
    }
    return GWIN_OK;
}

// One launch of the recurrence kernel (or, recur = false, of the direct kernel).
//   blk_off, n_blk   the part of the block schedule the launch covers
//   ks_arr, ke_arr   per-walker band
//   col, grp_b       tails of the table path (see loglr_recur_kernel), grid.y = n_chunks slots
//   list, n_list     slow path of the direct kernel (see loglr_direct_kernel), grid.x = grid_x CTAs
struct MainLaunch {
    bool recur, calibrated;
    int W, wpad, n_chunks, per_chunk, n_blk, blk_off;
    const int *ks_arr, *ke_arr;
    double2* partial;
    const int *col, *grp_b;
    const int *list, *n_list;
    int grid_x;
};
template <int ND>
static void launch_main(gwin_plan* p, const MainLaunch& a, cudaStream_t st) {
    dim3 grid(a.grid_x > 0 ? a.grid_x : (a.W + CTA_THREADS - 1) / CTA_THREADS, a.n_chunks);
    // diagnostic: GWIN_SMEM_PAD=<bytes> of unused dynamic shared memory lowers the occupancy
    static const int smem_pad = getenv("GWIN_SMEM_PAD") ? atoi(getenv("GWIN_SMEM_PAD")) : 0;
    constexpr int NWARP = CTA_THREADS / WARP;
    const int smem = (int)(sizeof(double2) * NWARP * 2 * TileStream<ND>::TILE * ND
                           + sizeof(double) * (a.calibrated ? RC_ROWS_CAL(ND) : RC_ROWS(ND)) * CTA_THREADS
                           + sizeof(uint64_t) * NWARP * 2) + smem_pad;
    // function attributes belong to the device's context: once per (ND, CALIB) and device
    static bool attr_set_dev[64][2] = {};
    bool& attr_flag = attr_set_dev[p->device & 63][a.calibrated];
    if (a.recur && !attr_flag) {
        if (a.calibrated)
            cudaFuncSetAttribute(loglr_recur_kernel<ND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        else
            cudaFuncSetAttribute(loglr_recur_kernel<ND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_flag = true;
    }
    if (a.recur && a.calibrated)
        loglr_recur_kernel<ND, true><<<grid, CTA_THREADS, smem, st>>>(p->d, a.W, a.wpad, a.per_chunk, a.n_blk, p->d_blk_start + a.blk_off,
                                                                  p->d_blk_mat + a.blk_off, p->d_mats,
                                                                  p->d_wc, a.ks_arr, a.ke_arr, a.partial, a.col, a.grp_b);
    else if (a.recur)
        loglr_recur_kernel<ND, false><<<grid, CTA_THREADS, smem, st>>>(p->d, a.W, a.wpad, a.per_chunk, a.n_blk, p->d_blk_start + a.blk_off,
                                                                   p->d_blk_mat + a.blk_off, p->d_mats,
                                                                   p->d_wc, a.ks_arr, a.ke_arr, a.partial, a.col, a.grp_b);
    else if (a.calibrated)
        loglr_direct_kernel<ND, true><<<grid, CTA_THREADS, 0, st>>>(p->d, a.W, a.wpad, a.per_chunk, a.n_chunks,
                                                                   p->d_wc, a.ks_arr, a.ke_arr, a.partial, a.list, a.n_list);
    else
        loglr_direct_kernel<ND, false><<<grid, CTA_THREADS, 0, st>>>(p->d, a.W, a.wpad, a.per_chunk, a.n_chunks,
                                                                    p->d_wc, a.ks_arr, a.ke_arr, a.partial, a.list, a.n_list);
}
static void launch_main_nd(gwin_plan* p, const MainLaunch& a, cudaStream_t st) {
    switch (p->d.n_det) {
        case 1: launch_main<1>(p, a, st); break;
        case 2: launch_main<2>(p, a, st); break;
        case 3: launch_main<3>(p, a, st); break;
        default: launch_main<4>(p, a, st); break;
    }
}

// fine = false: sub-blocks [blk_lo, blk_hi) of the schedule, per_chunk of them per CTA.
// fine = true: the tails' fine sub-blocks, walkers in the order of their last bins, `slots` CTAs per group.
template <int ND, bool CALIB>
static void launch_table(gwin_plan* p, int W, int wpad, int per_chunk, int blk_lo, int blk_hi, const int* col,
                         bool fine, int slots, cudaStream_t st) {
    if (!fine && blk_hi <= blk_lo) return;
    dim3 grid((W + CTA_THREADS - 1) / CTA_THREADS, fine ? slots : (blk_hi - blk_lo + per_chunk - 1) / per_chunk);
    const int fixed = (TAB_CROWS(ND) + TAB_WPITCH) * CTA_THREADS * (int)sizeof(double);
    const int smem = (fine ? 0 : per_chunk) * NPH * TAB_NTP * (int)sizeof(double) + fixed;
    static bool attr_set_dev[64] = {};
    if (!attr_set_dev[p->device & 63]) {
        cudaFuncSetAttribute(loglr_table_kernel<ND, false, CALIB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TAB_CHUNK_MAX * NPH * TAB_NTP * (int)sizeof(double) + fixed);
        cudaFuncSetAttribute(loglr_table_kernel<ND, false, CALIB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(loglr_table_kernel<ND, true, CALIB>, cudaFuncAttributeMaxDynamicSharedMemorySize, fixed);
        cudaFuncSetAttribute(loglr_table_kernel<ND, true, CALIB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_set_dev[p->device & 63] = true;
    }
    if (fine)
        loglr_table_kernel<ND, true, CALIB><<<grid, CTA_THREADS, smem, st>>>(p->d, W, wpad, 0, 0, 0, p->d_tblk, p->d_cmat,
                                                                         p->d_tref, p->d_tab, p->d_wc, p->d_ke_fast,
                                                                         p->d_partialF, p->d_slow, col,
                                                                         p->d_ks_tail, p->d_ks_fine, p->d_tail_parent);
    else
        loglr_table_kernel<ND, false, CALIB><<<grid, CTA_THREADS, smem, st>>>(p->d, W, wpad, per_chunk, blk_lo, blk_hi, p->d_tblk, p->d_cmat,
                                                                          p->d_tref, p->d_tab, p->d_wc, p->d_ke_fast,
                                                                          p->d_partialT, p->d_slow, col, nullptr, nullptr, nullptr);
}
static void launch_table_nd(gwin_plan* p, int W, int wpad, int per_chunk, int blk_lo, int blk_hi, const int* col,
                            bool fine, int slots, bool calib, cudaStream_t st) {
#define GWIN_LT(ND) do { if (calib) launch_table<ND, true>(p, W, wpad, per_chunk, blk_lo, blk_hi, col, fine, slots, st); \
                         else launch_table<ND, false>(p, W, wpad, per_chunk, blk_lo, blk_hi, col, fine, slots, st); } while (0)
    switch (p->d.n_det) {
        case 1: GWIN_LT(1); break;
        case 2: GWIN_LT(2); break;
        case 3: GWIN_LT(3); break;
        default: GWIN_LT(4); break;
    }
#undef GWIN_LT
}

extern "C" int gwin_loglr_batch(gwin_plan* p, int mode, int64_t W64,
                                const double* params, const uint8_t* skip,
                                double* loglr, double* hd, double* hh, void* stream) {
    return gwin_loglr_batch_strided(p, mode, W64, params, GWIN_NPARAM, 1, skip, loglr, hd, hh, stream);
}

static int loglr_batch_impl(gwin_plan* p, int mode, int64_t W64,
                            const double* params, int64_t stride_walker, int64_t stride_param,
                            const double* calib, int64_t cstride_walker, int64_t cstride_value,
                            const uint8_t* skip,
                            double* loglr, double* hd, double* hh, void* stream);

extern "C" int gwin_loglr_batch_strided(gwin_plan* p, int mode, int64_t W64,
                                        const double* params, int64_t stride_walker, int64_t stride_param,
                                        const uint8_t* skip,
                                        double* loglr, double* hd, double* hh, void* stream) {
    return loglr_batch_impl(p, mode, W64, params, stride_walker, stride_param, nullptr, 0, 0, skip, loglr, hd, hh, stream);
}

extern "C" int gwin_loglr_batch_calib(gwin_plan* p, int mode, int64_t W64,
                                      const double* params, int64_t stride_walker, int64_t stride_param,
                                      const double* calib, int64_t cstride_walker, int64_t cstride_value,
                                      const uint8_t* skip,
                                      double* loglr, double* hd, double* hh, void* stream) {
    if (p && calib && p->d.n_cal == 0)
        return fail(GWIN_EINVAL, "calibration values given but gwin_plan_set_calibration was not called");
    if (calib && (cstride_walker < 1 || cstride_value < 1)) return fail(GWIN_EINVAL, "calibration strides must be >= 1");
    return loglr_batch_impl(p, mode, W64, params, stride_walker, stride_param, calib, cstride_walker, cstride_value,
                            skip, loglr, hd, hh, stream);
}

// Would this batch go through the moment tables (once they cover it)?
static bool table_wanted(const gwin_plan* p, int64_t W, bool calibrated) {
    static const bool table_default = !(getenv("GWIN_TABLE") && atoi(getenv("GWIN_TABLE")) == 0);
    static const bool table_calib = !(getenv("GWIN_TABLE_CALIB") && atoi(getenv("GWIN_TABLE_CALIB")) == 0);
    if (calibrated && !table_calib) return false;
    if (p->kernel == GWIN_KERNEL_TABLE) return true;
    if (p->kernel != GWIN_KERNEL_AUTO || !table_default) return false;
    // AUTO: tables that exist are used where they pay; they are built for batches worth the build.
    // A segment too short for the tables to pay at all is recognised from the band alone.
    const PlanDev& d = p->d;
    const int band0 = std::max(d.kmin, d.istart);
    if (d.kmax - band0 < 32768) return false;
    if (p->n_tblk > 0 && !p->tab_dirty) return p->tab_worth;
    return W >= 256;
}

static int loglr_batch_impl(gwin_plan* p, int mode, int64_t W64,
                            const double* params, int64_t stride_walker, int64_t stride_param,
                            const double* calib, int64_t cstride_walker, int64_t cstride_value,
                            const uint8_t* skip,
                            double* loglr, double* hd, double* hh, void* stream) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (stride_walker < 1 || stride_param < 1) return fail(GWIN_EINVAL, "parameter strides must be >= 1");
    if (mode < GWIN_MODE_GAUSSIAN || mode > GWIN_MODE_MARG_DIST_PHASE) return fail(GWIN_EINVAL, "bad mode %d", mode);
    if (mode >= GWIN_MODE_MARG_DIST && p && p->n_marg == 0)
        return fail(GWIN_EINVAL, "distance marginalisation needs gwin_plan_set_distance_marg first");
    if (W64 < 0 || W64 > (1 << 26)) return fail(GWIN_EINVAL, "W out of range");
    p->last_launches = 0;
    if (W64 == 0) return GWIN_OK;
    if (!params || !loglr) return fail(GWIN_EINVAL, "params/loglr is NULL");
    CK(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int W = (int)W64;
    const PlanDev& d = p->d;
    const int nd = d.n_det;
    const bool recur = p->kernel != GWIN_KERNEL_DIRECT;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    // Calls on one plan share its workspace: a call on another stream than the previous one waits
    // for it on the device (nothing to order inside a graph capture: the graph's own edges do)
    if (!capturing && p->ev_done_set && p->last_stream != st) CK(cudaStreamWaitEvent(st, p->ev_done, 0));

    // ---- moment tables above k_split (uncalibrated batches); the recurrence kernel keeps [band0, k_split)
    bool table = recur && table_wanted(p, W, calib != nullptr);
    if (table && (!p->want.valid || p->replan) && !p->cover_manual) {
        // what the tables have to cover is read off this batch (first use, or after a batch that sent
        // many walkers down the slow path): one small kernel, one synchronisation
        if (capturing) {
            if (!p->want.valid) return fail(GWIN_EINVAL, "the moment tables' coverage is not known yet and cannot be measured inside a "
                                                         "stream capture: call gwin_plan_cover_device first");
        } else {
            Cover seen;
            memset(&seen, 0, sizeof seen);
            int rc = cover_from_device(p, W, params, stride_walker, stride_param, skip, st, &seen);
            if (rc) return rc;
            if (seen.valid) { cover_widen(seen); cover_merge(p->want, seen); cover_apply(p); }
            p->replan = false;
        }
    }
    if (recur && p->sched_dirty) {
        if (capturing) return fail(GWIN_EINVAL, "the block schedule has to be rebuilt: not inside a stream capture");
        int rc = upload_schedule(p); if (rc) return rc;
    }
    if (table) {
        if (p->tab_dirty) {
            if (capturing) return fail(GWIN_EINVAL, "the moment tables have to be rebuilt: not inside a stream capture");
            int rc = build_tables(p); if (rc) return rc;
        }
        table = p->n_tblk > 0 && (p->kernel == GWIN_KERNEL_TABLE || p->tab_worth);
    }
    const int n_blk = table ? p->n_blk_low : p->n_blk;
    DBG("batch W=%d kernel=%d table=%d n_blk=%d/%d n_tblk=%d k_split=%d mchirp_min=%g", W, p->kernel, (int)table, n_blk, p->n_blk,
        p->n_tblk, p->k_split, p->mchirp_min);

    // frequency chunking: enough CTAs to fill 148 SMs x ~5 resident CTAs several times over
    const int band0 = std::max(d.kmin, d.istart);
    const int nbins = std::max(d.kmax - band0, 1);
    const int groups = (W + CTA_THREADS - 1) / CTA_THREADS;
    static const int chunk_mult = getenv("GWIN_CHUNK_MULT") ? atoi(getenv("GWIN_CHUNK_MULT")) : 192;
    // at most 512 chunks: every chunk costs one partial per walker and detector in the combine pass
    int want_chunks = std::min(512, std::max(1, (148 * chunk_mult + groups - 1) / groups));
    if (table) want_chunks = std::min(want_chunks, std::max(1, (148 * 16 + groups - 1) / groups));     // the short low band
    int n_chunks, per_chunk;
    if (recur) {
        const int min_blocks = 2;
        per_chunk = std::max(min_blocks, (n_blk + want_chunks - 1) / want_chunks);
        n_chunks = (n_blk + per_chunk - 1) / per_chunk;
    } else {
        per_chunk = std::max(RESYNC, (nbins + want_chunks - 1) / want_chunks);
        per_chunk = (per_chunk + RESYNC - 1) / RESYNC * RESYNC;
        n_chunks = (nbins + per_chunk - 1) / per_chunk;
    }
    // tails of the table path: second pass of the recurrence kernel, walkers in the order of their last
    // bin, every CTA over its own walkers' blocks in chunks of tail_blocks, TAIL_SLOTS partial sums
    static const int tail_blocks = std::max(1, getenv("GWIN_TAIL_BLOCKS") ? atoi(getenv("GWIN_TAIL_BLOCKS")) : 2);
    static const int tail_slots = std::max(1, getenv("GWIN_TAIL_SLOTS") ? atoi(getenv("GWIN_TAIL_SLOTS")) : 8);
    const int n_chunks3 = table ? tail_slots : 0;
    int rc = ensure_workspace(p, W, std::max(n_chunks + n_chunks3, 1));
    if (rc) return rc;
    const int wpad = (int)p->w_cap;
    double* hh_rows = p->d_wc + (size_t)NCOEF(nd) * wpad;
    // table chunks: a sub-block costs about as much as 70 recurrence bins
    int n_chunksT = 0, per_chunkT = 1;
    if (table) {
        static const int tab_ctas = getenv("GWIN_TAB_CTAS") ? atoi(getenv("GWIN_TAB_CTAS")) : 24;      // CTAs per SM aimed at
        const int wantT = std::max(1, std::min(p->n_tblk, (148 * tab_ctas + groups - 1) / groups));
        per_chunkT = std::min(TAB_CHUNK_MAX, (p->n_tblk + wantT - 1) / wantT);
        n_chunksT = (p->n_tblk + per_chunkT - 1) / per_chunkT;
        const size_t need = (size_t)n_chunksT * nd * p->w_cap;
        if (need > p->partialT_cap) {
            CK(cudaDeviceSynchronize());
            cudaFree(p->d_partialT); p->d_partialT = nullptr; p->partialT_cap = 0;
            CK(cudaMalloc(&p->d_partialT, sizeof(double2) * need));
            p->partialT_cap = need;
        }
    }
    static const int fine_slots = std::max(1, getenv("GWIN_FINE_SLOTS") ? atoi(getenv("GWIN_FINE_SLOTS")) : 7);
    const int n_chunksF = (table && p->n_fine > 0) ? fine_slots : 0;
    if (n_chunksF > 0) {
        const size_t need = (size_t)n_chunksF * nd * p->w_cap;
        if (need > p->partialF_cap) {
            CK(cudaDeviceSynchronize());
            cudaFree(p->d_partialF); p->d_partialF = nullptr; p->partialF_cap = 0;
            CK(cudaMalloc(&p->d_partialF, sizeof(double2) * need));
            p->partialF_cap = need;
        }
    }
    // slow path (direct kernel over a list of walkers): its own chunking and partial sums
    const int slow_gx = std::min(groups, 8);
    int per_chunkS = std::max(RESYNC, (nbins + 255) / 256);
    per_chunkS = (per_chunkS + RESYNC - 1) / RESYNC * RESYNC;
    const int n_chunksS = (nbins + per_chunkS - 1) / per_chunkS;
    if (recur) {
        const size_t need = (size_t)n_chunksS * nd * p->w_cap;
        if (need > p->partialS_cap) {
            CK(cudaDeviceSynchronize());
            cudaFree(p->d_partialS); p->d_partialS = nullptr; p->partialS_cap = 0;
            CK(cudaMalloc(&p->d_partialS, sizeof(double2) * need));
            p->partialS_cap = need;
        }
    }

    const int tb = 128, gb = (W + tb - 1) / tb;
    const int* perm = nullptr;
    bool time_order = false;
    if (W > WARP || table) {
        const double a_lo = p->cover.a_lo;
        const double a_scale = table ? (double)(1 << TAB_KEY_BITS) / (p->cover.a_hi - p->cover.a_lo) : 0.0;
        // the walkers' last bins lie in the covered range (table path) or anywhere in the band
        const double e_lo = table ? std::max(p->cover.ke_lo, (double)band0) : (double)band0;
        const double e_hi = table ? std::min(p->cover.ke_hi, (double)d.kmax) : (double)d.kmax;
        // the long sub-blocks with one or two reference chirps are taken with the walkers in time order
        // (off by default: with the window walk of the table kernel the lanes' row spread no longer costs
        //  shared-memory wavefronts, and in time order the lanes of a warp end at different sub-blocks)
        static const bool time_order_on = getenv("GWIN_TIME_ORDER") && atoi(getenv("GWIN_TIME_ORDER")) != 0;
        time_order = table && time_order_on && p->tab_split < p->n_tblk;
        const double f_split = time_order ? p->tblk[p->tab_split].start * d.delta_f : 1.0;
        sort_key_kernel<<<gb, tb, 0, st>>>(d, W, params, stride_walker, stride_param, skip, p->d_key_in, p->d_perm_in,
                                           table, a_lo, a_scale, table ? p->d_key2_in : nullptr,
                                           time_order ? p->d_key3_in : nullptr, 5.0 / 3.0 * std::pow(f_split, -8.0 / 3.0),
                                           e_lo, 4095.0 / std::max(e_hi - e_lo, 1.0));
        const int bits = table ? TAB_KEY_BITS : 12;
        static const bool block_sort = !(getenv("GWIN_BLOCK_SORT") && atoi(getenv("GWIN_BLOCK_SORT")) == 0);
        if (W <= 16384 && block_sort) {
            // one launch: CTA 0 sorts the processing order, CTAs 1, 2 (table path) the tails' and the time order
            const int items = W <= 4096 ? 4 : 16;
            const size_t smem = items == 4 ? sizeof(cub::BlockRadixSort<int, 1024, 4, int>::TempStorage)
                                           : sizeof(cub::BlockRadixSort<int, 1024, 16, int>::TempStorage);
            static bool attr_set_dev[64] = {};
            if (!attr_set_dev[p->device & 63]) {
                cudaFuncSetAttribute(block_sort_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(cub::BlockRadixSort<int, 1024, 4, int>::TempStorage));
                cudaFuncSetAttribute(block_sort_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(cub::BlockRadixSort<int, 1024, 16, int>::TempStorage));
                attr_set_dev[p->device & 63] = true;
            }
            SortJobs J;
            J.j[0] = SortJob{p->d_key_in, bits, p->d_perm_out};
            J.j[1] = SortJob{p->d_key2_in, 12, p->d_perm_e};
            J.j[2] = SortJob{p->d_key3_in, 22, p->d_perm_t};
            const int njobs = table ? (time_order ? 3 : 2) : 1;
            static const bool exact_sort = getenv("GWIN_EXACT_SORT") && atoi(getenv("GWIN_EXACT_SORT")) != 0;
            if (!exact_sort) {
                if (items == 4) bucket_order_kernel<4><<<njobs, 1024, 0, st>>>(W, J);
                else bucket_order_kernel<16><<<njobs, 1024, 0, st>>>(W, J);
            } else if (items == 4) block_sort_kernel<4><<<njobs, 1024, smem, st>>>(W, J);
            else block_sort_kernel<16><<<njobs, 1024, smem, st>>>(W, J);
        } else {
            size_t bytes = p->cub_bytes;
            CK(cub::DeviceRadixSort::SortPairs(p->d_cub, bytes, p->d_key_in, p->d_key_out,
                                               p->d_perm_in, p->d_perm_out, W, 0, bits, st));
            if (table) {
                bytes = p->cub_bytes;
                CK(cub::DeviceRadixSort::SortPairs(p->d_cub, bytes, p->d_key2_in, p->d_key2_out,
                                                   p->d_perm_in, p->d_perm_e, W, 0, 12, st));
            }
            if (time_order) {
                bytes = p->cub_bytes;
                CK(cub::DeviceRadixSort::SortPairs(p->d_cub, bytes, p->d_key3_in, p->d_key2_out,
                                                   p->d_perm_in, p->d_perm_t, W, 0, 22, st));
            }
        }
        perm = p->d_perm_out;
        p->last_launches += 2;   // (CUB's device-wide radix-sort passes are library code, not counted)
    }
    FastArgs F;
    memset(&F, 0, sizeof F);
    // the recurrence schedule holds for chirp masses >= mchirp_min, i.e. a0 <= a0(mchirp_min)
    F.a_sched_max = recur ? 3.0 / (256.0 * G_PI) * std::pow(G_PI * p->mchirp_min * G_MTSUN, -5.0 / 3.0) * (1.0 + 1e-12) : INFINITY;
    F.table = table;
    F.cov_lo = p->cover.a_lo; F.cov_hi = p->cover.a_hi;
    F.k_split = p->k_split; F.k_tab_end = p->k_tab_end; F.n_tblk = p->n_tblk; F.blks = p->d_tblk;
    F.ke_fast = p->d_ke_fast; F.ke_low = p->d_ke_low; F.ks_tail = p->d_ks_tail;
    F.inv = p->d_idx2_in; F.slow = p->d_slow;
    F.ks_fine = p->d_ks_fine; F.ks_rec = p->d_ks_rec; F.tail_parent = p->d_tail_parent;
    if (recur) CK(cudaMemsetAsync(p->d_slow_n, 0, sizeof(int), st));
    walker_setup_kernel<<<gb, tb, 0, st>>>(d, W, wpad, params, stride_walker, stride_param, skip, perm, p->d_wc, p->d_ks, p->d_ke, hh_rows,
                                           calib, cstride_walker, cstride_value, F);
    const int* ke_main = table ? p->d_ke_low : p->d_ke_fast;
    double2* partial3 = p->d_partial + (size_t)n_chunks * nd * wpad;
    if (table) {
        // the tails are taken in the order of the walkers' last bins
        tail_plan_kernel<<<groups, CTA_THREADS, 0, st>>>(W, p->n_blk, p->d_blk_start, p->d_perm_e, p->d_idx2_in, p->d_key2_out,
                                                        p->d_ks_rec, p->d_ke_fast, p->d_grp_b,
                                                        time_order ? p->d_perm_t : nullptr, p->d_col_t,
                                                        p->d_tail_parent, p->d_grp_p);
        p->last_launches += 1;
    }
    if (p->timing) {
        timing_harvest(p);              // the previous bracket (it has long finished)
        CK(cudaEventRecord(p->ev0, st));
    }
    MainLaunch M;
    memset(&M, 0, sizeof M);
    M.recur = recur; M.calibrated = calib != nullptr; M.W = W; M.wpad = wpad;
    if (n_chunks > 0 && (!recur || n_blk > 0)) {
        M.n_chunks = n_chunks; M.per_chunk = per_chunk; M.n_blk = n_blk; M.blk_off = 0;
        M.ks_arr = p->d_ks; M.ke_arr = ke_main; M.partial = p->d_partial;
        launch_main_nd(p, M, st);
        p->last_launches += 1;
    }
    if (p->timing) CK(cudaEventRecord(p->evd[0], st));
    if (table) {
        // sub-blocks with several reference chirps: walkers in a0 order; the long ones above: in time order
        int split = time_order ? std::min(p->n_tblk, (p->tab_split + per_chunkT - 1) / per_chunkT * per_chunkT) : p->n_tblk;
        launch_table_nd(p, W, wpad, per_chunkT, 0, split, nullptr, false, 0, calib != nullptr, st);
        launch_table_nd(p, W, wpad, per_chunkT, split, p->n_tblk, p->d_col_t, false, 0, calib != nullptr, st);
        p->last_launches += split > 0 && split < p->n_tblk ? 1 : 0;
        if (p->timing) CK(cudaEventRecord(p->evd[1], st));
        // tails: fine sub-blocks (walkers in the order of their last bins), then the recurrence kernel
        // over what is left (less than the finest sub-block)
        if (n_chunksF > 0) {
            launch_table_nd(p, W, wpad, 0, 0, 0, p->d_key2_out, true, n_chunksF, calib != nullptr, st);
            p->last_launches += 1;
        }
        if (p->timing) CK(cudaEventRecord(p->evd[2], st));
        MainLaunch T = M;
        T.recur = true; T.calibrated = calib != nullptr;
        T.n_chunks = n_chunks3; T.per_chunk = tail_blocks; T.n_blk = p->n_blk; T.blk_off = 0;
        T.ks_arr = p->d_ks_rec; T.ke_arr = p->d_ke_fast; T.partial = partial3;
        T.col = p->d_key2_out; T.grp_b = p->d_grp_b;
        launch_main_nd(p, T, st);
        p->last_launches += 2;
    }
    if (p->timing) {
        if (!table) { CK(cudaEventRecord(p->evd[1], st)); CK(cudaEventRecord(p->evd[2], st)); }
        CK(cudaEventRecord(p->ev1, st));
        p->ev_pending = true;
    }
    combine_kernel<<<(W + 31) / 32, dim3(32, COMB_SEG), 0, st>>>(nd, W, wpad, (n_blk > 0 || !recur) ? n_chunks : 0, mode, per_chunk, n_blk,
                                                                 recur ? p->d_blk_start : nullptr, band0, d.kmax,
                                                                 perm, p->d_wc, p->d_ks, ke_main,
                                                                 p->d_partial, hh_rows, loglr, hd, hh, p->d_marg_opt,
                                                                 n_chunksT, per_chunkT, p->d_tblk, p->d_partialT, p->d_ke_fast,
                                                                 n_chunks3, partial3, n_chunksF, p->d_partialF);
    p->last_launches += 2;
    if (recur) {
        // slow path: whoever the fast kernels could not take goes through the direct kernel (no one, normally:
        // the CTAs of the three launches then exit at once)
        slow_compact_kernel<<<gb, tb, 0, st>>>(W, p->d_slow, p->d_slow_list, p->d_slow_n);
        MainLaunch S = M;
        S.recur = false;
        S.n_chunks = n_chunksS; S.per_chunk = per_chunkS; S.ks_arr = p->d_ks; S.ke_arr = p->d_ke; S.partial = p->d_partialS;
        S.list = p->d_slow_list; S.n_list = p->d_slow_n; S.grid_x = slow_gx;
        launch_main_nd(p, S, st);
        slow_finish_kernel<<<slow_gx, CTA_THREADS, 0, st>>>(nd, wpad, n_chunksS, mode, per_chunkS, band0, d.kmax, perm, p->d_wc,
                                                            p->d_ks, p->d_ke, p->d_partialS, hh_rows, p->d_slow_list,
                                                            p->d_slow_n, loglr, hd, hh, p->d_marg_opt);
        p->last_launches += 3;
    }
    if (mode >= GWIN_MODE_MARG_DIST) {
        marg_dist_kernel<<<(W + 7) / 8, 256, 0, st>>>(W, p->n_marg, p->d_marg_u, p->d_marg_logw, p->d_marg_opt, loglr);
        p->last_launches += 1;
    }
    CK(cudaGetLastError());
    if (!capturing) {
        if (!p->ev_done_set) { CK(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming)); p->ev_done_set = true; }
        CK(cudaEventRecord(p->ev_done, st));
        p->last_stream = st;
    }
    return GWIN_OK;
}

static int ensure_stage(gwin_plan* p, int64_t W) {
    if (W <= p->stage_cap) return GWIN_OK;
    const int nd = p->d.n_det;
    int64_t cap = std::max<int64_t>(W, 1024);
    CK(cudaDeviceSynchronize());
    cudaFree(p->d_params); cudaFree(p->d_out); cudaFree(p->d_skip); cudaFree(p->d_points);
    cudaFreeHost(p->h_params); cudaFreeHost(p->h_out); cudaFreeHost(p->h_skip);
    p->d_params = p->d_out = nullptr; p->d_skip = nullptr;
    p->h_params = p->h_out = nullptr; p->h_skip = nullptr; p->stage_cap = 0;
    const size_t out_row = 1 + 3 * (size_t)nd;
    CK(cudaMalloc(&p->d_params, sizeof(double) * GWIN_NPARAM * cap));
    CK(cudaMalloc(&p->d_points, sizeof(double) * GWIN_NPARAM * cap));
    CK(cudaMalloc(&p->d_out, sizeof(double) * out_row * cap));
    CK(cudaMalloc(&p->d_skip, cap));
    CK(cudaMallocHost(&p->h_params, sizeof(double) * GWIN_NPARAM * cap));
    CK(cudaMallocHost(&p->h_out, sizeof(double) * (out_row * cap + 2)));
    CK(cudaMallocHost(&p->h_skip, cap));
    p->stage_cap = cap;
    return GWIN_OK;
}

extern "C" int gwin_loglr_batch_host(gwin_plan* p, int mode, int64_t W,
                                     const double* params, const uint8_t* skip,
                                     double* loglr, double* hd, double* hh) {
    return gwin_loglr_batch_host_strided(p, mode, W, params, GWIN_NPARAM, 1, skip, loglr, hd, hh);
}

// params is either [W][13] (stride_walker = 13, stride_param = 1) or [13][W] (1, W); the
// buffer is copied as one contiguous block of 13*W doubles
static int loglr_batch_host_impl(gwin_plan* p, int mode, int64_t W,
                                 const double* params, int64_t stride_walker, int64_t stride_param,
                                 const double* calib, int64_t cstride_walker, int64_t cstride_value,
                                 const uint8_t* skip, double* loglr, double* hd, double* hh);

extern "C" int gwin_loglr_batch_host_strided(gwin_plan* p, int mode, int64_t W,
                                             const double* params, int64_t stride_walker, int64_t stride_param,
                                             const uint8_t* skip,
                                             double* loglr, double* hd, double* hh) {
    return loglr_batch_host_impl(p, mode, W, params, stride_walker, stride_param, nullptr, 0, 0, skip, loglr, hd, hh);
}

extern "C" int gwin_loglr_batch_calib_host(gwin_plan* p, int mode, int64_t W,
                                           const double* params, int64_t stride_walker, int64_t stride_param,
                                           const double* calib, int64_t cstride_walker, int64_t cstride_value,
                                           const uint8_t* skip,
                                           double* loglr, double* hd, double* hh) {
    if (p && calib) {
        const int64_t nv = (int64_t)p->d.n_det * 2 * p->d.n_cal;
        if (nv == 0) return fail(GWIN_EINVAL, "calibration values given but gwin_plan_set_calibration was not called");
        if (!((cstride_walker == nv && cstride_value == 1) || (cstride_walker == 1 && cstride_value == W)))
            return fail(GWIN_EINVAL, "host calibration values must be a dense [W][%lld] or [%lld][W] block",
                        (long long)nv, (long long)nv);
    }
    return loglr_batch_host_impl(p, mode, W, params, stride_walker, stride_param, calib, cstride_walker, cstride_value,
                                 skip, loglr, hd, hh);
}

static int ensure_calib_stage(gwin_plan* p, size_t n) {
    if (n <= p->calib_cap) return GWIN_OK;
    CK(cudaDeviceSynchronize());
    cudaFree(p->d_calib); cudaFreeHost(p->h_calib);
    p->d_calib = p->h_calib = nullptr; p->calib_cap = 0;
    const size_t cap = std::max<size_t>(n, 4096);
    CK(cudaMalloc(&p->d_calib, sizeof(double) * cap));
    CK(cudaMallocHost(&p->h_calib, sizeof(double) * cap));
    p->calib_cap = cap;
    return GWIN_OK;
}

static int loglr_batch_host_impl(gwin_plan* p, int mode, int64_t W,
                                 const double* params, int64_t stride_walker, int64_t stride_param,
                                 const double* calib, int64_t cstride_walker, int64_t cstride_value,
                                 const uint8_t* skip, double* loglr, double* hd, double* hh) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (!((stride_walker == GWIN_NPARAM && stride_param == 1) || (stride_walker == 1 && stride_param == W)))
        return fail(GWIN_EINVAL, "host parameters must be a dense [W][13] or [13][W] block");
    if (W < 0) return fail(GWIN_EINVAL, "W < 0");
    if (W == 0) { p->last_launches = 0; return GWIN_OK; }
    if (!params || !loglr) return fail(GWIN_EINVAL, "params/loglr is NULL");
    CK(cudaSetDevice(p->device));
    int rc = ensure_stage(p, W);
    if (rc) return rc;
    const int nd = p->d.n_det;
    cudaStream_t st = p->stream;
    memcpy(p->h_params, params, sizeof(double) * GWIN_NPARAM * W);
    const bool table = p->kernel != GWIN_KERNEL_DIRECT && table_wanted(p, W, calib != nullptr);
    if (p->auto_mchirp && p->kernel != GWIN_KERNEL_DIRECT && !table) {
        // the block schedule must cover the smallest chirp mass of the batch: quantise it
        // down to a quarter octave so the (cheap) schedule rebuild is rare.  (On the table path the
        // coverage measured on the device sets it.)
        double mc5_min = 1e300;                   // chirp mass^5 = (m1 m2)^3 / (m1 + m2): no pow per walker
        for (int64_t i = 0; i < W; ++i) {
            if (skip && skip[i]) continue;
            const double m1 = params[i * stride_walker], m2 = params[i * stride_walker + stride_param];
            if (!(m1 > 0.0) || !(m2 > 0.0) || !std::isfinite(m1 + m2)) continue;
            const double pm = m1 * m2;
            mc5_min = std::min(mc5_min, pm * pm * pm / (m1 + m2));
        }
        double mc_min = mc5_min < 1e300 ? std::pow(mc5_min, 0.2) : 1e300;
        if (mc_min < 1e300) {
            mc_min = std::min(std::max(mc_min, 0.02), 1e4);
            const double q = std::exp2(std::floor(4.0 * std::log2(mc_min)) / 4.0);
            if (q != p->mchirp_min) { p->mchirp_min = q; p->sched_dirty = true; }
        }
    }
    CK(cudaMemcpyAsync(p->d_params, p->h_params, sizeof(double) * GWIN_NPARAM * W, cudaMemcpyHostToDevice, st));
    if (calib) {
        const size_t nv = (size_t)nd * 2 * p->d.n_cal * (size_t)W;
        rc = ensure_calib_stage(p, nv);
        if (rc) return rc;
        memcpy(p->h_calib, calib, sizeof(double) * nv);
        CK(cudaMemcpyAsync(p->d_calib, p->h_calib, sizeof(double) * nv, cudaMemcpyHostToDevice, st));
    }
    if (skip) {
        memcpy(p->h_skip, skip, W);
        CK(cudaMemcpyAsync(p->d_skip, p->h_skip, W, cudaMemcpyHostToDevice, st));
    }
    double* d_lr = p->d_out;
    double* d_hd = p->d_out + p->stage_cap;
    double* d_hh = d_hd + 2 * (size_t)nd * p->stage_cap;
    rc = loglr_batch_impl(p, mode, W, p->d_params, stride_walker, stride_param,
                          calib ? p->d_calib : nullptr, cstride_walker, cstride_value,
                          skip ? p->d_skip : nullptr, d_lr, hd ? d_hd : nullptr, hh ? d_hh : nullptr, st);
    if (rc) return rc;
    double* h_lr = p->h_out;
    double* h_hd = p->h_out + p->stage_cap;
    double* h_hh = h_hd + 2 * (size_t)nd * p->stage_cap;
    int* h_slow = reinterpret_cast<int*>(p->h_out + (1 + 3 * (size_t)nd) * p->stage_cap);
    *h_slow = 0;
    CK(cudaMemcpyAsync(h_lr, d_lr, sizeof(double) * W, cudaMemcpyDeviceToHost, st));
    if (hd) CK(cudaMemcpyAsync(h_hd, d_hd, sizeof(double) * 2 * nd * W, cudaMemcpyDeviceToHost, st));
    if (hh) CK(cudaMemcpyAsync(h_hh, d_hh, sizeof(double) * nd * W, cudaMemcpyDeviceToHost, st));
    if (p->kernel != GWIN_KERNEL_DIRECT) CK(cudaMemcpyAsync(h_slow, p->d_slow_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // a batch that sent more than a few walkers down the slow path: re-measure what the schedule and
    // the tables have to cover before the next batch (this batch's results are right as they are)
    if (*h_slow > 4 + W / 64 && !p->cover_manual) {
        if (table) p->replan = true;
    }
    memcpy(loglr, h_lr, sizeof(double) * W);
    if (hd) memcpy(hd, h_hd, sizeof(double) * 2 * nd * W);
    if (hh) memcpy(hh, h_hh, sizeof(double) * nd * W);
    return GWIN_OK;
}

static bool host_pinned(const void* ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int gwin_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(GWIN_EINVAL, "NULL argument");
    *ptr = nullptr;
    CK(cudaMallocHost(ptr, bytes ? bytes : 1));
    return GWIN_OK;
}
extern "C" void gwin_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

extern "C" int gwin_loglr_batch_points_host(gwin_plan* p, int mode, int64_t W,
                                            const double* points, int ndim, int64_t stride_walker, int64_t stride_col,
                                            const int* map, const double* fixed, const uint8_t* skip,
                                            double* loglr, double* hd, double* hh) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (W < 0) return fail(GWIN_EINVAL, "W < 0");
    if (W == 0) { p->last_launches = 0; return GWIN_OK; }
    if (!map || !fixed || !loglr) return fail(GWIN_EINVAL, "map/fixed/loglr is NULL");
    if (ndim < 0 || ndim > GWIN_NPARAM) return fail(GWIN_EINVAL, "ndim must be 0..%d", GWIN_NPARAM);
    if (ndim > 0 && !points) return fail(GWIN_EINVAL, "points is NULL");
    if (ndim > 0 && !((stride_walker == ndim && stride_col == 1) || (stride_walker == 1 && stride_col == W)))
        return fail(GWIN_EINVAL, "host points must be a dense [W][ndim] or [ndim][W] block");
    PointMap M;
    for (int i = 0; i < GWIN_NPARAM; ++i) {
        if (map[i] >= ndim) return fail(GWIN_EINVAL, "map[%d] = %d is not a column of the points", i, map[i]);
        M.col[i] = map[i] < 0 ? -1 : map[i];
        M.fixed[i] = fixed[i];
    }
    CK(cudaSetDevice(p->device));
    int rc = ensure_stage(p, W);
    if (rc) return rc;
    const int nd = p->d.n_det;
    cudaStream_t st = p->stream;
    // parameters: straight from the caller's buffer when it is page-locked, else through the staging buffer
    if (ndim > 0) {
        const double* src = points;
        if (!host_pinned(points)) {
            memcpy(p->h_params, points, sizeof(double) * (size_t)ndim * W);
            src = p->h_params;
        }
        CK(cudaMemcpyAsync(p->d_points, src, sizeof(double) * (size_t)ndim * W, cudaMemcpyHostToDevice, st));
    }
    if (skip) {
        memcpy(p->h_skip, skip, W);
        CK(cudaMemcpyAsync(p->d_skip, p->h_skip, W, cudaMemcpyHostToDevice, st));
    }
    expand_points_kernel<<<(int)((W + 127) / 128), 128, 0, st>>>((int)W, p->d_points, stride_walker, stride_col, M, p->d_params);
    const bool table = p->kernel != GWIN_KERNEL_DIRECT && table_wanted(p, W, false);
    if (p->auto_mchirp && p->kernel != GWIN_KERNEL_DIRECT && !table) {
        // (as in the plain host entry point: the schedule follows the lightest system of the batch)
        const int c1 = M.col[0], c2 = M.col[1];
        double mc5_min = 1e300;
        for (int64_t i = 0; i < W; ++i) {
            if (skip && skip[i]) continue;
            const double m1 = c1 >= 0 ? points[i * stride_walker + c1 * stride_col] : M.fixed[0];
            const double m2 = c2 >= 0 ? points[i * stride_walker + c2 * stride_col] : M.fixed[1];
            if (!(m1 > 0.0) || !(m2 > 0.0) || !std::isfinite(m1 + m2)) continue;
            const double pm = m1 * m2;
            mc5_min = std::min(mc5_min, pm * pm * pm / (m1 + m2));
            if (c1 < 0 && c2 < 0) break;
        }
        if (mc5_min < 1e300) {
            const double mc_min = std::min(std::max(std::pow(mc5_min, 0.2), 0.02), 1e4);
            const double q = std::exp2(std::floor(4.0 * std::log2(mc_min)) / 4.0);
            if (q != p->mchirp_min) { p->mchirp_min = q; p->sched_dirty = true; }
        }
    }
    double* d_lr = p->d_out;
    double* d_hd = p->d_out + p->stage_cap;
    double* d_hh = d_hd + 2 * (size_t)nd * p->stage_cap;
    rc = loglr_batch_impl(p, mode, W, p->d_params, 1, W, nullptr, 0, 0, skip ? p->d_skip : nullptr,
                          d_lr, hd ? d_hd : nullptr, hh ? d_hh : nullptr, st);
    if (rc) return rc;
    p->last_launches += 1;
    double* h_lr = p->h_out;
    double* h_hd = p->h_out + p->stage_cap;
    double* h_hh = h_hd + 2 * (size_t)nd * p->stage_cap;
    int* h_slow = reinterpret_cast<int*>(p->h_out + (1 + 3 * (size_t)nd) * p->stage_cap);
    *h_slow = 0;
    const bool direct_out = host_pinned(loglr);
    CK(cudaMemcpyAsync(direct_out ? loglr : h_lr, d_lr, sizeof(double) * W, cudaMemcpyDeviceToHost, st));
    if (hd) CK(cudaMemcpyAsync(h_hd, d_hd, sizeof(double) * 2 * nd * W, cudaMemcpyDeviceToHost, st));
    if (hh) CK(cudaMemcpyAsync(h_hh, d_hh, sizeof(double) * nd * W, cudaMemcpyDeviceToHost, st));
    if (p->kernel != GWIN_KERNEL_DIRECT) CK(cudaMemcpyAsync(h_slow, p->d_slow_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (*h_slow > 4 + W / 64 && !p->cover_manual && table) p->replan = true;
    if (!direct_out) memcpy(loglr, h_lr, sizeof(double) * W);
    if (hd) memcpy(hd, h_hd, sizeof(double) * 2 * nd * W);
    if (hh) memcpy(hh, h_hh, sizeof(double) * nd * W);
    return GWIN_OK;
}

// ---- asynchronous host entry -------------------------------------------------------------------
static int async_setup(gwin_plan* p) {
    if (p->async_ready) return GWIN_OK;
    memset(p->aslot, 0, sizeof p->aslot);
    CK(cudaStreamCreateWithFlags(&p->stream_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->stream_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&p->aslot[i].ev_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->aslot[i].ev_k, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->aslot[i].ev_out, cudaEventDisableTiming));
    }
    p->async_ready = true;
    return GWIN_OK;
}

static int async_stage(gwin_plan* p, gwin_plan::AsyncSlot& a, int64_t W) {
    if (W <= a.cap) return GWIN_OK;
    const int nd = p->d.n_det;
    const int64_t cap = std::max<int64_t>(W, 1024);
    const size_t out_row = 1 + 3 * (size_t)nd;
    // (the slot is idle: its previous call has been waited for; cudaFree orders itself after queued work)
    cudaFree(a.d_points); cudaFree(a.d_params); cudaFree(a.d_out); cudaFree(a.d_skip);
    cudaFreeHost(a.h_points); cudaFreeHost(a.h_out); cudaFreeHost(a.h_skip);
    a.d_points = a.d_params = a.d_out = a.h_points = a.h_out = nullptr; a.d_skip = a.h_skip = nullptr; a.cap = 0;
    CK(cudaMalloc(&a.d_points, sizeof(double) * GWIN_NPARAM * cap));
    CK(cudaMalloc(&a.d_params, sizeof(double) * GWIN_NPARAM * cap));
    CK(cudaMalloc(&a.d_out, sizeof(double) * out_row * cap));
    CK(cudaMalloc(&a.d_skip, cap));
    CK(cudaMallocHost(&a.h_points, sizeof(double) * GWIN_NPARAM * cap));
    CK(cudaMallocHost(&a.h_out, sizeof(double) * (out_row * cap + 2)));
    CK(cudaMallocHost(&a.h_skip, cap));
    a.cap = cap;
    return GWIN_OK;
}

extern "C" int gwin_loglr_batch_points_submit(gwin_plan* p, int mode, int64_t W,
                                              const double* points, int ndim, int64_t stride_walker, int64_t stride_col,
                                              const int* map, const double* fixed, const uint8_t* skip,
                                              double* loglr, double* hd, double* hh, int* ticket) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (!ticket) return fail(GWIN_EINVAL, "ticket is NULL");
    *ticket = -1;
    if (W <= 0) return fail(GWIN_EINVAL, "W must be positive");
    if (!map || !fixed || !loglr) return fail(GWIN_EINVAL, "map/fixed/loglr is NULL");
    if (ndim < 0 || ndim > GWIN_NPARAM) return fail(GWIN_EINVAL, "ndim must be 0..%d", GWIN_NPARAM);
    if (ndim > 0 && !points) return fail(GWIN_EINVAL, "points is NULL");
    if (ndim > 0 && !((stride_walker == ndim && stride_col == 1) || (stride_walker == 1 && stride_col == W)))
        return fail(GWIN_EINVAL, "host points must be a dense [W][ndim] or [ndim][W] block");
    PointMap M;
    for (int i = 0; i < GWIN_NPARAM; ++i) {
        if (map[i] >= ndim) return fail(GWIN_EINVAL, "map[%d] = %d is not a column of the points", i, map[i]);
        M.col[i] = map[i] < 0 ? -1 : map[i];
        M.fixed[i] = fixed[i];
    }
    CK(cudaSetDevice(p->device));
    int rc = async_setup(p);
    if (rc) return rc;
    const int s = !p->aslot[0].busy ? 0 : (!p->aslot[1].busy ? 1 : -1);
    if (s < 0) return fail(GWIN_EINVAL, "two calls are already in flight: gwin_loglr_batch_points_wait for one of them first");
    gwin_plan::AsyncSlot& a = p->aslot[s];
    rc = async_stage(p, a, W);
    if (rc) return rc;
    const int nd = p->d.n_det;
    if (ndim > 0) {
        const double* src = points;
        if (!host_pinned(points)) {
            memcpy(a.h_points, points, sizeof(double) * (size_t)ndim * W);
            src = a.h_points;
        }
        CK(cudaMemcpyAsync(a.d_points, src, sizeof(double) * (size_t)ndim * W, cudaMemcpyHostToDevice, p->stream_in));
    }
    if (skip) {
        memcpy(a.h_skip, skip, W);
        CK(cudaMemcpyAsync(a.d_skip, a.h_skip, W, cudaMemcpyHostToDevice, p->stream_in));
    }
    CK(cudaEventRecord(a.ev_in, p->stream_in));
    cudaStream_t st = p->stream;
    CK(cudaStreamWaitEvent(st, a.ev_in, 0));
    expand_points_kernel<<<(int)((W + 127) / 128), 128, 0, st>>>((int)W, a.d_points, stride_walker, stride_col, M, a.d_params);
    const bool table = p->kernel != GWIN_KERNEL_DIRECT && table_wanted(p, W, false);
    if (p->auto_mchirp && p->kernel != GWIN_KERNEL_DIRECT && !table) {
        const int c1 = M.col[0], c2 = M.col[1];
        double mc5_min = 1e300;
        for (int64_t i = 0; i < W; ++i) {
            if (skip && skip[i]) continue;
            const double m1 = c1 >= 0 ? points[i * stride_walker + c1 * stride_col] : M.fixed[0];
            const double m2 = c2 >= 0 ? points[i * stride_walker + c2 * stride_col] : M.fixed[1];
            if (!(m1 > 0.0) || !(m2 > 0.0) || !std::isfinite(m1 + m2)) continue;
            const double pm = m1 * m2;
            mc5_min = std::min(mc5_min, pm * pm * pm / (m1 + m2));
            if (c1 < 0 && c2 < 0) break;
        }
        if (mc5_min < 1e300) {
            const double mc_min = std::min(std::max(std::pow(mc5_min, 0.2), 0.02), 1e4);
            const double q = std::exp2(std::floor(4.0 * std::log2(mc_min)) / 4.0);
            if (q != p->mchirp_min) { p->mchirp_min = q; p->sched_dirty = true; }
        }
    }
    double* d_lr = a.d_out;
    double* d_hd = a.d_out + a.cap;
    double* d_hh = d_hd + 2 * (size_t)nd * a.cap;
    rc = loglr_batch_impl(p, mode, W, a.d_params, 1, W, nullptr, 0, 0, skip ? a.d_skip : nullptr,
                          d_lr, hd ? d_hd : nullptr, hh ? d_hh : nullptr, st);
    if (rc) return rc;
    p->last_launches += 1;
    int* h_slow = reinterpret_cast<int*>(a.h_out + (1 + 3 * (size_t)nd) * a.cap);
    *h_slow = 0;
    // (this batch's slow-path count: read on the compute stream, before the next batch's kernels reset it)
    if (p->kernel != GWIN_KERNEL_DIRECT) CK(cudaMemcpyAsync(h_slow, p->d_slow_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(a.ev_k, st));
    cudaStream_t so = p->stream_out;
    CK(cudaStreamWaitEvent(so, a.ev_k, 0));
    double* h_lr = a.h_out;
    double* h_hd = a.h_out + a.cap;
    double* h_hh = h_hd + 2 * (size_t)nd * a.cap;
    a.direct_out = host_pinned(loglr);
    CK(cudaMemcpyAsync(a.direct_out ? loglr : h_lr, d_lr, sizeof(double) * W, cudaMemcpyDeviceToHost, so));
    if (hd) CK(cudaMemcpyAsync(h_hd, d_hd, sizeof(double) * 2 * nd * W, cudaMemcpyDeviceToHost, so));
    if (hh) CK(cudaMemcpyAsync(h_hh, d_hh, sizeof(double) * nd * W, cudaMemcpyDeviceToHost, so));
    CK(cudaEventRecord(a.ev_out, so));
    a.busy = true; a.table = table; a.W = W; a.loglr = loglr; a.hd = hd; a.hh = hh;
    *ticket = s;
    return GWIN_OK;
}

extern "C" int gwin_loglr_batch_points_wait(gwin_plan* p, int ticket) {
    if (!p) return fail(GWIN_EINVAL, "plan is NULL");
    if (ticket < 0 || ticket > 1 || !p->async_ready || !p->aslot[ticket].busy)
        return fail(GWIN_EINVAL, "no call in flight with ticket %d", ticket);
    gwin_plan::AsyncSlot& a = p->aslot[ticket];
    CK(cudaSetDevice(p->device));
    CK(cudaEventSynchronize(a.ev_out));
    a.busy = false;
    const int nd = p->d.n_det;
    const int64_t W = a.W;
    const double* h_lr = a.h_out;
    const double* h_hd = a.h_out + a.cap;
    const double* h_hh = h_hd + 2 * (size_t)nd * a.cap;
    const int* h_slow = reinterpret_cast<const int*>(a.h_out + (1 + 3 * (size_t)nd) * a.cap);
    if (*h_slow > 4 + W / 64 && !p->cover_manual && a.table) p->replan = true;
    if (!a.direct_out) memcpy(a.loglr, h_lr, sizeof(double) * W);
    if (a.hd) memcpy(a.hd, h_hd, sizeof(double) * 2 * nd * W);
    if (a.hh) memcpy(a.hh, h_hh, sizeof(double) * nd * W);
    return GWIN_OK;
}

extern "C" int gwin_generate_host(gwin_plan* p, const double* params, double* h_out, int64_t* length) {
    return gwin_generate_calib_host(p, params, nullptr, h_out, length);
}

extern "C" int gwin_generate_calib_host(gwin_plan* p, const double* params, const double* calib,
                                        double* h_out, int64_t* length) {
    if (!p || !params || !h_out) return fail(GWIN_EINVAL, "NULL argument");
    if (calib && p->d.n_cal == 0)
        return fail(GWIN_EINVAL, "calibration values given but gwin_plan_set_calibration was not called");
    CK(cudaSetDevice(p->device));
    int rc = ensure_stage(p, 1); if (rc) return rc;
    const size_t nv = (size_t)p->d.n_det * 2 * p->d.n_cal;
    if (calib) { rc = ensure_calib_stage(p, nv); if (rc) return rc; }
    rc = ensure_workspace(p, 2, 1); if (rc) return rc;
    const PlanDev& d = p->d;
    const int wpad = (int)p->w_cap;
    cudaStream_t st = p->stream;
    memcpy(p->h_params, params, sizeof(double) * GWIN_NPARAM);
    CK(cudaMemcpyAsync(p->d_params, p->h_params, sizeof(double) * GWIN_NPARAM, cudaMemcpyHostToDevice, st));
    if (calib) {
        memcpy(p->h_calib, calib, sizeof(double) * nv);
        CK(cudaMemcpyAsync(p->d_calib, p->h_calib, sizeof(double) * nv, cudaMemcpyHostToDevice, st));
    }
    double* hh_rows = p->d_wc + (size_t)NCOEF(d.n_det) * wpad;
    FastArgs F;
    memset(&F, 0, sizeof F);
    walker_setup_kernel<<<1, 32, 0, st>>>(d, 1, wpad, p->d_params, GWIN_NPARAM, 1, nullptr, nullptr, p->d_wc, p->d_ks, p->d_ke, hh_rows,
                                          calib ? p->d_calib : nullptr, (int64_t)nv, 1, F);
    generate_len_kernel<<<1, 1, 0, st>>>(d, p->d_params, p->d_ke);
    double2* dout = nullptr;
    CK(cudaMalloc(&dout, sizeof(double2) * (size_t)d.n_det * d.n_f));
    generate_kernel<<<(d.n_f + 255) / 256, 256, 0, st>>>(d, wpad, p->d_wc, p->d_ke, dout, calib != nullptr);
    int n = 0;
    auto fetch = [&]() -> int {
        CK(cudaMemcpyAsync(h_out, dout, sizeof(double2) * (size_t)d.n_det * d.n_f, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&n, p->d_ke + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return GWIN_OK;
    };
    rc = fetch();
    cudaFree(dout);
    if (rc) return rc;
    if (length) *length = n;
    p->last_launches = 3;
    return GWIN_OK;
}

template <int V>
static int run_probe(int device, double* rate) {
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    const double per_iter = V == 2 ? 16.0 : 8.0;      // FP64 instructions (3, 4: MMAs) per thread per iteration
    double* out = nullptr;
    CK(cudaMalloc(&out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dfma_probe_kernel<V><<<blocks, threads>>>(out, 1024);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        dfma_probe_kernel<V><<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, (double)threads * blocks * iters * per_iter / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *rate = best;
    return GWIN_OK;
}

extern "C" int gwin_fp64_peak(int device, double* dfma_per_sec) {
    if (!dfma_per_sec) return fail(GWIN_EINVAL, "NULL argument");
    return run_probe<0>(device, dfma_per_sec);
}

extern "C" int gwin_fp64_probe(int device, int variant, double* ops_per_sec) {
    if (!ops_per_sec) return fail(GWIN_EINVAL, "NULL argument");
    switch (variant) {
        case 0: return run_probe<0>(device, ops_per_sec);
        case 1: return run_probe<1>(device, ops_per_sec);
        case 2: return run_probe<2>(device, ops_per_sec);
        case 3: return run_probe<3>(device, ops_per_sec);
        case 4: return run_probe<4>(device, ops_per_sec);
        default: return fail(GWIN_EINVAL, "variant must be 0..4");
    }
}
