/*
 * oracle/loglr_ref.c -- plain-C restatement of the reference's per-walker loglr
 * chain.  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/gwin_oracle.py header):
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build or call it.  PARITY UNPINNED upstream (the reference's
 * arithmetic lives in un-vendored pycbc >= 1.12.1 / lalsuite, not installed here).
 *
 * It keeps the reference's PASS STRUCTURE -- one array pass per step, with the
 * intermediate complex128 arrays the Python stack allocates -- so that its speed
 * is an honest stand-in for "gwin + pycbc + lalsimulation on one core":
 *
 *   1. XLALSimInspiralTaylorF2Core loop      [UPSTREAM]  cbrt, log, cos, sin per bin
 *   2. ChooseFDWaveform  hp = pfac*h, hc = -i cfac*h     (2 arrays)
 *   3. per detector: thish = fp*hp + fc*hc               generator.py [UPSTREAM]
 *   4. apply_fseries_time_shift: phasor recurrence, re-sync every 100 bins
 *   5. h[kmin:kmax] *= weight                            gaussian_noise.py:337
 *   6. <d|h> = data.inner(h), <h|h> = h.inner(h).real    gaussian_noise.py:339-340
 *   7. combine                                           gaussian_noise.py:341-350,
 *                                                        marginalized_gaussian_noise.py:456-511
 *
 * Walkers are distributed over pthreads the way the reference distributes them
 * over processes (option_utils.py:171-185).
 */
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LAL_PI 3.141592653589793238462643383279502884
#define LAL_TWOPI 6.283185307179586476925286766559005768
#define LAL_PI_4 0.785398163397448309615660845819875721
#define LAL_GAMMA 0.577215664901532860606512090082402431
#define LAL_C_SI 299792458.0
#define LAL_MTSUN_SI 4.925491025543575903411922162094833998e-6
#define LAL_MRSUN_SI 1.476625061404649406193430731479084713e3
#define LAL_PC_SI 3.085677581491367278913937957796471611e16
#define MAX_DET 4

typedef double complex cplx;

typedef struct {
    int n_det;
    long n_f, kmin, kmax;
    double delta_f, epoch, f_lower_wf;
    int leaps;
    double resp[MAX_DET][9], loc[MAX_DET][3];
    double* weight[MAX_DET]; /* sqrt(norm/psd)               gaussian_noise.py:253-262 */
    cplx* data[MAX_DET];     /* whitened copy of the data     gaussian_noise.py:264-265 */
    double lognl_det[MAX_DET], lognl;
} oracle_model;

static const long LEAPS[] = {46828800, 78364801, 109900802, 173059203, 252028804, 315187205,
    346723206, 394156807, 425692808, 457228809, 504489610, 551750411, 599184012, 820108813,
    914803214, 1025136015, 1119744016, 1167264017};

oracle_model* oracle_create(int n_det, long n_f, double delta_f, double epoch, double f_lower,
                            double f_upper, double f_lower_wf, const double* resp,
                            const double* loc, const double* data, const double* psd,
                            double norm) {
    if (n_det < 1 || n_det > MAX_DET) return NULL;
    oracle_model* m = calloc(1, sizeof *m);
    m->n_det = n_det; m->n_f = n_f; m->delta_f = delta_f; m->epoch = epoch;
    m->f_lower_wf = f_lower_wf;
    /* pycbc.filter.get_cutoff_indices(f_lower, f_upper, df, (n_f-1)*2)   gaussian_noise.py:247 */
    long N = (n_f - 1) * 2;
    m->kmin = f_lower ? (long)(f_lower / delta_f) : 1;
    m->kmax = (long)((N + 1) / 2.);
    if (f_upper > 0 && (long)(f_upper / delta_f) < m->kmax) m->kmax = (long)(f_upper / delta_f);
    if (m->kmax <= m->kmin) { free(m); return NULL; }
    if (!(norm > 0)) norm = 4 * delta_f;
    for (size_t i = 0; i < sizeof LEAPS / sizeof *LEAPS; ++i)
        if ((long)floor(epoch) >= LEAPS[i]) m->leaps++;
    memcpy(m->resp, resp, sizeof(double) * 9 * n_det);
    memcpy(m->loc, loc, sizeof(double) * 3 * n_det);
    m->lognl = 0;
    for (int d = 0; d < n_det; ++d) {
        m->weight[d] = malloc(sizeof(double) * n_f);
        m->data[d] = malloc(sizeof(cplx) * n_f);
        for (long k = 0; k < n_f; ++k) {
            m->weight[d][k] = psd ? sqrt(norm / psd[d * n_f + k]) : sqrt(norm);
            m->data[d][k] = data[2 * (d * n_f + k)] + I * data[2 * (d * n_f + k) + 1];
        }
        for (long k = m->kmin; k < m->kmax; ++k) m->data[d][k] *= m->weight[d][k];
        /* _lognl  gaussian_noise.py:275-291 */
        cplx s = 0;
        for (long k = m->kmin; k < m->kmax; ++k) s += conj(m->data[d][k]) * m->data[d][k];
        m->lognl_det[d] = -0.5 * creal(s);
        m->lognl += m->lognl_det[d];
    }
    return m;
}

void oracle_destroy(oracle_model* m) {
    if (!m) return;
    for (int d = 0; d < m->n_det; ++d) { free(m->weight[d]); free(m->data[d]); }
    free(m);
}

double oracle_lognl(const oracle_model* m, double* per_det) {
    if (per_det) for (int d = 0; d < m->n_det; ++d) per_det[d] = m->lognl_det[d];
    return m->lognl;
}

void oracle_cutoffs(const oracle_model* m, long* kmin, long* kmax) { *kmin = m->kmin; *kmax = m->kmax; }

/* [UPSTREAM] XLALGreenwichMeanSiderealTime: integer-second Julian day + nanoseconds */
static double gmst_lal(double t_gps, int leaps) {
    double sec_f = floor(t_gps);
    double ns = rint((t_gps - sec_f) * 1e9);
    long sec = (long)sec_f;
    if (ns >= 1e9) { sec += 1; ns -= 1e9; }
    long utc = sec - leaps;
    long days = utc / 86400, sod = utc % 86400;
    if (sod < 0) { sod += 86400; days -= 1; }
    /* XLALConvertCivilTimeToJD of GPS day 0 (1980-01-06) gives 2444245 before "- 0.5" */
    volatile double jd = (double)(2444245L + days);
    jd += (double)sod / 86400.0 - 0.5;
    double t_hi = (jd - 2451545.0) / 36525.0;
    double t_lo = ns / (1e9 * 36525.0 * 86400.0);
    double t = t_hi + t_lo;
    double st = (-6.2e-6 * t + 0.093104) * t * t + 67310.54841;
    st += 8640184.812866 * t_lo;
    st += 3155760000.0 * t_lo;
    st += 8640184.812866 * t_hi;
    st += 3155760000.0 * t_hi;
    return st * LAL_PI / 43200.0;
}

/* log I0(x) = what numpy.log(special.i0(x)) returns where that is finite */
static double log_i0(double x) {
    x = fabs(x);
    if (x < 600.0) {
        double q = 0.25 * x * x, term = 1.0, sum = 1.0;
        for (int k = 1; k < 2000; ++k) {
            term *= q / ((double)k * (double)k);
            sum += term;
            if (term < 1e-17 * sum) break;
        }
        return log(sum);
    }
    double y = 1.0 / (8.0 * x);
    double ser = 1.0 + y * (1.0 + y * (4.5 + y * (37.5 + y * (459.375 + y * (7441.875 + y * 150077.8125)))));
    return x - 0.5 * log(LAL_TWOPI * x) + log(ser);
}

typedef struct { cplx *h, *hp, *hc, *hd; } scratch;

/* one walker; returns loglr, fills hd[n_det] (complex) and hh[n_det] */
static double one_loglr(const oracle_model* m, int mode, const double* p, scratch* s,
                        cplx* hd_out, double* hh_out) {
    const double m1 = p[0], m2 = p[1], chi1L = p[2], chi2L = p[3], distance = p[4];
    const double incl = p[5], phi_ref = p[6], ra = p[7], dec = p[8], psi = p[9], tc = p[10];
    const double lambda1 = p[11], lambda2 = p[12];
    const int nd = m->n_det;
    for (int d = 0; d < nd; ++d) { hd_out[d] = 0; hh_out[d] = 0; }
    for (int i = 0; i < 13; ++i) if (!isfinite(p[i])) return -INFINITY;
    if (!(m1 > 0) || !(m2 > 0) || !(distance > 0)) return -INFINITY;
    /* ---- [UPSTREAM] XLALSimInspiralPNPhasing_F2 ---- */
    const double mtot = m1 + m2, dm = (m1 - m2) / (m1 + m2);
    const double eta = m1 * m2 / mtot / mtot, m1M = m1 / mtot, m2M = m2 / mtot;
    const double pfaN = 3.0 / (128.0 * eta);
    double v[8] = {0}, vl[8] = {0};
    v[0] = 1.0;
    v[2] = 5.0 * (743.0 / 84.0 + 11.0 * eta) / 9.0;
    v[3] = -16.0 * LAL_PI;
    v[4] = 5.0 * (3058.673 / 7.056 + 5429.0 / 7.0 * eta + 617.0 * eta * eta) / 72.0;
    v[5] = 5.0 / 9.0 * (7729.0 / 84.0 - 13.0 * eta) * LAL_PI;
    vl[5] = 5.0 / 3.0 * (7729.0 / 84.0 - 13.0 * eta) * LAL_PI;
    v[6] = (11583.231236531 / 4.694215680 - 640.0 / 3.0 * LAL_PI * LAL_PI - 6848.0 / 21.0 * LAL_GAMMA)
         + eta * (-15737.765635 / 3.048192 + 2255.0 / 12.0 * LAL_PI * LAL_PI)
         + eta * eta * 76055.0 / 1728.0 - eta * eta * eta * 127825.0 / 1296.0;
    v[6] += (-6848.0 / 21.0) * log(4.0);
    vl[6] = -6848.0 / 21.0;
    v[7] = LAL_PI * (77096675.0 / 254016.0 + 378515.0 / 1512.0 * eta - 74045.0 / 756.0 * eta * eta);
    {
        const double chi1sq = chi1L * chi1L, chi2sq = chi2L * chi2L, chi1dotchi2 = chi1L * chi2L;
        const double SL = m1M * m1M * chi1L + m2M * m2M * chi2L;
        const double dSigmaL = dm * (m2M * chi2L - m1M * chi1L);
        double pn_sigma = eta * (721.0 / 48.0 * chi1L * chi2L - 247.0 / 48.0 * chi1dotchi2);
        pn_sigma += (720.0 - 1.0) / 96.0 * m1M * m1M * chi1L * chi1L;
        pn_sigma += (720.0 - 1.0) / 96.0 * m2M * m2M * chi2L * chi2L;
        pn_sigma -= (240.0 - 7.0) / 96.0 * m1M * m1M * chi1sq;
        pn_sigma -= (240.0 - 7.0) / 96.0 * m2M * m2M * chi2sq;
        double pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1L * chi2L;
        pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m1M - 120.0 * m1M * m1M)
                   + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1sq;
        pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m2M - 120.0 * m2M * m2M)
                   + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2sq;
        const double pn_gamma = (554345.0 / 1134.0 + 110.0 * eta / 9.0) * SL
                              + (13915.0 / 84.0 - 10.0 * eta / 3.0) * dSigmaL;
        v[7] += (-8980424995.0 / 762048.0 + 6586595.0 * eta / 756.0 - 305.0 * eta * eta / 36.0) * SL
              - (170978035.0 / 48384.0 - 2876425.0 * eta / 672.0 - 4735.0 * eta * eta / 144.0) * dSigmaL;
        v[6] += LAL_PI * (3760.0 * SL + 1490.0 * dSigmaL) / 3.0 + pn_ss3;
        v[5] += -1.0 * pn_gamma;
        vl[5] += -3.0 * pn_gamma;
        v[4] += -10.0 * pn_sigma;
        v[3] += 188.0 * SL / 3.0 + 25.0 * dSigmaL;
    }
    for (int i = 0; i < 8; ++i) { v[i] *= pfaN; vl[i] *= pfaN; }
    /* [UPSTREAM] tidal terms: XLALSimInspiralTaylorF2Phasing_{10,12}PNTidalCoeff */
    const double m1M4 = m1M * m1M * m1M * m1M, m2M4 = m2M * m2M * m2M * m2M;
    const double pft10 = pfaN * (lambda1 * (-288.0 + 264.0 * m1M) * m1M4 + lambda2 * (-288.0 + 264.0 * m2M) * m2M4);
    const double pft12 = pfaN * (lambda1 * (-15895.0 / 28.0 + 4595.0 / 28.0 * m1M + 5715.0 / 14.0 * m1M * m1M - 325.0 / 7.0 * m1M * m1M * m1M) * m1M4
                               + lambda2 * (-15895.0 / 28.0 + 4595.0 / 28.0 * m2M + 5715.0 / 14.0 * m2M * m2M - 325.0 / 7.0 * m2M * m2M * m2M) * m2M4);
    /* ---- [UPSTREAM] XLALSimInspiralTaylorF2 ---- */
    const double m_sec = mtot * LAL_MTSUN_SI, piM = LAL_PI * m_sec;
    const double vISCO = 1.0 / sqrt(6.0);
    const double fISCO = vISCO * vISCO * vISCO / piM;
    if (fISCO <= m->f_lower_wf) return -INFINITY;       /* NoWaveformError -> -inf */
    long n = (long)(fISCO / m->delta_f + 1);
    const long iStart = (long)ceil(m->f_lower_wf / m->delta_f);
    const double r = distance * 1e6 * LAL_PC_SI;
    const double amp0 = -4.0 * m1 * m2 / r * LAL_MRSUN_SI * LAL_MTSUN_SI * sqrt(LAL_PI / 12.0);
    const double shft = LAL_TWOPI * (-1.0 / m->delta_f);
    const double FTaN = 32.0 * eta * eta / 5.0, dETaN = 2.0 * (-eta / 2.0);
    const long nuse = n < m->kmax ? n : m->kmax;   /* bins >= kmax are never read: do not generate */
    cplx* h = s->h;
    for (long k = 0; k < (iStart < nuse ? iStart : nuse); ++k) h[k] = 0;
    for (long k = iStart; k < nuse; ++k) {
        const double f = k * m->delta_f;
        const double vv = cbrt(piM * f), logv = log(vv);
        const double v2 = vv * vv, v3 = vv * v2, v4 = vv * v3, v5 = vv * v4, v6 = vv * v5, v7 = vv * v6;
        const double v10 = v5 * v5;
        double phasing = 0.0;
        phasing += v[7] * v7;
        phasing += (v[6] + vl[6] * logv) * v6;
        phasing += (v[5] + vl[5] * logv) * v5;
        phasing += v[4] * v4;
        phasing += v[3] * v3;
        phasing += v[2] * v2;
        phasing += v[1] * vv;
        phasing += v[0];
        phasing += pft12 * (v10 * v2);
        phasing += pft10 * v10;
        phasing /= v5;
        const double flux = FTaN * v10, dEnergy = dETaN * vv;
        phasing += shft * f - 2.0 * phi_ref;
        const double amp = amp0 * sqrt(-dEnergy / flux) * vv;
        h[k] = amp * cos(phasing - LAL_PI_4) - amp * sin(phasing - LAL_PI_4) * I;
    }
    /* ---- ChooseFDWaveform ---- */
    const double cfac = cos(incl), pfac = 0.5 * (1.0 + cfac * cfac);
    for (long k = 0; k < nuse; ++k) { s->hp[k] = pfac * h[k]; s->hc[k] = (-I * cfac) * h[k]; }
    /* ---- FDomainDetFrameGenerator.generate + gwin inner products ---- */
    const double gmst = gmst_lal(tc, m->leaps);
    const double gha = gmst - ra;
    const double cosgha = cos(gha), singha = sin(gha), cosdec = cos(dec), sindec = sin(dec);
    const double cospsi = cos(psi), sinpsi = sin(psi);
    const double X[3] = {-cospsi * singha - sinpsi * cosgha * sindec,
                         -cospsi * cosgha + sinpsi * singha * sindec, sinpsi * cosdec};
    const double Y[3] = {sinpsi * singha - cospsi * cosgha * sindec,
                         sinpsi * cosgha + cospsi * singha * sindec, cospsi * cosdec};
    const double eh[3] = {cosdec * cosgha, cosdec * -singha, sindec};
    double lr = 0, opt = 0;
    cplx mf = 0;
    for (int d = 0; d < nd; ++d) {
        double fp = 0, fc = 0;
        for (int i = 0; i < 3; ++i) {
            const double dx = m->resp[d][3 * i] * X[0] + m->resp[d][3 * i + 1] * X[1] + m->resp[d][3 * i + 2] * X[2];
            const double dy = m->resp[d][3 * i] * Y[0] + m->resp[d][3 * i + 1] * Y[1] + m->resp[d][3 * i + 2] * Y[2];
            fp += X[i] * dx - Y[i] * dy;
            fc += X[i] * dy + Y[i] * dx;
        }
        const double delay = ((0 - m->loc[d][0]) * eh[0] + (0 - m->loc[d][1]) * eh[1] + (0 - m->loc[d][2]) * eh[2]) / LAL_C_SI;
        cplx* hdt = s->hd;
        for (long k = 0; k < nuse; ++k) hdt[k] = fp * s->hp[k] + fc * s->hc[k];
        /* apply_fseries_time_shift */
        const double dt = (tc + delay) - m->epoch;
        const double phi = -2 * M_PI * dt * m->delta_f;
        const double cphi = cos(phi), sphi = sin(phi);
        for (long k0 = 0; k0 < nuse; k0 += 100) {
            double re = cos(phi * k0), im = sin(phi * k0);
            const long k1 = k0 + 100 < nuse ? k0 + 100 : nuse;
            for (long k = k0; k < k1; ++k) {
                hdt[k] *= re + I * im;
                const double t = re * cphi - im * sphi;
                im = re * sphi + im * cphi;
                re = t;
            }
        }
        /* gaussian_noise.py:328-340 */
        const long kmax = nuse;
        cplx hd = 0; double hh = 0;
        if (m->kmin < kmax) {
            for (long k = m->kmin; k < kmax; ++k) hdt[k] *= m->weight[d][k];
            cplx a = 0, b = 0;
            for (long k = m->kmin; k < kmax; ++k) a += conj(m->data[d][k]) * hdt[k];
            for (long k = m->kmin; k < kmax; ++k) b += conj(hdt[k]) * hdt[k];
            hd = a; hh = creal(b);
        }
        hd_out[d] = hd; hh_out[d] = hh;
        lr += creal(hd - 0.5 * hh);
        opt += hh; mf += hd;
    }
    if (mode == 1) return log_i0(cabs(mf)) - 0.5 * opt;
    return lr;
}

typedef struct {
    const oracle_model* m; int mode; long w0, w1; const double* params;
    double *loglr, *hd, *hh;
} job;

static void* worker(void* arg) {
    job* j = arg;
    const oracle_model* m = j->m;
    scratch s;
    s.h = malloc(sizeof(cplx) * m->n_f); s.hp = malloc(sizeof(cplx) * m->n_f);
    s.hc = malloc(sizeof(cplx) * m->n_f); s.hd = malloc(sizeof(cplx) * m->n_f);
    cplx hd[MAX_DET]; double hh[MAX_DET];
    for (long w = j->w0; w < j->w1; ++w) {
        j->loglr[w] = one_loglr(m, j->mode, j->params + 13 * w, &s, hd, hh);
        for (int d = 0; d < m->n_det; ++d) {
            if (j->hd) { j->hd[2 * (w * m->n_det + d)] = creal(hd[d]); j->hd[2 * (w * m->n_det + d) + 1] = cimag(hd[d]); }
            if (j->hh) j->hh[w * m->n_det + d] = hh[d];
        }
    }
    free(s.h); free(s.hp); free(s.hc); free(s.hd);
    return NULL;
}

/* params [W][13] in the order of include/gwin_b200.h; hd [W][n_det][2]; hh [W][n_det] */
int oracle_loglr(const oracle_model* m, int mode, long W, const double* params,
                 double* loglr, double* hd, double* hh, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > W) nthreads = W > 0 ? (int)W : 1;
    pthread_t* th = malloc(sizeof(pthread_t) * nthreads);
    job* jobs = malloc(sizeof(job) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (job){m, mode, W * t / nthreads, W * (t + 1) / nthreads, params, loglr, hd, hh};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
