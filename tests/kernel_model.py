"""numpy model of the CUDA kernels' arithmetic (test aid, runs without a GPU).

Mirrors ``gwin_b200/csrc/gwin_cuda.cu`` step by step -- table build, per-walker
coefficient folding, the direct per-bin evaluation and the block-anchored
quartic phasor recurrence with its block schedule -- so that the *algorithm*
(not the CUDA code generation) can be checked against the oracle on the CPU:
coefficient folding, node placement, forward differences, the phase-error bound
of the schedule.  It is not part of the product and is not an oracle.
"""
import math

import numpy as np

PI = math.pi
TWOPI = 2 * math.pi
GAMMA = 0.577215664901532860606512090082402431
C_SI = 299792458.0
MTSUN = 4.925491025543575903411922162094833998e-6
MRSUN = 1.476625061404649406193430731479084713e3
PC_SI = 3.085677581491367278913937957796471611e16
LEAPS = (46828800, 78364801, 109900802, 173059203, 252028804, 315187205,
         346723206, 394156807, 425692808, 457228809, 504489610, 551750411,
         599184012, 820108813, 914803214, 1025136015, 1119744016, 1167264017)
MAX_BLOCK, MIN_BLOCK = 256, 8


class PlanModel(object):
    def __init__(self, delta_f, epoch, f_lower, f_upper, f_lower_wf, resp, loc,
                 data, psd, norm=None, phase_tol=1e-10, mchirp_min=0.5):
        data = np.asarray(data, dtype=np.complex128)
        self.n_det, self.n_f = data.shape
        n_f = self.n_f
        N = (n_f - 1) * 2
        self.kmin = int(f_lower / delta_f) if f_lower else 1
        self.kmax = int((N + 1) / 2.)
        if f_upper:
            self.kmax = min(int(f_upper / delta_f), self.kmax)
        self.istart = max(int(math.ceil(f_lower_wf / delta_f)), 1)
        self.f_lower_wf = f_lower_wf
        self.delta_f, self.epoch = delta_f, epoch
        self.gps_utc = sum(1 for t in LEAPS if math.floor(epoch) >= t)
        self.resp = np.asarray(resp, dtype=float).reshape(self.n_det, 3, 3)
        self.loc = np.asarray(loc, dtype=float).reshape(self.n_det, 3)
        if norm is None:
            norm = 4.0 * delta_f
        ld = np.longdouble
        k = np.arange(n_f)
        f = k.astype(ld) * ld(delta_f)
        f[0] = 1.0
        x = np.cbrt(f)
        self.bx = x.astype(float)
        self.bxm5 = (1 / x ** 5).astype(float)
        self.blx = np.log(x).astype(float)
        band = (k >= self.kmin) & (k < self.kmax)
        x7 = x ** 7
        f73 = np.where(band, 1 / x7, 0)
        f76 = np.where(band, 1 / np.sqrt(x7), 0)
        self.T = np.zeros((self.n_det, n_f), dtype=np.complex128)
        self.C = np.zeros((self.n_det, n_f + 1))
        self.lognl_det = np.zeros(self.n_det)
        for d in range(self.n_det):
            w2 = np.where(band, ld(norm) / np.where(band, psd[d], 1).astype(ld)
                          if psd is not None else ld(norm), 0)
            dd = data[d]
            self.T[d] = (dd.real.astype(ld) * w2 * f76).astype(float) - \
                1j * (dd.imag.astype(ld) * w2 * f76).astype(float)
            self.C[d, 1:] = np.cumsum(f73 * w2).astype(float)
            self.lognl_det[d] = float(-0.5 * np.sum(
                (dd.real.astype(ld) ** 2 + dd.imag.astype(ld) ** 2) * w2))
            if d == 0:
                self._w73 = np.zeros((self.n_det, n_f), dtype=ld)
            self._w73[d] = f73 * w2
        self.phase_tol, self.mchirp_min = phase_tol, mchirp_min
        self.blk_start = self.build_schedule()
        self.cal = None

    # -- gwin_plan_set_calibration ----------------------------------------------
    def set_calibration(self, spline_freqs):
        """spline_freqs [n_det][n_points].  Per detector: z = (f - f_mid) / f_half maps the
        spline range to [-1, 1]; calM = the least-squares map from the n_points values to the
        cubic's monomial coefficients in z (what UnivariateSpline's default smoothing returns);
        Ccal[j] = prefix sums of w f^(-7/3) z^j, j = 0..6, for the closed-form <h|h>."""
        ld = np.longdouble
        sf = np.asarray(spline_freqs, dtype=float)
        n_f = self.n_f
        cal = {"zs": [], "z0": [], "M": [], "C": np.zeros((self.n_det, 7, n_f + 1))}
        for d in range(self.n_det):
            fm, fh = 0.5 * (sf[d].max() + sf[d].min()), 0.5 * (sf[d].max() - sf[d].min())
            zs, z0 = self.delta_f / fh, -fm / fh
            z = (sf[d].astype(ld) - ld(fm)) / ld(fh)
            V = np.stack([z ** 0, z, z ** 2, z ** 3], axis=1)
            Q, R = np.linalg.qr(V.astype(float))
            cal["M"].append(np.linalg.solve(R, Q.T))
            cal["zs"].append(zs)
            cal["z0"].append(z0)
            zk = np.arange(n_f).astype(ld) * ld(zs) + ld(z0)
            for j in range(7):
                cal["C"][d, j, 1:] = np.cumsum(self._w73[d] * zk ** j).astype(float)
        self.cal = cal

    def calib_coef(self, values):
        """values [n_det][2][n_points] -> cubic coefficients [n_det][2][4] (z^0..z^3)."""
        v = np.asarray(values, dtype=float)
        return np.array([[self.cal["M"][d].dot(v[d, w]) for w in range(2)]
                         for d in range(self.n_det)])

    def calib_factor(self, coef, d, k):
        """(1 + dA)(2 + i dphi)/(2 - i dphi) at bins k, as the direct kernel forms it."""
        z = np.asarray(k, dtype=float) * self.cal["zs"][d] + self.cal["z0"][d]
        a, ph = coef[d, 0], coef[d, 1]
        A = ((a[3] * z + a[2]) * z + a[1]) * z + a[0]
        P = ((ph[3] * z + ph[2]) * z + ph[1]) * z + ph[0]
        return (1 + A) / (4 + P * P) * ((4 - P * P) + 4j * P)

    def calib_log(self, coef, d, k):
        """log of the factor, as the recurrence kernel samples it at block nodes."""
        z = np.asarray(k, dtype=float) * self.cal["zs"][d] + self.cal["z0"][d]
        a, ph = coef[d, 0], coef[d, 1]
        A = ((a[3] * z + a[2]) * z + a[1]) * z + a[0]
        P = ((ph[3] * z + ph[2]) * z + ph[1]) * z + ph[0]
        return np.log1p(A) + 2j * np.arctan(0.5 * P)

    def calib_hh(self, w, coef):
        """closed-form <h|h> with the amplitude factor: (1 + A(z))^2 is a sextic in z."""
        hh = []
        for d in range(self.n_det):
            q = np.convolve(coef[d, 0] + np.array([1., 0, 0, 0]), coef[d, 0] + np.array([1., 0, 0, 0]))
            v = 0.0
            if w['ke'] > w['ks']:
                C = self.cal["C"][d]
                v = abs(w['A'][d]) ** 2 * float(np.dot(q, C[:, w['ke']] - C[:, w['ks']]))
            hh.append(v)
        return hh

    @staticmethod
    def block_mat(B, deg):
        """make_block_mat: nodes and the linear map samples -> forward differences."""
        if deg == 0:
            return None
        e = B - 1
        if deg == 4:
            q = (e + 4) // 8
            x = [0, q, e // 2, e - q, e]
        else:
            q = (e + 2) // 4
            x = [0, q, e - q, e]
        ld = np.longdouble
        m = np.zeros((4, 4))
        for col in range(1, deg + 1):
            dd = [ld(0)] * (deg + 1)
            dd[col] = ld(1)
            nw = [dd[0]]
            for lvl in range(1, deg + 1):
                for i in range(0, deg + 1 - lvl):
                    dd[i] = (dd[i + 1] - dd[i]) / ld(x[i + lvl] - x[i])
                nw.append(dd[0])
            c = [ld(0)] * 6
            basis = [ld(1)] + [ld(0)] * 5
            for lvl in range(deg + 1):
                for q_ in range(lvl + 1):
                    c[q_] += nw[lvl] * basis[q_]
                nb = [ld(0)] * 6
                for q_ in range(lvl + 1):
                    nb[q_ + 1] += basis[q_]
                    nb[q_] -= basis[q_] * ld(x[lvl])
                basis = nb
            c1, c2, c3 = c[1], c[2], c[3]
            c4 = c[4] if deg == 4 else ld(0)
            m[0, col - 1] = float(c1 + c2 + c3 + c4)
            m[1, col - 1] = float(2 * c2 + 6 * c3 + 14 * c4)
            m[2, col - 1] = float(6 * c3 + 36 * c4)
            m[3, col - 1] = float(24 * c4)
        return {"deg": deg, "node": x[1:], "m": m}

    def build_schedule(self):
        band0 = max(self.kmin, self.istart)
        K4 = (5. / 3) * (8. / 3) * (11. / 3) * (14. / 3)
        K5 = K4 * (17. / 3)
        chirp = 2.0 * (3. / 128.) * (PI * self.mchirp_min * MTSUN) ** (-5. / 3)
        setup_cost, bin4, bin3 = 450.0, 80.0, 71.0
        out, self.blk_deg = [], []
        k = max(band0, 1)
        while k < self.kmax:
            f = k * self.delta_f
            d5 = chirp * K5 * self.delta_f ** 5 * f ** (-20. / 3)
            d4 = chirp * K4 * self.delta_f ** 4 * f ** (-17. / 3)
            B4 = min((self.phase_tol * 120.0 / (0.005 * d5)) ** 0.2 + 1.0, MAX_BLOCK)
            B3 = min((self.phase_tol * 24.0 / (0.015625 * d4)) ** 0.25 + 1.0, MAX_BLOCK)
            deg, Bi = 4, int(B4)
            if B3 >= MIN_BLOCK and bin3 + setup_cost / B3 < bin4 + setup_cost / B4:
                deg, Bi = 3, int(B3)
            if Bi >= MIN_BLOCK:
                Bi &= ~3
            else:
                deg = 4
                for Bi in (7, 6, 5):
                    if Bi == 5:
                        break
                    e = Bi - 1
                    q = (e + 4) // 8
                    x = [0, q, e // 2, e - q, e]
                    worst = max(abs(np.prod([j - xi for xi in x])) for j in range(e + 1))
                    if d5 / 120.0 * worst <= self.phase_tol:
                        break
            if k + Bi > self.kmax:
                Bi = self.kmax - k
            if Bi < 5:
                deg = 0
            out.append(k)
            self.blk_deg.append(deg)
            k += Bi
        out.append(self.kmax)
        return out

    # -- walker_setup_kernel --------------------------------------------------
    def walker(self, p):
        m1, m2, chi1, chi2, dist, incl, phic, ra, dec, psi, tc, lam1, lam2 = [float(v) for v in p]
        mtot = m1 + m2
        dm = (m1 - m2) / (m1 + m2)
        eta = m1 * m2 / mtot / mtot
        m1M, m2M = m1 / mtot, m2 / mtot
        pfaN = 3.0 / (128.0 * eta)
        v2 = 5.0 * (743.0 / 84.0 + 11.0 * eta) / 9.0
        v3 = -16.0 * PI
        v4 = 5.0 * (3058.673 / 7.056 + 5429.0 / 7.0 * eta + 617.0 * eta * eta) / 72.0
        v5 = 5.0 / 9.0 * (7729.0 / 84.0 - 13.0 * eta) * PI
        vl5 = 5.0 / 3.0 * (7729.0 / 84.0 - 13.0 * eta) * PI
        v6 = (11583.231236531 / 4.694215680 - 640.0 / 3.0 * PI * PI - 6848.0 / 21.0 * GAMMA) \
            + eta * (-15737.765635 / 3.048192 + 2255.0 / 12.0 * PI * PI) \
            + eta * eta * 76055.0 / 1728.0 - eta * eta * eta * 127825.0 / 1296.0
        v6 += (-6848.0 / 21.0) * math.log(4.0)
        vl6 = -6848.0 / 21.0
        v7 = PI * (77096675.0 / 254016.0 + 378515.0 / 1512.0 * eta - 74045.0 / 756.0 * eta * eta)
        SL = m1M * m1M * chi1 + m2M * m2M * chi2
        dSigmaL = dm * (m2M * chi2 - m1M * chi1)
        pn_sigma = eta * (721.0 / 48.0 * chi1 * chi2 - 247.0 / 48.0 * chi1 * chi2)
        pn_sigma += 719.0 / 96.0 * m1M * m1M * chi1 * chi1
        pn_sigma += 719.0 / 96.0 * m2M * m2M * chi2 * chi2
        pn_sigma -= 233.0 / 96.0 * m1M * m1M * chi1 * chi1
        pn_sigma -= 233.0 / 96.0 * m2M * m2M * chi2 * chi2
        pn_ss3 = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * chi1 * chi2
        pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m1M - 120.0 * m1M * m1M)
                   + (-4108.25 / 6.72 - 108.5 / 1.2 * m1M + 125.5 / 3.6 * m1M * m1M)) * m1M * m1M * chi1 * chi1
        pn_ss3 += ((4703.5 / 8.4 + 2935.0 / 6.0 * m2M - 120.0 * m2M * m2M)
                   + (-4108.25 / 6.72 - 108.5 / 1.2 * m2M + 125.5 / 3.6 * m2M * m2M)) * m2M * m2M * chi2 * chi2
        pn_gamma = (554345.0 / 1134.0 + 110.0 * eta / 9.0) * SL + (13915.0 / 84.0 - 10.0 * eta / 3.0) * dSigmaL
        v7 += (-8980424995.0 / 762048.0 + 6586595.0 * eta / 756.0 - 305.0 * eta * eta / 36.0) * SL \
            - (170978035.0 / 48384.0 - 2876425.0 * eta / 672.0 - 4735.0 * eta * eta / 144.0) * dSigmaL
        v6 += PI * (3760.0 * SL + 1490.0 * dSigmaL) / 3.0 + pn_ss3
        v5 += -pn_gamma
        vl5 += -3.0 * pn_gamma
        v4 += -10.0 * pn_sigma
        v3 += 188.0 * SL / 3.0 + 25.0 * dSigmaL
        piM = PI * (mtot * MTSUN)
        m3 = piM ** (1. / 3)
        m3 = float(np.cbrt(piM))
        lm3 = math.log(m3)
        s = pfaN / TWOPI
        im3 = 1.0 / m3
        w = {}
        w['a0'] = s * im3 ** 5
        w['a2'] = s * v2 * im3 ** 3
        w['a3'] = s * v3 * im3 ** 2
        w['a4'] = s * v4 * im3
        w['a5'] = s * (v5 + vl5 * lm3) - (2.0 * phic + 0.25 * PI) / TWOPI
        w['b5'] = s * vl5
        w['a6'] = s * (v6 + vl6 * lm3) * m3
        w['b6'] = s * vl6 * m3
        w['a7'] = s * v7 * m3 * m3
        c10 = lambda x: (-288.0 + 264.0 * x) * x ** 4
        c12 = lambda x: (-15895.0 / 28.0 + 4595.0 / 28.0 * x + 5715.0 / 14.0 * x * x - 325.0 / 7.0 * x ** 3) * x ** 4
        w['a10'] = s * (lam1 * c10(m1M) + lam2 * c10(m2M)) * m3 ** 5
        w['a12'] = s * (lam1 * c12(m1M) + lam2 * c12(m2M)) * m3 ** 7
        vI = 1.0 / math.sqrt(6.0)
        fisco = vI * vI * vI / piM
        n = int(fisco / self.delta_f + 1.0)
        w['valid'] = fisco > self.f_lower_wf
        w['ke'] = min(n, self.kmax)
        w['ks'] = max(self.kmin, self.istart)
        r = dist * 1e6 * PC_SI
        amp0 = -4.0 * m1 * m2 / r * MRSUN * MTSUN * math.sqrt(PI / 12.0)
        ampfac = amp0 * math.sqrt(5.0 / (32.0 * eta)) * (im3 ** 3 / math.sqrt(m3))
        cfac = math.cos(incl)
        pfac = 0.5 * (1.0 + cfac * cfac)
        # XLALGreenwichMeanSiderealTime with LAL's integer-second Julian day
        sec = math.floor(tc)
        ns = round((tc - sec) * 1e9)
        if ns >= 1000000000:
            sec += 1
            ns -= 1000000000
        utc = int(sec) - int(self.gps_utc)
        days, sod = divmod(utc, 86400)
        jd = (2444245 + days) + sod / 86400.0 - 0.5
        t_hi = (jd - 2451545.0) / 36525.0
        t_lo = ns / (1e9 * 36525.0 * 86400.0)
        tj = t_hi + t_lo
        sid = (-6.2e-6 * tj + 0.093104) * tj * tj + 67310.54841
        sid += 8640184.812866 * t_lo
        sid += 3155760000.0 * t_lo
        sid += 8640184.812866 * t_hi
        sid += 3155760000.0 * t_hi
        gmst = sid * PI / 43200.0
        gha = gmst - ra
        sg, cg = math.sin(gha), math.cos(gha)
        sd, cd = math.sin(dec), math.cos(dec)
        sp, cp = math.sin(psi), math.cos(psi)
        X = np.array([-cp * sg - sp * cg * sd, -cp * cg + sp * sg * sd, sp * cd])
        Y = np.array([sp * sg - cp * cg * sd, sp * cg + cp * sg * sd, cp * cd])
        eh = np.array([cd * cg, -cd * sg, sd])
        w['tau'], w['A'], w['hh'] = [], [], []
        for d in range(self.n_det):
            dx, dy = self.resp[d].dot(X), self.resp[d].dot(Y)
            fp = (X * dx - Y * dy).sum()
            fc = (X * dy + Y * dx).sum()
            delay = -(self.loc[d].dot(eh)) / C_SI
            dt = (tc + delay) - self.epoch
            tau = dt * self.delta_f
            tau = tau - round(tau)
            A = complex(fp * pfac * ampfac, -fc * cfac * ampfac)
            w['tau'].append(tau)
            w['A'].append(A)
            hh = 0.0
            if w['ke'] > w['ks']:
                hh = abs(A) ** 2 * (self.C[d, w['ke']] - self.C[d, w['ks']])
            w['hh'].append(hh)
        return w

    def phase_turns(self, w, k):
        x, xm5, lx = self.bx[k], self.bxm5[k], self.blx[k]
        lo = (w['a4'] * x + w['a3']) * x + w['a2']
        lo = (lo * (x * x) + w['a0']) * xm5
        l5 = w['b5'] * lx + w['a5']
        l6 = w['b6'] * lx + w['a6']
        tid = (w['a12'] * (x * x) + w['a10']) * (x * x * x)
        return lo + (((w['a7'] + tid) * x + l6) * x + l5)

    def finish(self, w, S, marg=False):
        hd = np.array([w['A'][d] * S[d] for d in range(self.n_det)])
        hh = np.array(w['hh'])
        if marg:
            from scipy import special
            mf = abs(hd.sum())
            lr = mf + math.log(special.i0e(mf)) - 0.5 * hh.sum()
        else:
            lr = float((hd.real - 0.5 * hh).sum())
        return lr, hd, hh

    # -- loglr_direct_kernel ------------------------------------------------------
    def loglr_direct(self, p, marg=False, calib=None):
        w = self.walker(p)
        if not w['valid']:
            return -np.inf, None, None
        ks, ke = w['ks'], w['ke']
        S = np.zeros(self.n_det, dtype=complex)
        coef = None
        if calib is not None:
            coef = self.calib_coef(calib)
            w['hh'] = self.calib_hh(w, coef)
        if ke > ks:
            k = np.arange(ks, ke)
            psi = self.phase_turns(w, k)
            for d in range(self.n_det):
                ph = psi + np.modf(k * w['tau'][d])[0]
                ph = ph - np.rint(ph)
                u = np.exp(-2j * PI * ph)
                if coef is not None:
                    u = u * self.calib_factor(coef, d, k)
                S[d] = np.sum(self.T[d, ks:ke] * u)
        return self.finish(w, S, marg)

    # -- loglr_table_kernel -------------------------------------------------------
    # Above k_split the band is cut into sub-blocks of m bins (2^k or 3 2^(k-1)).  The TaylorF2
    # phase is a combination of 11 fixed functions of f, so its Taylor coefficients about a
    # sub-block's centre are a matrix (taylor_matrix) applied to the walker's phase coefficients.
    # Per sub-block there are K reference chirps (equal masses, no spin, a0 on a uniform grid over
    # the covered range) whose exact non-linear phase is baked into the tabulated moments
    #   M_n(nu) = sum_j (j/h)^n T[kc+j] e^{-2 pi i N_ref(j)} e^{-2 pi i nu j}      (grid nu = g/(2m))
    # and the block sum is e^{-2 pi i th0} sum_n e_n M_n(nu), the M_n interpolated by a TAB_W-tap
    # exponential-of-semicircle kernel, e_n the Taylor coefficients of e^{-2 pi i R},
    # R = sum_{n=2..8} (t_n - t_n^ref) s^n.  The dropped e_9, e_10 give the run-time estimate
    # that sends a walker outside the coverage to the slow path.
    TAB_DEG, TAB_NT, TAB_NTH, TAB_W, TAB_MIN, TAB_MAX = 8, 9, 13, 13, 128, 4096
    TAB_EST_DESIGN, TAB_EST_MAX, TAB_RHO_DESIGN, TAB_RHO_MAX = 2.0e-9, 1.6e-8, 0.05, 0.10
    BASIS = (("a0", -5, 0), ("a2", -3, 0), ("a3", -2, 0), ("a4", -1, 0), ("a5", 0, 0), ("b5", 0, 1),
             ("a6", 1, 0), ("b6", 1, 1), ("a7", 2, 0), ("a10", 5, 0), ("a12", 7, 0))

    @classmethod
    def table_sizes(cls):
        """Admissible sub-block lengths: 2^k and 3 2^(k-1)."""
        out, m = [], cls.TAB_MIN
        while m <= cls.TAB_MAX:
            out.append(m)
            if m + m // 2 <= cls.TAB_MAX:
                out.append(m + m // 2)
            m *= 2
        return sorted(out)

    def taylor_matrix(self, kc, h):
        """c[i][n]: basis function i at f = (kc + h s) delta_f as a series in s (long double)."""
        ld = np.longdouble
        n = self.TAB_NTH
        fc = ld(kc) * ld(self.delta_f)
        eps = ld(h) / ld(kc)
        ls = np.zeros(n, dtype=ld)
        ls[0] = np.log(fc) / 3
        for k in range(1, n):
            ls[k] = ld(-1) ** (k + 1) * eps ** k / k / 3
        out = np.zeros((len(self.BASIS), n), dtype=ld)
        for i, (_, pw, lg) in enumerate(self.BASIS):
            q = ld(pw) / 3
            bs = np.zeros(n, dtype=ld)
            bs[0] = fc ** q
            for k in range(n - 1):
                bs[k + 1] = bs[k] * (q - k) / (k + 1) * eps
            out[i] = np.convolve(bs, ls)[:n] if lg else bs
        return out

    def coef_vector(self, w):
        return np.array([w[name] for name, _, _ in self.BASIS])

    def ref_walker(self, a0):
        """The reference chirp of Newtonian coefficient a0: equal masses, no spins, no tides."""
        piM = (a0 * (64.0 * PI / 3.0)) ** -0.6
        m = 0.5 * piM / (PI * MTSUN)
        return self.walker([m, m, 0, 0, 1.0, 0, 0, 0, 0, 0, self.epoch, 0, 0])

    def cover_stats(self, rows):
        """cover_stats_kernel + cover_widen: (a_lo, a_hi, dev[11]) of a set of parameter rows."""
        a0, dev = [], np.zeros(len(self.BASIS))
        for p in rows:
            q = list(p)
            q[6] = 0.0
            w = self.walker(q)
            if not w['valid']:
                continue
            a = self.coef_vector(w)
            ar = self.coef_vector(self.ref_walker(a[0]))
            d = np.abs(a - ar)
            d[0] = d[4] = 0.0
            dev = np.maximum(dev, d)
            a0.append(a[0])
        lo, hi = min(a0), max(a0)
        pad = 0.05 * (hi - lo) + 2e-3 * hi
        return max(lo - pad, 0.25 * lo), hi + pad, 1.5 * dev

    def table_admissible(self, c, dev, da_half):
        n = self.TAB_NTH
        u = np.zeros(n)
        for k in range(2, n):
            u[k] = TWOPI * float(da_half * abs(c[0][k]) + np.dot(dev[1:], np.abs(c[1:, k]).astype(float)))
        e = np.zeros(n)
        e[0] = 1.0
        for m in range(2, n):
            e[m] = sum(k * u[k] * e[m - k] for k in range(2, m + 1)) / m
        return e[self.TAB_DEG + 1:].sum() <= self.TAB_EST_DESIGN and u[2:self.TAB_DEG + 1].sum() <= self.TAB_RHO_DESIGN

    def table_refs_needed(self, cover, k, m, kmax_refs=4096):
        a_lo, a_hi, dev = cover
        c = self.taylor_matrix(k + m // 2, 0.5 * m)
        half = 0.5 * (a_hi - a_lo)
        if self.table_admissible(c, dev, half):
            return 1
        if not self.table_admissible(c, dev, half / kmax_refs):
            return 0
        lo, hi = half / kmax_refs, half
        for _ in range(14):
            mid = math.sqrt(lo * hi)
            if self.table_admissible(c, dev, mid):
                lo = mid
            else:
                hi = mid
        return min(kmax_refs, int(math.ceil(half / lo)))

    def build_tables(self, cover, K0=64.0):
        """table_schedule + the lazily filled tables for the coverage (a_lo, a_hi, dev)."""
        self.cover = cover
        band0 = max(max(self.kmin, self.istart), 1)
        f0 = band0 * self.delta_f
        sizes = self.table_sizes()

        def pick(k):
            cap = max(1.0, K0 * (k * self.delta_f / f0) ** (-11.0 / 9.0))
            for m in reversed(sizes):
                if k + m > self.kmax:
                    continue
                K = self.table_refs_needed(cover, k, m)
                if 0 < K <= cap:
                    return m, K
            return 0, 0
        bs = list(self.blk_start)
        ib = 0
        while ib < len(bs) - 1 and pick(bs[ib])[0] == 0:
            k_next = bs[ib] + 64
            while ib < len(bs) - 1 and bs[ib] < k_next:
                ib += 1
        self.tblk = []
        self.k_split = self.k_tab_end = self.kmax
        if ib >= len(bs) - 1:
            return
        k = bs[ib]
        self.k_split = k
        while k < self.kmax:
            m, K = pick(k)
            if m == 0:
                break
            self.tblk.append(dict(start=k, m=m, K=K, da=(cover[1] - cover[0]) / K, c=self.taylor_matrix(k + m // 2, 0.5 * m).astype(float),
                                  tabs={}, tref={}))
            k += m
        self.k_tab_end = k
        self._hat = {}

    def _table(self, B, r, d):
        if (r, d) not in B['tabs']:
            m, L, s = B['m'], 2 * B['m'], B['start']
            if m not in self._hat:
                self._hat[m] = self.es_hat((np.arange(m) - m // 2) / float(L), self.TAB_W, 2.30 * self.TAB_W)
            wr = self.ref_walker(self.cover[0] + (r + 0.5) * B['da'])
            tr = self.coef_vector(wr).dot(B['c'])
            B['tref'][r] = tr
            jc = np.arange(m) - m // 2
            nref = (self.phase_turns(wr, s + m // 2 + jc) - tr[0]) - tr[1] * (jc / (m / 2.0))
            seg = self.T[d, s:s + m] * np.exp(-2j * PI * nref) / self._hat[m]
            g = np.arange(L)
            rows = np.zeros((self.TAB_DEG + 1, L), dtype=complex)
            for n in range(self.TAB_DEG + 1):
                rows[n] = np.fft.fft(seg * (jc / (m / 2.0)) ** n, L) * np.exp(2j * np.pi * ((g * (m // 2)) % L) / L)
            B['tabs'][(r, d)] = rows
        return B['tabs'][(r, d)]

    @staticmethod
    def es_kernel(x, w, beta):
        z = 2.0 * np.asarray(x, dtype=float) / w
        out = np.zeros_like(z)
        ok = np.abs(z) < 1
        out[ok] = np.exp(beta * (np.sqrt(1 - z[ok] ** 2) - 1))
        return out

    @classmethod
    def es_hat(cls, xi, w, beta, nq=2048):
        x = (np.arange(nq) + 0.5) / nq * w - w / 2
        ph = cls.es_kernel(x, w, beta)
        return (ph[None, :] * np.cos(2 * np.pi * np.asarray(xi)[:, None] * x[None, :])).sum(axis=1) * (w / nq)

    _dweights = None

    @classmethod
    def tab_weights(cls, v, derivative=False):
        """TAB_WEIGHTS / TAB_DWEIGHTS of csrc/tab_poly.inc evaluated as the kernel evaluates them
        (nested FMAs with the generated literal coefficients)."""
        import os
        import re
        if cls._dweights is None:
            src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "gwin_b200", "csrc", "tab_poly.inc")).read()
            progs = {}
            for macro in ("TAB_WEIGHTS", "TAB_DWEIGHTS"):
                body = src[src.index("#define %s(v, w)" % macro):]
                body = body[:body.index("} while (0)")]
                code = "\n".join(ln.strip().rstrip("\\").strip() for ln in body.split("\n")[1:])
                code = code.replace("const double ", "").replace("double tw_e, tw_o;", "")
                code = code.replace("(w)", "w").replace("(v)", "v")
                code = re.sub(r"(-?0x[0-9a-f.]+p[+-]?\d+)", lambda mm: 'float.fromhex("%s")' % mm.group(1), code)
                progs[macro] = [st.strip() for st in code.split(";") if st.strip()]
            cls._dweights = progs
        wts = [0.0] * cls.TAB_W
        env = {"fma": lambda a, b, c: a * b + c, "w": wts, "v": float(v)}
        for st in cls._dweights["TAB_DWEIGHTS" if derivative else "TAB_WEIGHTS"]:
            exec(st, env)
        return np.array(wts)

    def calib_series(self, coef, d, kc, h):
        """Taylor coefficients L_1..L_6 in s = (k - kc) / h of log c_d, c_d the calibration factor, as
        loglr_table_kernel<CALIB> forms them: the two cubics shifted to the centre, then log(1 + r) and
        atan by their differential equations (no transcendental function); and c_d(z_c)."""
        zs, z0 = self.cal["zs"][d], self.cal["z0"][d]
        a, ph = coef[d, 0], coef[d, 1]
        zc, ez = kc * zs + z0, h * zs
        A0 = ((a[3] * zc + a[2]) * zc + a[1]) * zc + a[0]
        P0 = ((ph[3] * zc + ph[2]) * zc + ph[1]) * zc + ph[0]
        ia = 1.0 / (1.0 + A0)
        rr = [0.0, ez * ((3 * a[3] * zc + 2 * a[2]) * zc + a[1]) * ia, ez * ez * (3 * a[3] * zc + a[2]) * ia,
              ez ** 3 * a[3] * ia]
        xx = [0.5 * P0, 0.5 * ez * ((3 * ph[3] * zc + 2 * ph[2]) * zc + ph[1]),
              0.5 * ez * ez * (3 * ph[3] * zc + ph[2]), 0.5 * ez ** 3 * ph[3]]
        qq = [1 + xx[0] ** 2, 2 * xx[0] * xx[1], xx[1] ** 2 + 2 * xx[0] * xx[2], 2 * (xx[0] * xx[3] + xx[1] * xx[2]),
              xx[2] ** 2 + 2 * xx[1] * xx[3], 2 * xx[2] * xx[3]]
        ff, gg = [0.0] * 7, [0.0] * 7
        for n in range(1, 7):
            sf = sum((n - k) * rr[k] * ff[n - k] for k in range(1, min(n, 4)))
            sg = sum((n - k) * qq[k] * gg[n - k] for k in range(1, min(n, 6)))
            ff[n] = (rr[n] if n <= 3 else 0.0) - sf / n
            gg[n] = ((xx[n] if n <= 3 else 0.0) - sg / n) / qq[0]
        L = np.array([ff[n] + 2j * gg[n] for n in range(7)])
        c0 = (1 + A0) / (4 + P0 * P0) * ((4 - P0 * P0) + 4j * P0)
        return L, c0, 1 + A0

    def table_block(self, w, b, d, coef=None):
        """One full sub-block for one detector: the kernel's arithmetic.  Sets ``last_est`` /
        ``last_rho`` (what the kernel compares with TAB_EST_MAX / TAB_RHO_MAX).  ``coef``: calibration
        cubics (calib_coef) -- the CALIB variant: per-detector series with the log-calibration's Taylor
        coefficients, first moment from the derivative of the interpolation weights."""
        B = self.tblk[b]
        s, m = B['start'], B['m']
        N, W, beta = self.TAB_DEG, self.TAB_W, 2.30 * self.TAB_W
        h, L, kc = m / 2.0, 2 * m, s + m // 2
        t = self.coef_vector(w).dot(B['c'])[:self.TAB_NT]
        r = min(max(int((w['a0'] - self.cover[0]) / B['da']), 0), B['K'] - 1)
        rows = self._table(B, r, d)
        dt = t - B['tref'][r][:self.TAB_NT]
        th0 = t[0] + kc * w['tau'][d]
        nu = t[1] / h + w['tau'][d]
        u = np.zeros(N + 1, dtype=complex)
        u[2:] = -2j * PI * dt[2:N + 1]
        c0, extra_est = 1.0, 0.0
        if coef is not None:
            Ls, c0, amp = self.calib_series(coef, d, kc, h)
            u[1:6] += Ls[1:6]
            extra_est = abs(Ls[6].real) + abs(Ls[6].imag)
            if not amp > 0.25:
                extra_est = np.inf
        e = np.zeros(N + 3, dtype=complex)
        e[0] = 1.0
        for n in range(1, N + 3):
            e[n] = sum(k * u[k] * e[n - k] for k in range(1, min(n, N) + 1)) / n
        eps = h / kc
        self.last_est = float(np.abs(e[N + 1:].real).sum() + np.abs(e[N + 1:].imag).sum()
                              + 1.2 * eps * abs(TWOPI * dt[N]) * (1 + 1.2 * eps) + extra_est)
        self.last_rho = TWOPI * float(np.sum(np.abs(dt[2:])))
        x = (nu - math.floor(nu)) * L
        g0 = int(math.ceil(x - W / 2.0))
        taps = g0 + np.arange(W)
        wt = self.es_kernel(x - taps, W, beta)
        M = (rows[:, taps % L] * wt[None, :]).sum(axis=1)
        if coef is not None:
            # the first moment is not tabulated (e_1 = 0 without calibration):
            # M_1 = (i / 2 pi h) dM_0/dnu = (2 i / pi) sum_i w_i' tab_0[g0 + i]
            dw = self.tab_weights(2.0 * (x - g0) - (W - 1), derivative=True)
            M[1] = 2j / PI * np.sum(rows[0, taps % L] * dw)
        return c0 * np.exp(-2j * PI * (th0 - math.floor(th0))) * np.dot(e[:N + 1], M)

    @staticmethod
    def fine_decomposition(length, pm, fmin):
        """loglr_table_kernel<FINE>: the fine sub-blocks (offset, size) a walker takes between the start
        of the sub-block of ``pm`` bins it ends in and its last bin, ``length`` bins further on: at the
        level of sub-blocks of ml = pm / 2, pm / 4, ... >= fmin bins, the one at offset (q - 1) ml when
        q = (length // fmin * fmin) // ml is odd; and the number of bins left after the last one."""
        taken = []
        stretch = length // fmin * fmin
        ml = pm // 2
        while ml >= fmin:
            q = stretch // ml
            if q & 1:
                taken.append(((q - 1) * ml, ml))
            ml //= 2
        return taken, length - stretch

    def loglr_table(self, p, marg=False, calib=None):
        """Returns (loglr, hd, hh) and sets ``slow`` when the walker would leave the table path."""
        w = self.walker(p)
        if not w['valid']:
            return -np.inf, None, None
        ks, ke = w['ks'], w['ke']
        coef = None
        if calib is not None:
            coef = self.calib_coef(calib)
            w['hh'] = self.calib_hh(w, coef)
        S = np.zeros(self.n_det, dtype=complex)
        self.slow = not (self.cover[0] <= w['a0'] <= self.cover[1])
        self.worst_est = 0.0

        def exact(k_lo, k_hi):
            k = np.arange(k_lo, k_hi)
            psi = self.phase_turns(w, k)
            for d in range(self.n_det):
                u = np.exp(-2j * PI * (psi + k * w['tau'][d]))
                if coef is not None:
                    u = u * self.calib_factor(coef, d, k)
                S[d] += np.sum(self.T[d, k] * u)
        # low band: any exact method (the recurrence kernel on the GPU)
        if min(ke, self.k_split) > ks:
            exact(ks, min(ke, self.k_split))
        done = self.k_split
        for b, B in enumerate(self.tblk):
            if B['start'] + B['m'] > ke:
                break
            for d in range(self.n_det):
                S[d] += self.table_block(w, b, d, coef)
                self.worst_est = max(self.worst_est, self.last_est)
                if self.last_est > self.TAB_EST_MAX or self.last_rho > self.TAB_RHO_MAX:
                    self.slow = True
            done = B['start'] + B['m']
        if ke > done:                   # the tail: any exact method (second recurrence pass)
            exact(done, ke)
        return self.finish(w, S, marg)

    # -- loglr_recur_kernel -------------------------------------------------------
    def loglr_recur(self, p, marg=False, return_phase_err=False, calib=None):
        w = self.walker(p)
        if not w['valid']:
            return -np.inf, None, None
        ks, ke = w['ks'], w['ke']
        coef = None
        if calib is not None:
            coef = self.calib_coef(calib)
            w['hh'] = self.calib_hh(w, coef)
        S = np.zeros(self.n_det, dtype=complex)
        max_err = 0.0
        bs = self.blk_start
        for b in range(len(bs) - 1):
            kb, kn = bs[b], bs[b + 1]
            if kn <= ks or kb >= ke:
                continue
            B = kn - kb
            j = np.arange(B)
            k = kb + j
            on = (k >= ks) & (k < ke)
            M = self.block_mat(B, self.blk_deg[b])
            if M is None:
                psi = self.phase_turns(w, k)
                for d in range(self.n_det):
                    ph = psi + k * w['tau'][d]
                    u = np.exp(-2j * PI * ph)
                    if coef is not None:
                        u = u * self.calib_factor(coef, d, k)
                    S[d] += np.sum(np.where(on, self.T[d, k], 0) * u)
                continue
            deg = M["deg"]
            f0 = self.phase_turns(w, kb)
            fn = self.phase_turns(w, kb + np.array(M["node"]))
            D = M["m"][:, :deg].dot(fn - f0)
            D1, D2, D3, D4 = D
            if return_phase_err:
                exact = self.phase_turns(w, k)
                # the polynomial the differences describe, by summation
                d1 = D1 + np.concatenate([[0.], np.cumsum(D2 + np.concatenate([[0.], np.cumsum(
                    D3 + D4 * np.arange(B - 2))])[:B - 1])])[:B - 1] if B > 2 else np.array([D1])
                poly = f0 + np.concatenate([[0.], np.cumsum(d1)])
                max_err = max(max_err, float(np.max(np.abs(exact - poly)[on])) * TWOPI
                              if on.any() else 0.0)
            for d in range(self.n_det):
                Tk = np.where(on, self.T[d, k], 0)
                E1, E2, E3, E4 = D1, D2, D3, D4
                cend = 1.0
                if coef is not None:
                    # the per-detector log-calibration joins the block polynomial as an
                    # imaginary phase: Theta' = Theta + i L / (2 pi)
                    L0 = self.calib_log(coef, d, kb)
                    Ln = self.calib_log(coef, d, kb + np.array(M["node"]))
                    DL = M["m"][:, :deg].dot(Ln - L0) * (1j / TWOPI)
                    E1, E2, E3, E4 = D1 + DL[0], D2 + DL[1], D3 + DL[2], D4 + DL[3]
                    cend = np.exp(Ln[-1])
                tq = np.exp(2j * PI * E4)
                g = np.exp(2j * PI * (E1 + w['tau'][d]))
                rq = np.exp(2j * PI * E2)
                sq = np.exp(2j * PI * E3)
                acc = Tk[0]
                for jj in range(1, B):
                    acc = acc * g + Tk[jj]
                    g = g * rq
                    rq = rq * sq
                    if deg == 4:
                        sq = sq * tq
                kend = kb + B - 1
                anchor = fn[-1] + np.modf(kend * w['tau'][d])[0]
                S[d] += acc * np.exp(-2j * PI * anchor) * cend
        out = self.finish(w, S, marg)
        if return_phase_err:
            return out + (max_err,)
        return out
