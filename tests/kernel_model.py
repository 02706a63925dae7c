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
    # Above k_split the band is cut into sub-blocks of m (a power of two) bins.  Per sub-block
    # the moments M_n(nu) = sum_j (j/h)^n T[kc+j] e^{-2 pi i nu j} are tabulated on the grid
    # nu = g/(2m) (zero-padded FFT of the kernel-deconvolved sequence); per walker the block sum
    # is e^{-2 pi i th0} sum_n c_n M_n(nu) with the M_n interpolated by a TAB_W-tap
    # exponential-of-semicircle kernel and c_n the Taylor coefficients of e^{-2 pi i R(j)},
    # R = the quadratic..sextic part of the polynomial through TAB_PH + 1 exact phase samples.
    TAB_DEG, TAB_W, TAB_MIN, TAB_MAX, TAB_RHO, TAB_PH = 8, 13, 128, 4096, 0.025, 6

    @classmethod
    def table_nodes(cls, m):
        """TAB_PH + 1 nodes of the phase polynomial on [0, m-1]: rounded Chebyshev-Lobatto points
        with the centre moved to m/2."""
        n = cls.TAB_PH
        x = [int(round((m - 1) * 0.5 * (1 - math.cos(math.pi * i / n)))) for i in range(n + 1)]
        x[n // 2] = m // 2
        return x

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

    def build_table_schedule(self):
        band0 = max(self.kmin, self.istart)
        # with a reference waveform (set_reference) the bounds apply to the batch's deviation from it
        ref = getattr(self, "ref", None)
        wfac = min(1.0, 1.3 * ref[1] + 0.02) if ref else 1.0
        chirp = wfac * 2.0 * (3. / 128.) * (PI * self.mchirp_min * MTSUN) ** (-5. / 3)
        sizes = self.table_sizes()
        # worst |prod (j - x_i)| over the sub-block, per size (interpolation error factor)
        self._tab_omega = {}
        for m in sizes:
            x = np.array(self.table_nodes(m), dtype=float)
            j = np.arange(m, dtype=float)
            self._tab_omega[m] = float(np.max(np.abs(np.prod(j[:, None] - x[None, :], axis=1))))
        kk = max(band0, 1)
        while kk < self.kmax and self._table_m(kk, chirp, sizes) < self.TAB_MIN:
            kk += 64
        self.k_split = min(kk, self.kmax)
        k = self.k_split
        self.tab_start, self.tab_m = [], []
        while k < self.kmax:
            m = max(self._table_m(k, chirp, sizes), self.TAB_MIN)
            while m > self.TAB_MIN and k + m > self.kmax:
                m = max(s for s in sizes if s < m)
            if k + m > self.kmax:           # sub-blocks lie wholly inside the band
                break
            self.tab_start.append(k)
            self.tab_m.append(m)
            k += m
        self.tab_start.append(k)
        return self.tab_start

    def _table_m(self, k, chirp, sizes):
        """Longest admissible sub-block starting at bin k: |R| <= TAB_RHO at its edge and the
        degree-TAB_PH interpolation of the phase good to phase_tol."""
        f = k * self.delta_f
        K = 1.0
        for i in range(2):
            K *= (5. / 3 + i)
        d2 = chirp * K * self.delta_f ** 2 * f ** (-5. / 3 - 2)
        n = self.TAB_PH + 1
        K = 1.0
        for i in range(n):
            K *= (5. / 3 + i)
        dn = chirp * K * self.delta_f ** n * f ** (-5. / 3 - n)
        best = 0
        for m in sizes:
            if 0.5 * d2 * (0.5 * m) ** 2 <= self.TAB_RHO and \
                    dn / math.factorial(n) * self._tab_omega[m] <= self.phase_tol:
                best = m
        return best

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

    def set_reference(self, ref_row=None, rel_width=0.0):
        """gwin_plan_set_reference: tables relative to the waveform of ``ref_row``; ``rel_width``
        bounds |(Mc/Mc_ref)^(-5/3) - 1| over the batches to come."""
        self.ref = (np.asarray(ref_row, dtype=float), float(rel_width)) if ref_row is not None else None

    def build_tables(self):
        self.build_table_schedule()
        w, beta, N = self.TAB_W, 2.30 * self.TAB_W, self.TAB_DEG
        self.tab = []
        self.tab_nref = []
        ref = getattr(self, "ref", None)
        ref_w = self.walker(ref[0]) if ref else None
        hat = {}
        for b, (s, m) in enumerate(zip(self.tab_start[:-1], self.tab_m)):
            L = 2 * m
            if m not in hat:
                hat[m] = self.es_hat((np.arange(m) - m // 2) / float(L), w, beta)
            jc = np.arange(m) - m // 2
            g = np.arange(L)
            rows = np.zeros((self.n_det, N + 1, L), dtype=complex)
            base = 1.0
            if ref_w is not None:
                # tab_ref_kernel: the reference's exact phase minus its centre value and an end-point
                # slope; the tables hold T e^{-2 pi i N_ref}
                th = self.phase_turns(ref_w, s + np.arange(m))
                nref = th - th[m // 2] - (th[-1] - th[0]) / (m - 1) * jc
                self.tab_nref.append(nref[np.array(self.table_nodes(m))])
                base = np.exp(-2j * PI * nref)
            for d in range(self.n_det):
                seg = self.T[d, s:s + m] * base
                for n in range(N + 1):
                    a = seg * (jc / (m / 2.0)) ** n / hat[m]
                    rows[d, n] = np.fft.fft(a, L) * np.exp(2j * np.pi * ((g * (m // 2)) % L) / L)
            self.tab.append(rows)

    def table_block(self, w, b, d):
        """One full sub-block for one detector: the kernel's arithmetic."""
        s, m = self.tab_start[b], self.tab_m[b]
        N, W, beta = self.TAB_DEG, self.TAB_W, 2.30 * self.TAB_W
        h = m / 2.0
        nodes = np.array(self.table_nodes(m))
        ic = self.TAB_PH // 2
        sn = (nodes - m // 2) / h
        th = self.phase_turns(w, s + nodes)
        keep = [i for i in range(len(nodes)) if i != ic]
        V = np.array([[sn[i] ** k for k in range(1, self.TAB_PH + 1)] for i in keep])
        v = th - th[ic]
        if getattr(self, "ref", None):
            v = v - self.tab_nref[b]                    # phase relative to the reference waveform
        t = np.concatenate([[0.0], np.linalg.solve(V, v[keep])])    # t_k s^k, t_0 = 0
        self.last_rho = TWOPI * float(np.sum(np.abs(t[2:])))
        th0 = th[ic] + (s + m // 2) * w['tau'][d]
        nu = t[1] / h + w['tau'][d]
        # Taylor coefficients of exp(U), U = -2 pi i sum_{k>=2} t_k s^k:  n e_n = sum_k k u_k e_{n-k}
        u = -2j * PI * t
        c = np.zeros(N + 1, dtype=complex)
        c[0] = 1.0
        for n in range(1, N + 1):
            acc = 0j
            for k in range(2, min(n, self.TAB_PH) + 1):
                acc += k * u[k] * c[n - k]
            c[n] = acc / n
        x = (nu - math.floor(nu)) * (2 * m)
        g0 = int(math.ceil(x - W / 2.0))
        taps = g0 + np.arange(W)
        wt = self.es_kernel(x - taps, W, beta)
        M = (self.tab[b][d][:, taps % (2 * m)] * wt[None, :]).sum(axis=1)
        return np.exp(-2j * PI * (th0 - math.floor(th0))) * np.dot(c, M)

    def loglr_table(self, p, marg=False):
        w = self.walker(p)
        if not w['valid']:
            return -np.inf, None, None
        ks, ke = w['ks'], w['ke']
        S = np.zeros(self.n_det, dtype=complex)
        # low band: any exact method (the recurrence kernel on the GPU)
        k_hi = min(ke, self.k_split)
        if k_hi > ks:
            k = np.arange(ks, k_hi)
            psi = self.phase_turns(w, k)
            for d in range(self.n_det):
                S[d] += np.sum(self.T[d, k] * np.exp(-2j * PI * (psi + k * w['tau'][d])))
        done = self.k_split
        for b, (s, m) in enumerate(zip(self.tab_start[:-1], self.tab_m)):
            if s + m > ke:
                break
            for d in range(self.n_det):
                S[d] += self.table_block(w, b, d)
            done = s + m
        if ke > done:                   # the tail: any exact method (second recurrence pass)
            k = np.arange(done, ke)
            psi = self.phase_turns(w, k)
            for d in range(self.n_det):
                S[d] += np.sum(self.T[d, k] * np.exp(-2j * PI * (psi + k * w['tau'][d])))
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
