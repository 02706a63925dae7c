"""numpy study of the round-2 moment tables: analytic Taylor coefficients of the phase about each
sub-block centre, a one-parameter family of reference chirps per sub-block, run-time truncation
estimate.  Not product code; checks the arithmetic the CUDA kernel uses.

  python tools/multiref_prototype.py [seglen]
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_model as KM  # noqa: E402

LD = np.longdouble
NT = 11          # Taylor orders kept per sub-block: s^0 .. s^10
DEG = 8          # moments n = 0, 2..DEG
TAB_W = 13
BETA = 2.30 * TAB_W
# phase basis of phase_turns(): coefficient name -> (power of x = f^(1/3), has ln x)
BASIS = [("a0", -5, 0), ("a2", -3, 0), ("a3", -2, 0), ("a4", -1, 0), ("a5", 0, 0), ("b5", 0, 1),
         ("a6", 1, 0), ("b6", 1, 1), ("a7", 2, 0), ("a10", 5, 0), ("a12", 7, 0)]


def binom_series(q, eps, n):
    """(1 + eps s)^q = sum_k out[k] s^k."""
    out = np.zeros(n, dtype=LD)
    c = LD(1)
    for k in range(n):
        out[k] = c * LD(eps) ** k
        c = c * (LD(q) - k) / (k + 1)
    return out


def log_series(eps, n):
    out = np.zeros(n, dtype=LD)
    for k in range(1, n):
        out[k] = LD(-1) ** (k + 1) * LD(eps) ** k / k
    return out


def taylor_matrix(kc, h, delta_f, n=NT):
    """c[i][k]: basis function i at f = (kc + h s) delta_f as a series in s."""
    fc = LD(kc) * LD(delta_f)
    eps = LD(h) / LD(kc)
    lnx = np.log(fc) / 3
    out = np.zeros((len(BASIS), n), dtype=LD)
    for i, (_, p, lg) in enumerate(BASIS):
        ser = fc ** (LD(p) / 3) * binom_series(LD(p) / 3, eps, n)
        if lg:
            l = log_series(eps, n) / 3
            l[0] = lnx
            ser = np.convolve(ser, l)[:n]
        out[i] = ser
    return out.astype(float)


def coef_vector(w):
    return np.array([w[name] for name, _, _ in BASIS])


def ref_walker(pm, a0):
    """Equal-mass, non-spinning, point-particle waveform with Newtonian coefficient a0."""
    s = 3.0 / (64.0 * math.pi)
    M = (a0 / s) ** (-3.0 / 5.0) / (math.pi * KM.MTSUN)
    p = [M / 2, M / 2, 0, 0, 1.0, 0, 0, 0, 0, 0, pm.epoch, 0, 0]
    return pm.walker(p)


def es_kernel(x):
    z = 2.0 * np.asarray(x, dtype=float) / TAB_W
    out = np.zeros_like(z)
    ok = np.abs(z) < 1
    out[ok] = np.exp(BETA * (np.sqrt(1 - z[ok] ** 2) - 1))
    return out


def es_hat(xi, nq=2048):
    x = (np.arange(nq) + 0.5) / nq * TAB_W - TAB_W / 2
    ph = es_kernel(x)
    return (ph[None, :] * np.cos(2 * np.pi * np.asarray(xi)[:, None] * x[None, :])).sum(axis=1) * (TAB_W / nq)


class SubBlock(object):
    def __init__(self, pm, s, m, a_refs):
        self.pm, self.s, self.m = pm, s, m
        self.kc, self.h = s + m // 2, m / 2.0
        self.c = taylor_matrix(self.kc, self.h, pm.delta_f)
        self.a_refs = np.asarray(a_refs)
        self.hat = es_hat((np.arange(m) - m // 2) / float(2 * m))
        self.tabs = {}
        self.tref = {}

    def table(self, r, d):
        if (r, d) in self.tabs:
            return self.tabs[(r, d)]
        pm, m, L = self.pm, self.m, 2 * self.m
        wr = ref_walker(pm, self.a_refs[r])
        tr = coef_vector(wr).dot(self.c)
        self.tref[r] = tr
        jc = np.arange(m) - m // 2
        th = pm.phase_turns(wr, self.kc + jc)
        nref = (th - tr[0]) - tr[1] * (jc / self.h)
        seg = pm.T[d, self.kc + jc] * np.exp(-2j * np.pi * nref) / self.hat
        g = np.arange(L)
        rows = np.zeros((DEG + 1, L), dtype=complex)
        for n in range(DEG + 1):
            rows[n] = np.fft.fft(seg * (jc / self.h) ** n, L) * np.exp(2j * np.pi * ((g * (m // 2)) % L) / L)
        self.tabs[(r, d)] = rows
        return rows

    def evaluate(self, w, d, r=None):
        """(table value, direct sum, run-time truncation estimate, rho)."""
        pm, m, L = self.pm, self.m, 2 * self.m
        a = coef_vector(w)
        t = a.dot(self.c)
        if r is None:
            r = int(np.argmin(np.abs(self.a_refs - w['a0'])))
        rows = self.table(r, d)
        dt = t - self.tref[r]
        tau = w['tau'][d]
        th0 = t[0] + self.kc * tau
        nu = t[1] / self.h + tau
        u = -2j * np.pi * dt
        e = np.zeros(NT, dtype=complex)
        e[0] = 1.0
        for n in range(2, NT):
            acc = 0j
            for k in range(2, n + 1):
                acc += k * u[k] * e[n - k]
            e[n] = acc / n
        x = (nu - math.floor(nu)) * L
        g0 = int(math.ceil(x - TAB_W / 2.0))
        taps = g0 + np.arange(TAB_W)
        wt = es_kernel(x - taps)
        M = (rows[:, taps % L] * wt[None, :]).sum(axis=1)
        val = np.exp(-2j * np.pi * (th0 - math.floor(th0))) * np.dot(e[:DEG + 1], M)
        jc = np.arange(m) - m // 2
        k = self.kc + jc
        ph = pm.phase_turns(w, k) + k * tau
        direct = np.sum(pm.T[d, k] * np.exp(-2j * np.pi * (ph - np.floor(ph))))
        scale = np.sum(np.abs(pm.T[d, k]))
        return val, direct, float(np.sum(np.abs(e[DEG + 1:]))), 2 * np.pi * float(np.sum(np.abs(dt[2:]))), scale


def make_plan(seglen, srate=4096.0, f_lower=20.0):
    n_f = int(seglen * srate) // 2 + 1
    rng = np.random.default_rng(1)
    data = (rng.standard_normal((1, n_f)) + 1j * rng.standard_normal((1, n_f)))
    psd = np.ones((1, n_f))
    resp = np.array([[-0.3926141, -0.0776130, -0.2473886, -0.0776130, 0.3195244, 0.2279981,
                      -0.2473886, 0.2279981, 0.0730903]])
    loc = np.array([[-2.16141492636e6, -3.83469517889e6, 4.60035022664e6]])
    return KM.PlanModel(1.0 / seglen, 0.0, f_lower, None, f_lower, resp, loc, data, psd)


def draw(rng, W, tc0, spin=0.0, lam=0.0, mlo=1.2, mhi=1.6):
    P = np.zeros((W, 13))
    P[:, 0] = rng.uniform(mlo, mhi, W)
    P[:, 1] = rng.uniform(mlo, mhi, W)
    P[:, 2] = rng.uniform(-spin, spin, W)
    P[:, 3] = rng.uniform(-spin, spin, W)
    P[:, 4] = 40.0
    P[:, 5:10] = rng.uniform(0, 1, (W, 5))
    P[:, 10] = tc0 + rng.uniform(-0.05, 0.05, W)
    P[:, 11] = rng.uniform(0, lam, W)
    P[:, 12] = rng.uniform(0, lam, W)
    return P


def main():
    seglen = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    pm = make_plan(seglen)
    rng = np.random.default_rng(2)
    P = draw(rng, 24, seglen - 6.0)
    ws = [pm.walker(p) for p in P]
    a0 = np.array([w['a0'] for w in ws])
    a_lo, a_hi = a0.min() * 0.999, a0.max() * 1.001
    print("a0 range %.6g .. %.6g (rel %.3f)" % (a_lo, a_hi, (a_hi - a_lo) / a_lo))
    for f0, m in ((20.0, 128), (20.0, 256), (40.0, 512), (100.0, 1024), (150.0, 4096), (400.0, 4096), (1000.0, 4096)):
        s = int(f0 * seglen)
        if s + m > pm.kmax:
            continue
        kc, h = s + m // 2, m / 2.0
        c = taylor_matrix(kc, h, pm.delta_f)
        # reference spacing: |delta t_2| 2 pi <= rho_a for |delta a| <= spacing / 2
        rho_a = 0.02
        spacing = 2 * rho_a / (2 * np.pi * abs(c[0, 2]))
        K = max(1, int(math.ceil((a_hi - a_lo) / spacing)))
        a_refs = a_lo + (np.arange(K) + 0.5) * (a_hi - a_lo) / K
        sb = SubBlock(pm, s, m, a_refs)
        worst, worst_est, worst_rho = 0.0, 0.0, 0.0
        for w in ws:
            if w['ke'] < s + m:
                continue
            val, direct, est, rho, scale = sb.evaluate(w, 0)
            err = abs(val - direct) / scale
            worst, worst_est, worst_rho = max(worst, err), max(worst_est, est), max(worst_rho, rho)
        print("f=%6.1f Hz m=%5d eps=%.4f  K=%5d  max rel err %.2e  run-time estimate %.2e  rho %.4f"
              % (f0, m, h / kc, K, worst, worst_est, worst_rho))


if __name__ == "__main__":
    main()
