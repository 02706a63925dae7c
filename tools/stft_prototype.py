"""Prototype (numpy, CPU) of a table-driven <d|h>: per sub-block moments of the data looked up
by NUFFT-style interpolation in the chirp's local time, instead of a per-bin phasor recurrence.

For a sub-block of m bins centred on kc the detector phase is Theta(kc + j) = th0 + nu j + R(j),
R = th2 j^2 + th3 j^3 + th4 j^4 (the block quartic the recurrence kernel already builds).  With
e^{-2 pi i R(j)} = sum_n c_n (j/h)^n, h = m/2, truncated at degree N,

    sum_j T[kc+j] e^{-2 pi i Theta} = e^{-2 pi i th0} sum_n c_n M_n(nu),
    M_n(nu) = sum_j (j/h)^n T[kc+j] e^{-2 pi i nu j}

M_n is a trigonometric polynomial of nu (period 1): tabulate it once per plan on the grid
nu = g / (os m) by a zero-padded FFT of the kernel-deconvolved sequence, and evaluate it per walker
with a w-tap "exponential of semicircle" kernel (the type-2 NUFFT).  Per (walker, sub-block,
detector) the work is O(w (N + 1)), independent of m.

This script measures the error of that scheme against the direct sum for a C3-like segment so
that (m, N, w, os) can be chosen; see DESIGN.md section 9.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import util
from kernel_model import PlanModel


def es_kernel(x, w, beta):
    """exp(beta (sqrt(1 - (2x/w)^2) - 1)) on |x| <= w/2."""
    z = 2.0 * x / w
    out = np.zeros_like(x)
    ok = np.abs(z) <= 1
    out[ok] = np.exp(beta * (np.sqrt(1 - z[ok] ** 2) - 1))
    return out


def es_hat(xi, w, beta, nq=4000):
    """Fourier transform of the kernel at frequencies xi (cycles per grid unit), by quadrature."""
    x = (np.arange(nq) + 0.5) / nq * w - w / 2
    ph = es_kernel(x, w, beta)
    return (ph[None, :] * np.cos(2 * np.pi * xi[:, None] * x[None, :])).sum(axis=1) * (w / nq)


class StftPlan(object):
    def __init__(self, T, k0, k1, m, N, w=12, osf=2):
        """T: [n_f] complex table of one detector; sub-blocks of m bins covering [k0, k1)."""
        self.m, self.N, self.w, self.os = m, N, w, osf
        self.beta = 2.30 * w
        self.L = osf * m
        self.h = m / 2.0
        self.starts = np.arange(k0, k1, m)
        jc = np.arange(m) - m // 2          # integer offsets: M_n(nu) has period 1
        what = es_hat(jc / self.L, w, self.beta)
        self.tab = np.zeros((len(self.starts), N + 1, self.L), dtype=complex)
        g = np.arange(self.L)
        for b, s in enumerate(self.starts):
            seg = np.zeros(m, dtype=complex)
            n_have = min(m, k1 - s)
            seg[:n_have] = T[s:s + n_have]
            for n in range(N + 1):
                a = seg * (jc / self.h) ** n / what
                # A[g] = sum_j a_j e^{-2 pi i jc_j g / L}
                A = np.fft.fft(a, self.L) * np.exp(2j * np.pi * g * (m // 2) / self.L)
                self.tab[b, n] = A

    def moments(self, b, nu):
        """M_n(nu), n = 0..N, by w-tap interpolation."""
        x = (nu % 1.0) * self.L
        g0 = int(np.ceil(x - self.w / 2.0))
        taps = g0 + np.arange(self.w)
        wt = es_kernel(x - taps, self.w, self.beta)
        return (self.tab[b][:, taps % self.L] * wt[None, :]).sum(axis=1)


def block_sum_table(sp, b, th, N):
    """th = (th0, nu, th2, th3, th4) about the block centre [turns]."""
    th0, nu, th2, th3, th4 = th
    h = sp.h
    # e^{u R}, R = r2 s^2 + r3 s^3 + r4 s^4 in s = j/h, as a polynomial in s truncated at degree N
    r = np.zeros(N + 1, dtype=complex)
    for n, c in ((2, th2 * h ** 2), (3, th3 * h ** 3), (4, th4 * h ** 4)):
        if n <= N:
            r[n] = -2j * np.pi * c
    c = np.zeros(N + 1, dtype=complex)
    c[0] = 1.0
    term = c.copy()
    for kk in range(1, N // 2 + 1):
        term = np.convolve(term, r)[:N + 1] / kk
        c = c + term
    M = sp.moments(b, nu)
    return np.exp(-2j * np.pi * th0) * np.dot(c, M)


def main():
    seglen = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    cfg = util.c3(seglen=seglen)
    import test_kernel_model as tk
    pm = tk._plan(cfg, mchirp_min=1.0)
    rng = np.random.default_rng(5)
    P = util.draw_bns(rng, 3, cfg)
    d = 0
    fband = [(100., 200.), (200., 400.), (400., 800.), (800., 1500.)]
    print("T=%d s; errors are |table - direct| / sqrt(sum |T|^2) per band, worst of %d walkers" % (seglen, len(P)))
    for m in (64, 128, 256):
        for N in (4, 6, 8):
            for w in (10, 13):
                line = []
                for f0, f1 in fband:
                    k0, k1 = int(f0 / cfg.delta_f), int(f1 / cfg.delta_f)
                    k1 = k0 + ((k1 - k0) // m) * m
                    sp = StftPlan(pm.T[d], k0, k1, m, N, w=w)
                    worst = 0.0
                    for row in P:
                        wk = pm.walker(row)
                        tau = wk['tau'][d]
                        tot_t, tot_d = 0j, 0j
                        for b, s in enumerate(sp.starts):
                            k = s + np.arange(m)
                            th_exact = pm.phase_turns(wk, k) + k * tau
                            direct = np.sum(pm.T[d, k] * np.exp(-2j * np.pi * th_exact))
                            # quartic through 5 nodes about the centre (as the block polynomial)
                            jc = k - (s + m // 2)
                            nodes = np.array([0, m // 8, m // 2, m - 1 - m // 8, m - 1])
                            coef = np.polynomial.polynomial.polyfit(jc[nodes] / sp.h, th_exact[nodes] - th_exact[m // 2], 4)
                            th = (coef[0] + th_exact[m // 2], coef[1] / sp.h, coef[2] / sp.h ** 2,
                                  coef[3] / sp.h ** 3, coef[4] / sp.h ** 4)
                            tot_t += block_sum_table(sp, b, th, N)
                            tot_d += direct
                        norm = np.sqrt(np.sum(np.abs(pm.T[d, k0:k1]) ** 2))
                        worst = max(worst, abs(tot_t - tot_d) / norm)
                    line.append("%.1e" % worst)
                print("m=%3d N=%d w=%2d : %s" % (m, N, w, "  ".join(line)))


if __name__ == "__main__":
    main()
