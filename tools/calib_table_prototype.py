"""Prototype (numpy, CPU) for round 2: the calibration factor inside the moment-table kernel.

c_d(k) = (1 + dA)(2 + i dphi)/(2 - i dphi) = exp(L_d(k)), L_d smooth.  On a sub-block fit the sextic
L_d(kc + j) - L_d(kc) = sum_{k=1..6} lam_k (j/h)^k at the phase nodes.  Then
    T c e^{-2 pi i Theta} = c(kc) e^{-2 pi i th0} e^{-2 pi i nu' j} exp(U(s)) T,
    nu' = nu - Im(lam_1) / (2 pi h)            (the imaginary slope shifts the local time)
    U(s) = Re(lam_1) s + sum_{k>=2} (lam_k - 2 pi i t_k) s^k
so the Taylor coefficients e_n of exp(U) now have e_1 = Re(lam_1) != 0 and the sum needs M_1, which is
not tabulated: M_1(nu) = sum_j (j/h) T e^{-2 pi i nu j} = (i / (2 pi h)) dM_0/dnu, i.e. the same table
row of M_0 with the DERIVATIVE of the interpolation kernel as tap weights.
This script checks that against the direct calibrated sum."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import util
from kernel_model import PI, TWOPI


def es_kernel_deriv(x, w, beta):
    z = 2.0 * np.asarray(x, dtype=float) / w
    out = np.zeros_like(z)
    ok = np.abs(z) < 1
    r = np.sqrt(1 - z[ok] ** 2)
    out[ok] = np.exp(beta * (r - 1)) * beta * (-z[ok] / r) * (2.0 / w)
    return out


def table_block_cal(pm, w, b, d, coef):
    s, m = pm.tab_start[b], pm.tab_m[b]
    N, W, beta = pm.TAB_DEG, pm.TAB_W, 2.30 * pm.TAB_W
    h = m / 2.0
    nodes = np.array(pm.table_nodes(m))
    ic = pm.TAB_PH // 2
    sn = (nodes - m // 2) / h
    keep = [i for i in range(len(nodes)) if i != ic]
    V = np.array([[sn[i] ** k for k in range(1, pm.TAB_PH + 1)] for i in keep])
    th = pm.phase_turns(w, s + nodes)
    t = np.concatenate([[0.0], np.linalg.solve(V, (th - th[ic])[keep])])
    L = pm.calib_log(coef, d, s + nodes)
    lam = np.concatenate([[0.0], np.linalg.solve(V.astype(complex), (L - L[ic])[keep])])
    th0 = th[ic] + (s + m // 2) * w['tau'][d]
    nu = t[1] / h + w['tau'][d] - lam[1].imag / (TWOPI * h)
    u = lam - 2j * PI * t
    u[1] = lam[1].real
    c = np.zeros(N + 1, dtype=complex)
    c[0] = 1.0
    for n in range(1, N + 1):
        c[n] = sum(k * u[k] * c[n - k] for k in range(1, min(n, pm.TAB_PH) + 1)) / n
    x = (nu - math.floor(nu)) * (2 * m)
    g0 = int(math.ceil(x - W / 2.0))
    taps = g0 + np.arange(W)
    rows = pm.tab[b][d][:, taps % (2 * m)]
    M = (rows * pm.es_kernel(x - taps, W, beta)[None, :]).sum(axis=1)
    # M_1 from the M_0 row: d/dnu of sum_g A[g] phi(nu L - g) = L sum_g A[g] phi'(nu L - g)
    dM0 = (2 * m) * (rows[0] * es_kernel_deriv(x - taps, W, beta)).sum()
    M1 = 1j / (TWOPI * h) * dM0
    total = c[0] * M[0] + c[1] * M1 + np.dot(c[2:], M[2:])
    return np.exp(L[ic]) * np.exp(-2j * PI * (th0 - math.floor(th0))) * total, abs(M1 - M[1]) / max(abs(M[1]), 1e-300)


def main():
    import gwin_oracle as O
    import test_kernel_model as tk
    cfg = util.c3(seglen=32)
    pm = tk._plan(cfg, mchirp_min=1.0)
    pm.build_tables()
    n = 6
    rec = {d: O.CubicSpline(20., 2048., n, d.lower()) for d in cfg.dets}
    pm.set_calibration([rec[d].spline_points for d in cfg.dets])
    rng = np.random.default_rng(2)
    worst, worst_m1 = 0.0, 0.0
    for row in util.draw_bns(rng, 2, cfg):
        vals = rng.normal(0, 0.08, (len(cfg.dets), 2, n))
        coef = pm.calib_coef(vals)
        w = pm.walker(row)
        for d in range(pm.n_det):
            for b in range(len(pm.tab_m)):
                s, m = pm.tab_start[b], pm.tab_m[b]
                if s + m > w['ke']:
                    break
                k = np.arange(s, s + m)
                direct = np.sum(pm.T[d, k] * pm.calib_factor(coef, d, k)
                                * np.exp(-2j * PI * (pm.phase_turns(w, k) + k * w['tau'][d])))
                got, e1 = table_block_cal(pm, w, b, d, coef)
                norm = np.sqrt(np.sum(np.abs(pm.T[d, k]) ** 2))
                worst = max(worst, abs(got - direct) / norm)
                worst_m1 = max(worst_m1, e1)
    print("calibrated sub-block sums from the tables vs direct: worst |error| / rms %.2e; "
          "M_1 from the derivative kernel vs its table: worst relative %.2e" % (worst, worst_m1))


if __name__ == "__main__":
    main()
