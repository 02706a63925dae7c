"""Prototype (numpy, CPU) for round 2: moment tables RELATIVE TO A REFERENCE CHIRP.

The table kernel needs |2 pi R| <= rho at the edge of a sub-block, R the non-linear
part of the walker's phase.  If the tables are built from T[k] e^{-2 pi i Q_b(j)} with Q_b the
non-linear phase of a REFERENCE waveform (the middle of the batch's chirp-mass range, known on the
host), only the difference R - Q_b has to be small: sub-blocks grow as 1/sqrt(relative width of the
range) and k_split moves down.  This script measures, for a C3-like segment, the schedule and the
error of that scheme against the direct sum, for batches of decreasing width around a reference."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import util
from kernel_model import PlanModel, PI


class RefPlan(PlanModel):
    """PlanModel whose table schedule uses the residual curvature and whose tables carry the
    reference walker's non-linear phase."""
    width = 1.0          # relative half-width of the batch's leading-order curvature around the reference

    def _table_m(self, k, chirp, sizes):
        f = k * self.delta_f
        K2 = (5. / 3) * (8. / 3)
        d2 = chirp * K2 * self.delta_f ** 2 * f ** (-5. / 3 - 2) * self.width
        n = self.TAB_PH + 1
        K = 1.0
        for i in range(n):
            K *= (5. / 3 + i)
        dn = chirp * K * self.delta_f ** n * f ** (-5. / 3 - n) * self.width
        best = 0
        for m in sizes:
            if 0.5 * d2 * (0.5 * m) ** 2 <= self.TAB_RHO and dn / math.factorial(n) * self._tab_omega[m] <= self.phase_tol:
                best = m
        return best

    def build_tables_ref(self, ref_row):
        """Tables of T[k] e^{-2 pi i N_ref(j)}, N_ref the EXACT non-linear phase of the reference
        waveform on the sub-block (its value at the centre and an endpoint slope removed)."""
        self.ref_walker = self.walker(ref_row)
        self.build_table_schedule()
        w, beta, N = self.TAB_W, 2.30 * self.TAB_W, self.TAB_DEG
        self.tab, self.nref_nodes, hat = [], [], {}
        for b, (s, m) in enumerate(zip(self.tab_start[:-1], self.tab_m)):
            L = 2 * m
            if m not in hat:
                hat[m] = self.es_hat((np.arange(m) - m // 2) / float(L), w, beta)
            jc = np.arange(m) - m // 2
            sj = jc / (m / 2.0)
            th = self.phase_turns(self.ref_walker, s + np.arange(m))
            nu_ref = (th[-1] - th[0]) / (m - 1)
            nref = th - th[m // 2] - nu_ref * jc
            self.nref_nodes.append(nref[np.array(self.table_nodes(m))])
            base = np.exp(-2j * PI * nref)
            g = np.arange(L)
            rows = np.zeros((self.n_det, N + 1, L), dtype=complex)
            for d in range(self.n_det):
                seg = self.T[d, s:s + m] * base
                for n in range(N + 1):
                    rows[d, n] = np.fft.fft(seg * sj ** n / hat[m], L) * np.exp(2j * np.pi * ((g * (m // 2)) % L) / L)
            self.tab.append(rows)

    def table_block(self, w, b, d):
        s, m = self.tab_start[b], self.tab_m[b]
        N, W, beta = self.TAB_DEG, self.TAB_W, 2.30 * self.TAB_W
        h = m / 2.0
        nodes = np.array(self.table_nodes(m))
        ic = self.TAB_PH // 2
        sn = (nodes - m // 2) / h
        th = self.phase_turns(w, s + nodes)
        v = (th - th[ic]) - self.nref_nodes[b]          # the walker's phase relative to the reference
        keep = [i for i in range(len(nodes)) if i != ic]
        V = np.array([[sn[i] ** k for k in range(1, self.TAB_PH + 1)] for i in keep])
        t = np.concatenate([[0.0], np.linalg.solve(V, v[keep])])
        th0 = th[ic] + (s + m // 2) * w['tau'][d]
        nu = t[1] / h + w['tau'][d]
        u = -2j * PI * t
        u[1] = 0.0
        c = np.zeros(N + 1, dtype=complex)
        c[0] = 1.0
        for n in range(1, N + 1):
            c[n] = sum(k * u[k] * c[n - k] for k in range(2, min(n, self.TAB_PH) + 1)) / n
        x = (nu - math.floor(nu)) * (2 * m)
        g0 = int(math.ceil(x - W / 2.0))
        taps = g0 + np.arange(W)
        wt = self.es_kernel(x - taps, W, beta)
        M = (self.tab[b][d][:, taps % (2 * m)] * wt[None, :]).sum(axis=1)
        return np.exp(-2j * PI * (th0 - math.floor(th0))) * np.dot(c, M)


def main():
    seglen = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfg = util.c3(seglen=seglen)
    import test_kernel_model as tk
    base = tk._plan(cfg, mchirp_min=1.0)
    rng = np.random.default_rng(7)
    ref = util.inj_row(cfg)
    mc_ref = (ref[0] * ref[1]) ** 0.6 / (ref[0] + ref[1]) ** 0.2
    print("T = %d s, reference chirp mass %.4f; width = relative half-width of Mc^(-5/3) in the batch" % (seglen, mc_ref))
    for width in (1.0, 0.25, 0.05, 0.01):
        pm = RefPlan.__new__(RefPlan)
        pm.__dict__.update(base.__dict__)
        pm.width = width * 1.3 + 0.02                  # slack for the higher PN orders (eta, spins)
        pm.mchirp_min = mc_ref * (1 + width) ** (-0.6)
        pm.build_tables_ref(ref)
        errs = []
        for _ in range(3):
            row = ref.copy()
            f = (1 + width * rng.uniform(-1, 1)) ** (-0.6)       # chirp mass scaled inside the range
            row[0] *= f
            row[1] *= f
            row[10] += rng.uniform(-0.05, 0.05)
            a, b = pm.loglr_direct(row), pm.loglr_table(row)
            errs.append(max(abs(a[0] - b[0]), np.max(np.abs(a[1] - b[1]))))
        m = np.array(pm.tab_m)
        print("width %5.2f: k_split %6.1f Hz, %3d sub-blocks, mean %5.0f bins, low band %6d bins; max error %.1e"
              % (width, pm.k_split * cfg.delta_f, len(m), m.mean() if len(m) else 0, pm.k_split - max(pm.kmin, pm.istart), max(errs)))


if __name__ == "__main__":
    main()
