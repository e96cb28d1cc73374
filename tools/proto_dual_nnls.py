"""Design prototype (CPU, numpy): Lawson-Hanson for the Tikhonov problem with a DIAGONAL regularisation matrix
(reg_matrix I and InvT2) carried out in ECHO SPACE (m = nTE = 32 unknowns) instead of T2 space (p <= 60 unknowns).

    min_{x >= 0} |D x - b|^2 + lam |L x|^2,   L = diag(l),   Dt = D diag(1/l),   xt = l * x  (same sign pattern)

For a positive set P the stationary point is, by the push-through identity,

    v    = (lam I_m + M_P)^-1 b,        M_P = Dt_P Dt_P^T = sum_{j in P} dt_j dt_j^T     (m x m, independent of lam)
    zt_P = Dt_P^T v                     (the least-squares coefficients on P)
    w_Z  = lam Dt_Z^T v                 (the dual vector on the zero set: r = b - Dt_P zt_P = lam v exactly)

so ONE product g = Dt^T v (n x m) gives both the coefficients (on P) and the dual (on Z); a column entering or leaving
P is a rank-one change of M_P; a new lambda (next Brent abscissa) keeps M_P and costs one m x m factorisation
(m^3/3 = 11 k flops) instead of a p x p one (p^3/3 = 42 k at p = 50) — and the per-voxel factor is tri(32) = 528
doubles (4.2 KB) instead of tri(60) = 1830 (14.6 KB), which is what pins the T2 kernel at 10 warps per SM today
(DESIGN.md §8 item 1).  With lam > 0 both acceptance tests of nnls.f always pass (a regularised column is never
dependent and its entering coefficient w_j / (lam (1 + dt_j^T A^-1 dt_j)) is positive), so the control flow is the
main loop + the interpolation loop only.

This script measures what decides whether the round-2 kernel may be built this way: supports / lambda / spectra of the
echo-space solver inside the X2 search (SciPy's fminbound, as the reference) against voxels fitted by the UNMODIFIED
reference (tests/golden/config2_subset.npz: FA spline + X2-I), and against the oracle for InvT2.

    python tools/proto_dual_nnls.py [n_voxels] [I|InvT2] [fresh|update|rec|sm] [warm]
"""
import os
import sys
import time

import numpy as np
from scipy.optimize import fminbound

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import met2_oracle as O  # noqa: E402


class EchoFactor:
    """A = lam I + M_P as an upper-triangular inverse factor T, A^-1 = T T^T.  `fresh`: refactor from M_P at every
    solve; `update`: rank-one update / downdate of T when a column enters / leaves (dense algebra); `rec`: the same update
    as the O(m) per-row recurrence the kernel would run; `sm`: explicit inverse (Gauss-Jordan
    without pivoting at every new lambda, Sherman-Morrison when a column enters / leaves)."""

    def __init__(self, m, lam, mode):
        self.m, self.lam, self.mode = m, lam, mode
        self.M = np.zeros((m, m))
        self.T = np.eye(m) / np.sqrt(lam)
        self.Ainv = np.eye(m) / lam
        self.n_fact = self.n_up = 0

    def _refactor(self):
        A = self.M + self.lam * np.eye(self.m)
        # A = R^T R  (R upper) ->  A^-1 = R^-1 R^-T = T T^T with T = R^-1 upper
        R = np.linalg.cholesky(A).T
        self.T = np.linalg.solve(R, np.eye(self.m))
        self.n_fact += 1

    def _gauss_jordan(self):
        """Explicit inverse by Gauss-Jordan without pivoting (SPD), the arithmetic a warp with one matrix row per lane
        would do: 32 steps of `row_i -= (a_ik / a_kk) row_k`."""
        m = self.m
        W = np.concatenate((self.M + self.lam * np.eye(m), np.eye(m)), axis=1)
        for k in range(m):
            W[k] = W[k] / W[k, k]
            f = W[:, k].copy()
            f[k] = 0.0
            W -= np.outer(f, W[k])
        self.Ainv = W[:, m:]
        self.n_fact += 1

    def change(self, d, sign):
        self.M += sign * np.outer(d, d)
        self.n_up += 1
        if self.mode == "fresh":
            self._refactor()
            return
        if self.mode == "sm":
            # Sherman-Morrison on the explicit inverse: (A + s d d^T)^-1 = A^-1 - s u u^T / (1 + s d.u),  u = A^-1 d
            u = self.Ainv @ d
            self.Ainv = self.Ainv - (sign / (1.0 + sign * (d @ u))) * np.outer(u, u)
            return
        # A' = A + sign d d^T = T^-T (I + sign u u^T) T^-1,  u = T^T d   ->   A'^-1 = T (I + sign u u^T)^-1 T^T
        u = self.T.T @ d
        if self.mode == "rec":
            # The form the kernel uses: (I + s u u^T)^-1 = Q Q^T with Q upper triangular, Q_jj = delta_j and
            # Q_ij = u_i q_j (i < j).  With the inclusive prefix sums tau_k = sum_{j <= k} u_j^2 (one warp scan):
            #     h_k = 1 + s tau_k,   delta_k^2 = h_{k-1} / h_k,   q_k = -s u_k / (h_k delta_k)
            # (s = +1 entering column: h >= 1, no cancellation; s = -1 leaving column: h_k >= 1 - u.u > 0).
            # T' = T Q is a running sum per ROW of T (one row per lane):  T'[r][j] = delta_j T[r][j] + q_j acc,
            # acc += T[r][j] u_j.
            tau = np.cumsum(u * u)
            h = 1.0 + sign * tau
            hm1 = np.concatenate(([1.0], h[:-1]))
            delta = np.sqrt(hm1 / h)
            q = -sign * u / (h * delta)
            Tn = np.zeros_like(self.T)
            for r in range(self.m):
                acc = 0.0
                for j in range(r, self.m):          # row r of an upper triangular T starts at its diagonal
                    t = self.T[r, j]
                    Tn[r, j] = delta[j] * t + q[j] * acc
                    acc += t * u[j]
            self.T = Tn
            return
        C = np.eye(self.m) + sign * np.outer(u, u)
        # (I + s u u^T)^-1 = Q Q^T with Q upper triangular (Cholesky of the inverse, "UL" form): T' = T Q stays upper
        Ci = np.linalg.inv(C)
        J = np.eye(self.m)[::-1]
        Q = J @ np.linalg.cholesky(J @ Ci @ J) @ J          # upper triangular, Q Q^T = Ci
        self.T = self.T @ Q

    def solve(self, b):
        if self.mode == "sm":
            return self.Ainv @ b
        return self.T @ (self.T.T @ b)


def nnls_echo(Dt, b, lam, mode="fresh", start=None, stats=None):
    """Lawson-Hanson (nnls.f control flow: pivot = first arg-max of w over Z, interpolation with the first arg-min of
    x_i / (x_i - z_i), every x_i <= 0 leaves, itmax = 3 n) with the echo-space linear algebra above.
    start: optional support (list of columns) with feasible coefficients xt[start] > 0 — warm start."""
    m, n = Dt.shape
    F = EchoFactor(m, lam, mode)
    x = np.zeros(n)
    inP = np.zeros(n, bool)
    it, itmax = 0, 3 * n
    secondary_first = False
    if start is not None and len(start[0]):
        cols, xs = start
        inP[cols] = True
        x[cols] = xs
        F.M = Dt[:, cols] @ Dt[:, cols].T
        if mode == "sm":
            F._gauss_jordan()
        else:
            F._refactor()
        secondary_first = True
    while True:
        if not secondary_first:
            if inP.sum() >= n:
                break
            g = Dt.T @ F.solve(b)
            w = np.where(inP, -np.inf, lam * g)
            j = int(np.argmax(w))
            if not (w[j] > 0):
                break
            inP[j] = True
            F.change(Dt[:, j], +1.0)
            if stats is not None:
                stats["outer"] = stats.get("outer", 0) + 1
        secondary_first = False
        stop = False
        while True:
            it += 1
            if it > itmax:
                stop = True
                break
            g = Dt.T @ F.solve(b)
            z = np.where(inP, g, 0.0)
            bad = inP & (z <= 0.0)
            if not bad.any():
                x = z
                break
            ratio = np.where(bad, x / np.where(bad, x - z, 1.0), np.inf)
            jb = int(np.argmin(ratio))
            alpha = ratio[jb]
            x = np.where(inP, x + alpha * (z - x), 0.0)
            out = inP & ((x <= 0.0) | (np.arange(n) == jb))
            for k in np.nonzero(out)[0]:
                inP[k] = False
                x[k] = 0.0
                F.change(Dt[:, k], -1.0)
                if stats is not None:
                    stats["removed"] = stats.get("removed", 0) + 1
        if stop:
            break
    if stats is not None:
        stats["fact"] = stats.get("fact", 0) + F.n_fact
        stats["solves"] = stats.get("solves", 0) + 1
    return x


def x2_echo(D, M, l, factor, mode, warm, stats):
    """nnls_x2 (algorithms.py:211-233) with the Tikhonov solves in echo space; plain NNLS (SSE0) from SciPy."""
    f0, _ = O.nnls(D, M)
    SSE = np.sum((D @ f0 - M) ** 2)
    Dt = D / l[None, :]
    last = [None]

    def solve(lam):
        xt = nnls_echo(Dt, M, lam, mode, start=last[0] if warm else None, stats=stats)
        if warm:
            cols = np.nonzero(xt > 0)[0]
            last[0] = (cols, xt[cols])
        return xt / l

    def obj(lam):
        f = solve(lam)
        return np.abs(np.sum((D @ f - M) ** 2) - factor * SSE) / SSE

    with np.errstate(all="ignore"):
        reg = fminbound(obj, 0.0, 10.0, xtol=1e-5, maxfun=300, full_output=0, disp=0)
        f = solve(reg)
    return f, reg, np.sum((D @ f - M) ** 2) / SSE


def main():
    nvox = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    matrix = sys.argv[2] if len(sys.argv) > 2 else "I"
    mode = sys.argv[3] if len(sys.argv) > 3 else "fresh"
    warm = (len(sys.argv) > 4 and sys.argv[4] == "warm")
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "config2_subset.npz")))
    gr = O._grids("X2", matrix, "spline", 40.0, 32, 10.0, 1000.0)
    Dic = O.create_Dic_3D(60, gr["T2s"], gr["T1s"], 32, 10.0, gr["alpha_values"], 1000.0)
    l = np.diag(gr["L"]).copy()
    sup_ref = np.unpackbits(g["support"], axis=1)[:, :60].astype(bool)
    fref = np.zeros(sup_ref.shape)
    fref[sup_ref] = g["f_nz"]
    step = max(1, len(g["sig"]) // nvox)
    sel = np.arange(0, len(g["sig"]), step)[:nvox]
    stats = {}
    n_sup = 0
    worst_f = worst_k = worst_mwf = 0.0
    t0 = time.time()
    for v in sel:
        M = g["sig"][v]
        if not (M[0] > 0):
            continue
        D = np.ascontiguousarray(Dic[:, :, int(g["fa_idx"][v])])
        f, reg, k_est = x2_echo(D, M / M[0], l, 1.02, mode, warm, stats)
        f = f * M[0]
        if matrix == "I":
            fr, kr = fref[v], g["reg"][v]
        else:
            fr, _sig, kr = O.t2_fit_voxel(M, D, "X2", gr["L"], gr["lambda_reg"])
        same = np.array_equal(f > 0, fr > 0)
        n_sup += not same
        worst_f = max(worst_f, np.max(np.abs(f - fr)) / np.max(np.abs(fr)))
        worst_k = max(worst_k, abs(k_est - kr) / abs(kr))
        mwf = f[gr["ind_m"]].sum() / f.sum()
        mwr = fr[gr["ind_m"]].sum() / fr.sum()
        worst_mwf = max(worst_mwf, abs(mwf - mwr))
    print("X2-%s echo-space (%s, %s start): %d voxels, %.0f s" % (matrix, mode, "warm" if warm else "cold", len(sel),
                                                                  time.time() - t0))
    print("  support disagreements vs %s: %d" % ("the unmodified reference" if matrix == "I" else "the oracle", n_sup))
    print("  max rel spectrum diff %.2e   max rel k_est diff %.2e   max |dMWF| %.2e" % (worst_f, worst_k, worst_mwf))
    nv = len(sel)
    print("  per voxel: %.1f solves, %.1f entering columns, %.1f removals, %.1f m x m factorisations" % (
        stats["solves"] / nv, stats.get("outer", 0) / nv, stats.get("removed", 0) / nv, stats["fact"] / nv))


if __name__ == "__main__":
    main()
