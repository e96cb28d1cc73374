"""TEST INFRASTRUCTURE ONLY — CPU oracle for the per-voxel MET2 inverse-problem path.

A restatement (numpy + the same SciPy routines the reference calls) of the reference's algorithm, function by
function, with the reference file:line each one follows.  It exists so that the CUDA path can be checked on a box
that has no /root/reference (the GPU box) and so that a CPU baseline can be timed beside the GPU number.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module;
the product package never does.

PINNING.  The reference ships no tests, golden vectors or seeds — "parity unpinned" by the
reference's own fixtures.  The oracle is instead pinned against OUTPUTS OF THE REFERENCE ITSELF, imported read-only
in the build container through `oracle/ref_shim.py`: `oracle/make_golden*.py` write those outputs to
`tests/golden/*.npz` — per-function vectors (make_golden.py), 20 480 + 2 048 voxels of the config-2 phantom
(make_golden_config2.py, make_golden_methods.py) and END-TO-END runs of the reference's two orchestrators
(make_golden_pipeline.py: motor_recon_met2 with NESMA + smoothing; make_golden_roi.py: motor_recon_met2_ROIs) — and
`tests/test_oracle_golden.py`, `tests/test_oracle_vs_reference.py`, `tests/test_oracle_pipeline_golden.py` hold the
oracle to them (bit-for-bit for indices/supports and <= 1e-9 .. 1e-12 for values) on any box.

Third-party arithmetic not under /root/reference (reference pins scipy==1.5.2, numpy==1.19.2 in requirements.txt:6,9;
this image has scipy 1.18.1 / numpy 2.3.5 — the operative oracle):
  * Lawson-Hanson NNLS: `scipy.optimize._slsqplib.nnls` (C translation of nnls.f).  `lh_nnls` below is our own
    restatement of the published algorithm (Lawson & Hanson 1974, ch. 23) used for flop counting and as an
    independent check.
  * bounded Brent: `scipy.optimize.fminbound` / `minimize_scalar(method='Bounded')`; `brent_bounded` below restates
    it and is tested for identical iterates.
  * `scipy.interpolate.interp1d(kind='cubic')`, `scipy.linalg.cholesky`, `scipy.linalg.det`, `scipy.special.erf`,
    `numpy.linalg.lstsq`.
"""
import math

import numpy as np
from scipy.interpolate import interp1d
from scipy.linalg import cholesky, det
from scipy.optimize import _slsqplib, fminbound, minimize_scalar
from scipy.special import erf

EPSILON = 1.0e-16

# ----------------------------------------------------------------------------------------------------------------
# EPG dictionary — epg/epg.py:47-162
# ----------------------------------------------------------------------------------------------------------------


def _epg_matrices(n, alpha):
    """Dense S (shift) and T (RF) of size (3n+1)^2 — epg/epg.py:97-131."""
    size = 3 * n + 1
    S = np.zeros((size, size))
    S[0, 2] = 1.0          # F0   <- F-1
    S[1, 0] = 1.0          # F+1  <- F0
    if size > 5:
        S[2, 5] = 1.0      # F-1  <- F-2
    S[3, 3] = 1.0          # Z1
    for k in range(2, n + 1):
        S[3 * k - 2, 3 * k - 5] = 1.0            # F+k <- F+(k-1)
        if 3 * k + 2 < size:
            S[3 * k - 1, 3 * k + 2] = 1.0        # F-k <- F-(k+1)
        S[3 * k, 3 * k] = 1.0                    # Zk
    c2 = math.cos(alpha / 2.0) ** 2
    s2 = math.sin(alpha / 2.0) ** 2
    sa = math.sin(alpha)
    ca = math.cos(alpha)
    T0 = np.array([[c2, s2, sa], [s2, c2, -sa], [-0.5 * sa, 0.5 * sa, ca]])
    T = np.zeros((size, size))
    T[0, 0] = 1.0
    for k in range(n):
        T[3 * k + 1:3 * k + 4, 3 * k + 1:3 * k + 4] = T0
    return S, T


def epg_signal(n, tau, R1, R2, alpha, alpha_exc):
    """One decay curve [n] — epg/epg.py:64-95,143-153 (dense E = P T P, x <- E x per echo, record x[0])."""
    half = tau / 2.0
    S, T = _epg_matrices(n, alpha)
    size = 3 * n + 1
    r = np.empty(size)
    r[0] = np.exp(-half * R2)
    r[1::3] = np.exp(-half * R2)
    r[2::3] = np.exp(-half * R2)
    r[3::3] = np.exp(-half * R1)
    P = np.dot(np.diag(r), S)
    E = np.dot(np.dot(P, T), P)
    x = np.zeros((size, 1))
    x[0] = math.sin(alpha_exc)
    x[2] = math.cos(alpha_exc)
    out = np.empty(n)
    for i in range(n):
        x = np.dot(E, x)
        out[i] = x[0, 0]
    return out


def create_met2_design_matrix_epg(Npc, T2s, T1s, nEchoes, tau, flip_angle, TR):
    """[nEchoes, Npc] for one refocusing angle (deg) — epg/epg.py:47-62."""
    rad = np.pi / 180.0
    D = np.zeros((nEchoes, Npc))
    for c in range(Npc):
        D[:, c] = (1.0 - np.exp(-TR / T1s[c])) * epg_signal(nEchoes, tau, 1.0 / T1s[c], 1.0 / T2s[c],
                                                            flip_angle * rad, flip_angle / 2.0 * rad)
    return D


def create_Dic_3D(Npc, T2s, T1s, nEchoes, tau, alpha_values, TR):
    """Dic_3D[nEchoes, Npc, nAlpha] — epg/epg.py:155-162."""
    Dic = np.zeros((nEchoes, Npc, len(alpha_values)))
    for i, a in enumerate(alpha_values):
        Dic[:, :, i] = create_met2_design_matrix_epg(Npc, T2s, T1s, nEchoes, tau, a, TR)
    return Dic


# ----------------------------------------------------------------------------------------------------------------
# NNLS — intravoxel_algorithms/algorithms.py:55-82 (duplicate bayesian_interpolation.py:46-73)
# ----------------------------------------------------------------------------------------------------------------


def nnls(A, b):
    """Lawson-Hanson NNLS with itmax = 3n and no error on itmax — algorithms.py:55-82.

    `asarray_chkfinite` raises ValueError on NaN/Inf exactly like the reference (:56).
    """
    A = np.asarray_chkfinite(A)
    b = np.asarray_chkfinite(b)
    n = A.shape[1]
    x, rnorm, _info = _slsqplib.nnls(np.ascontiguousarray(A, dtype=np.float64),
                                     np.ascontiguousarray(b, dtype=np.float64), 3 * n)
    return x, rnorm


class FlopCounter:
    """Reference-formulation flop model of SURVEY.md §8d (2 flops per multiply-add)."""

    def __init__(self):
        self.flops = 0
        self.solves = 0
        self.outer = 0
        self.removals = 0


def lh_nnls(A, b, itmax=None, counter=None):
    """Own restatement of Lawson & Hanson's NNLS (the algorithm behind scipy's nnls.f / C port).

    Works on copies of A and b that are progressively QR-transformed by Householder reflections (column entering the
    positive set) and Givens rotations (column leaving it).  Returns (x, rnorm, mode) with mode 1 = converged,
    3 = itmax reached (current x returned).
    """
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    b = np.array(b, dtype=np.float64, copy=True)
    m, n = A.shape
    if itmax is None:
        itmax = 3 * n
    x = np.zeros(n)
    w = np.zeros(n)
    zz = np.zeros(m)
    inP = np.zeros(n, dtype=bool)
    order = []            # columns of the positive set in triangularisation order
    it = 0
    mode = 1
    c = counter
    factor = 0.01
    while True:
        npp = len(order)
        if npp >= n or npp >= m:
            break
        Z = np.nonzero(~inP)[0]
        w[:] = 0.0
        w[Z] = A[npp:, Z].T @ b[npp:]
        if c:
            c.flops += 2 * (m - npp) * len(Z)
        accepted = False
        while True:
            wz = w[Z]
            k = int(np.argmax(wz))
            if not (wz[k] > 0.0):
                break
            j = int(Z[k])
            # Householder that would zero A[npp+1:, j]
            col = A[npp:, j].copy()
            asave = col[0]
            cl = np.max(np.abs(col))
            if cl <= 0.0:
                w[j] = 0.0
                continue
            sm = np.sqrt(np.sum((col / cl) ** 2)) * cl
            if col[0] > 0.0:
                sm = -sm
            up = col[0] - sm
            a_new = sm
            if c:
                c.flops += 3 * (m - npp)
            unorm = np.sqrt(np.sum(A[:npp, j] ** 2)) if npp > 0 else 0.0
            if (unorm + abs(a_new) * factor) - unorm > 0.0:
                # apply to b (into zz) and test the candidate coefficient
                zz[:] = b
                _h12_apply(col, up, a_new, zz[npp:])
                if c:
                    c.flops += 4 * (m - npp)
                ztest = zz[npp] / a_new
                if ztest > 0.0:
                    accepted = True
                    break
            w[j] = 0.0
        if not accepted:
            break
        # accept column j
        b[:] = zz
        A[npp, j] = a_new
        others = [int(jj) for jj in Z if jj != j]
        for jj in others:
            _h12_apply(col, up, a_new, A[npp:, jj])
        if c:
            c.flops += 4 * (m - npp) * len(others)
            c.outer += 1
        A[npp + 1:, j] = 0.0
        inP[j] = True
        order.append(j)
        w[j] = 0.0
        # solve the triangular system
        z = _back_substitute(A, b, order)
        if c:
            c.flops += len(order) ** 2
        done = False
        while True:
            it += 1
            if it > itmax:
                mode = 3
                done = True
                break
            zneg = [(ip, jj) for ip, jj in enumerate(order) if z[ip] <= 0.0]
            if not zneg:
                break
            alpha = 2.0
            jblock = -1
            for ip, jj in zneg:
                t = -x[jj] / (z[ip] - x[jj])
                if alpha > t:
                    alpha = t
                    jblock = ip
            if jblock < 0:
                break
            for ip, jj in enumerate(order):
                x[jj] += alpha * (z[ip] - x[jj])
            if c:
                c.flops += 3 * len(order)
            # remove the blocking column, then any other column whose x dropped to <= 0
            ip = jblock
            while True:
                jj = order[ip]
                x[jj] = 0.0
                _remove_column(A, b, order, ip, n)
                if c:
                    c.flops += (6 * n + 12) * (len(order) - ip)
                    c.removals += 1
                inP[jj] = False
                nxt = [q for q, cj in enumerate(order) if x[cj] <= 0.0]
                if not nxt:
                    break
                ip = nxt[0]
            z = _back_substitute(A, b, order)
            if c:
                c.flops += len(order) ** 2
        if done:
            break
        for ip, jj in enumerate(order):
            x[jj] = z[ip]
    npp = len(order)
    rnorm = math.sqrt(float(np.sum(b[npp:] ** 2))) if npp < m else 0.0
    if c:
        c.solves += 1
    return x, rnorm, mode


def _h12_apply(u, up, s, v):
    """Apply the Householder reflection defined by pivot column `u` (pivot element replaced by `up`, new pivot `s`)."""
    bb = up * s
    if bb >= 0.0:
        return
    sm = v[0] * up + float(np.dot(u[1:], v[1:]))
    if sm != 0.0:
        sm = sm / bb
        v[0] += sm * up
        v[1:] += sm * u[1:]


def _back_substitute(A, b, order):
    p = len(order)
    z = np.zeros(p)
    rhs = b[:p].copy()
    for ip in range(p - 1, -1, -1):
        jj = order[ip]
        z[ip] = rhs[ip] / A[ip, jj]
        rhs[:ip] -= A[:ip, jj] * z[ip]
    return z


def _remove_column(A, b, order, ip, n):
    """Delete position `ip` of the positive set and re-triangularise with Givens rotations over all n columns."""
    p = len(order)
    for q in range(ip + 1, p):
        jj = order[q]
        a1, a2 = A[q - 1, jj], A[q, jj]
        if abs(a1) > abs(a2):
            xr = a2 / a1
            yr = math.sqrt(1.0 + xr * xr)
            cc = math.copysign(1.0 / yr, a1)
            ss = cc * xr
            sig = abs(a1) * yr
        elif a2 != 0.0:
            xr = a1 / a2
            yr = math.sqrt(1.0 + xr * xr)
            ss = math.copysign(1.0 / yr, a2)
            cc = ss * xr
            sig = abs(a2) * yr
        else:
            sig, cc, ss = 0.0, 0.0, 1.0
        A[q - 1, jj] = sig
        A[q, jj] = 0.0
        for l in range(n):
            if l != jj:
                t = A[q - 1, l]
                A[q - 1, l] = cc * t + ss * A[q, l]
                A[q, l] = -ss * t + cc * A[q, l]
        t = b[q - 1]
        b[q - 1] = cc * t + ss * b[q]
        b[q] = -ss * t + cc * b[q]
    del order[ip]


# ----------------------------------------------------------------------------------------------------------------
# Bounded Brent — scipy/optimize/_optimize.py:_minimize_scalar_bounded
# ----------------------------------------------------------------------------------------------------------------


def brent_bounded(func, x1, x2, xatol=1e-5, maxfun=500, trace=None):
    """Restatement of SciPy's bounded Brent; returns (xf, fval, nfev).  `trace` collects the evaluated abscissae."""
    sqrt_eps = math.sqrt(2.2e-16)
    golden_mean = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = x1, x2
    fulc = a + golden_mean * (b - a)
    nfc, xf = fulc, fulc
    rat = e = 0.0
    x = xf
    fx = func(x)
    if trace is not None:
        trace.append(x)
    num = 1
    ffulc = fnfc = fx
    xm = 0.5 * (a + b)
    tol1 = sqrt_eps * abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        golden = 1
        if abs(e) > tol1:
            golden = 0
            r = (xf - nfc) * (fx - ffulc)
            q = (xf - fulc) * (fx - fnfc)
            p = (xf - fulc) * q - (xf - nfc) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = rat
            if (abs(p) < abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                rat = (p + 0.0) / q
                x = xf + rat
                if ((x - a) < tol2) or ((b - x) < tol2):
                    si = np.sign(xm - xf) + ((xm - xf) == 0)
                    rat = tol1 * si
            else:
                golden = 1
        if golden:
            if xf >= xm:
                e = a - xf
            else:
                e = b - xf
            rat = golden_mean * e
        si = np.sign(rat) + (rat == 0)
        x = xf + si * max(abs(rat), tol1)
        fu = func(x)
        if trace is not None:
            trace.append(x)
        num += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            fulc, ffulc = nfc, fnfc
            nfc, fnfc = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if (fu <= fnfc) or (nfc == xf):
                fulc, ffulc = nfc, fnfc
                nfc, fnfc = x, fu
            elif (fu <= ffulc) or (fulc == xf) or (fulc == nfc):
                fulc, ffulc = x, fu
        xm = 0.5 * (a + b)
        tol1 = sqrt_eps * abs(xf) + xatol / 3.0
        tol2 = 2.0 * tol1
        if num >= maxfun:
            break
    return xf, fx, num


# ----------------------------------------------------------------------------------------------------------------
# Flip-angle estimation — flip_angle_algorithms/fa_estimation.py
# ----------------------------------------------------------------------------------------------------------------


def compute_optimal_FA(M, Dic_3D, alpha_values):
    """Brute-force search — fa_estimation.py:74-90.  Returns (index, alpha, km, SSE, f)."""
    nA = Dic_3D.shape[2]
    M = np.ascontiguousarray(M)
    residual = np.zeros(nA)
    for i in range(nA):
        residual[i] = nnls(np.ascontiguousarray(Dic_3D[:, :, i]), M)[1]
    index = int(np.argmin(residual))
    Di = np.ascontiguousarray(Dic_3D[:, :, index])
    f, _ = nnls(Di, M)
    km = np.sum(f)
    SSE = np.sum((np.dot(Di, f) - M) ** 2)
    return index, alpha_values[index], km, SSE, f


def spline_optimal_FA(M, Dic_3D_LR, Dic_3D, alpha_values_spline, alpha_values, return_debug=False):
    """Per-voxel body of fitting_slice_FA_spline_method — fa_estimation.py:46-65."""
    nk = Dic_3D_LR.shape[2]
    residual = np.zeros(nk)
    for i in range(nk):
        residual[i] = nnls(np.ascontiguousarray(Dic_3D_LR[:, :, i]), M)[1]
    f2 = interp1d(alpha_values_spline, residual, kind="cubic")
    res = minimize_scalar(f2, method="Bounded", bounds=(90.0, 180.0))
    index = int(np.argmin(np.abs(alpha_values - res.x)))
    fsol, _ = nnls(np.ascontiguousarray(Dic_3D[:, :, index]), M)
    km = np.sum(fsol)
    if return_debug:
        return index, alpha_values[index], km, fsol, residual, res.x
    return index, alpha_values[index], km, fsol


def _voxel_ok(mask_v, M):
    return (mask_v > 0.0) and (np.sum(M) > 0.0)


def fitting_slice_FA_brute_force(mask_1d, data_1d, nx, Dic_3D, alpha_values):
    """Row worker — fa_estimation.py:92-112."""
    FA = np.zeros(nx)
    FA_index = np.zeros(nx)
    KM = np.zeros(nx)
    Fsol = 0.0
    if np.count_nonzero(mask_1d) > 0:
        for v in range(nx):
            if _voxel_ok(mask_1d[v], data_1d[v, :]):
                idx, a, km, _sse, f = compute_optimal_FA(data_1d[v, :], Dic_3D, alpha_values)
                FA[v], FA_index[v], KM[v] = a, idx, km
                Fsol = Fsol + f
    return FA, FA_index, KM, Fsol


def fitting_slice_FA_spline_method(Dic_3D_LR, Dic_3D, data_1d, mask_1d, alpha_values_spline, nx, alpha_values):
    """Row worker — fa_estimation.py:35-70."""
    FA = np.zeros(nx)
    FA_index = np.zeros(nx)
    KM = np.zeros(nx)
    Fsol = 0.0
    if np.count_nonzero(mask_1d) > 0:
        for v in range(nx):
            if _voxel_ok(mask_1d[v], data_1d[v, :]):
                idx, a, km, f = spline_optimal_FA(data_1d[v, :], Dic_3D_LR, Dic_3D, alpha_values_spline, alpha_values)
                FA[v], FA_index[v], KM[v] = a, idx, km
                Fsol = Fsol + f
    return FA, FA_index, KM, Fsol


# ----------------------------------------------------------------------------------------------------------------
# Regularised solvers — intravoxel_algorithms/algorithms.py
# ----------------------------------------------------------------------------------------------------------------


def nnls_tik(D, M, L, reg):
    """nnls([D; sqrt(reg) L], [M; 0]) — algorithms.py:262-269."""
    n = D.shape[1]
    f, _ = nnls(np.concatenate((D, np.sqrt(reg) * L)), np.concatenate((M, np.zeros(n))))
    return f


def obj_nnls_x2(x, D, L, Maug, SSE, factor, M):
    """algorithms.py:226-233."""
    f, _ = nnls(np.concatenate((D, np.sqrt(x) * L)), Maug)
    SSEr = np.sum((np.dot(D, f) - M) ** 2)
    return np.abs(SSEr - factor * SSE) / SSE


def nnls_x2(D, M, L, factor):
    """Chi-square regularisation, bounded Brent on [0, 10] — algorithms.py:211-224.  Returns (f, reg_opt, k_est)."""
    f0, _ = nnls(D, M)
    SSE = np.sum((np.dot(D, f0) - M) ** 2)
    n = D.shape[1]
    Maug = np.concatenate((M, np.zeros(n)))
    with np.errstate(all="ignore"):
        reg = fminbound(obj_nnls_x2, 0.0, 10.0, args=(D, L, Maug, SSE, factor, M), xtol=1e-05, maxfun=300,
                        full_output=0, disp=0)
        f, _ = nnls(np.concatenate((D, np.sqrt(reg) * L)), Maug)
        k_est = np.sum((np.dot(D, f) - M) ** 2) / SSE
    return f, reg, k_est


def scale_curve_1(a):
    """algorithms.py:200-204 with l, u = -10, 10."""
    vmin, vmax = a.min(), a.max()
    l, u = -10, 10
    return ((u - l) / (vmax - vmin)) * (a - (u * vmin - l * vmax) / (u - l))


def select_corner(x, y):
    """Triangle method for the L-curve corner — algorithms.py:150-191."""
    with np.errstate(all="ignore"):
        x = scale_curve_1(x)
        y = scale_curve_1(y)
        n = len(x)
        corner = n - 1
        cte = 7.0 * np.pi / 8.0
        angmin = None
        cx, cy = x[-1], y[-1]
        for k in range(0, n - 2):
            bx, by = x[k], y[k]
            for j in range(k + 1, n - 1):
                ax, ay = x[j], y[j]
                ab = np.sqrt((ax - bx) ** 2 + (ay - by) ** 2)
                ac = np.sqrt((ax - cx) ** 2 + (ay - cy) ** 2)
                bc = np.sqrt((bx - cx) ** 2 + (by - cy) ** 2)
                cosa = (ab ** 2 + ac ** 2 - bc ** 2) / (2.0 * ab * ac)
                cosa = max(-1.0, min(cosa, 1.0))
                ang = np.arccos(cosa)
                area = 0.5 * ((bx - ax) * (ay - cy) - (ax - cx) * (by - ay))
                if area > 0 and (ang < cte and (angmin is None or ang < angmin)):
                    corner = j
                    angmin = ang
    return corner


def lcurve_curves(D, y, L, lambda_reg):
    """The two log curves of nnls_lcurve_wrapper — algorithms.py:88-108."""
    n = D.shape[1]
    b = np.concatenate((y, np.zeros(n)))
    nl = len(lambda_reg)
    log_err = np.zeros(nl)
    log_nrm = np.zeros(nl)
    for i in range(nl):
        x, _ = nnls(np.concatenate((D, np.sqrt(lambda_reg[i]) * L)), b)
        log_err[i] = np.log(np.sum((np.dot(D, x) - y) ** 2.0) + 1e-200)
        log_nrm[i] = np.log(np.sum((np.dot(L, x)) ** 2.0) + 1e-200)
    return log_err, log_nrm


def nnls_lcurve_wrapper(D, y, L, lambda_reg):
    """algorithms.py:88-113: returns the grid lambda at the corner."""
    log_err, log_nrm = lcurve_curves(D, y, L, lambda_reg)
    return lambda_reg[select_corner(log_err, log_nrm)]


def obj_nnls_gcv(x, D, L, Maug, m, Im):
    """algorithms.py:285-296 including its indexing quirk: L[f>0, f>0] is the 1-D vector of diagonal entries, so
    LTL is a scalar that is added to every entry of Dr^T Dr."""
    f, SSEr = nnls(np.concatenate((D, np.sqrt(x) * L)), Maug)
    sel = f > 0
    Dr = D[:, sel]
    Lr = L[sel, sel]
    DTD = np.matmul(Dr.T, Dr)
    LTL = np.matmul(Lr.T, Lr)
    A = np.matmul(Dr, np.linalg.lstsq(DTD + x * LTL, Dr.T, rcond=None)[0])
    cost = ((1.0 / m) * (SSEr ** 2.0)) / ((1.0 / m) * np.trace(Im - A)) ** 2.0
    return np.log(cost)


def nnls_gcv(D, M, L):
    """algorithms.py:276-283: bounded Brent on [1e-8, 10]."""
    m, n = D.shape
    Maug = np.concatenate((M, np.zeros(n)))
    Im = np.eye(m)
    with np.errstate(all="ignore"):
        reg = fminbound(obj_nnls_gcv, 1e-8, 10.0, args=(D, L, Maug, m, Im), xtol=1e-05, maxfun=300, full_output=0,
                        disp=0)
    f, _ = nnls(np.concatenate((D, np.sqrt(reg) * L)), Maug)
    return f, reg


def obj_BayesReg_nnls(x, D, L, Maug, M, m, n, B, det_L, beta, K):
    """bayesian_interpolation.py:107-126."""
    f, _ = nnls(np.concatenate((D, np.sqrt(x) * L)), Maug)
    ED = 0.5 * np.sum((np.dot(D, f) - M) ** 2)
    EW = 0.5 * np.sum(np.dot(L, f) ** 2)
    A = beta * B + (beta * x) * K
    U = cholesky(A, lower=False, overwrite_a=True, check_finite=False)
    det_U = np.prod(np.diag(U))
    err1 = 1.0 + erf((1.0 / np.sqrt(2.0)) * np.dot(U, f))
    series = np.sum(np.log(err1))
    c1 = beta * ED + beta * x * EW + np.log(det_U) - (n / 2.0) * np.log(np.pi / 2.0) - series
    c2 = ((m / 2.0) * np.log(2.0 * np.pi) - (m / 2.0) * np.log(beta) + (n / 2.0) * np.log(np.pi)
          - (n / 2.0) * np.log(2 * beta * x) - np.log(det_L))
    return c1 + c2


def BayesReg_nnls(D, M, L):
    """bayesian_interpolation.py:84-105: sigma from the plain NNLS residual once, then bounded Brent on [1e-8, 2]."""
    m, n = D.shape
    Maug = np.concatenate((M, np.zeros(n)))
    x0, _ = nnls(D, M)
    nnz = np.sum(x0 > 0)
    dof = np.max([m - nnz, 1.0])
    sigma = np.sqrt(np.sum((M - np.dot(D, x0)) ** 2) / dof)
    beta = 1.0 / sigma ** 2
    B = np.matmul(D.T, D)
    K = np.matmul(L.T, L)
    det_L = det(L)
    with np.errstate(all="ignore"):
        reg = fminbound(obj_BayesReg_nnls, 1e-8, 2.0, args=(D, L, Maug, M, m, n, B, det_L, beta, K), xtol=1e-05,
                        maxfun=200, full_output=0, disp=0)
    f, _ = nnls(np.concatenate((D, np.sqrt(reg) * L)), Maug)
    return f, reg


# ----------------------------------------------------------------------------------------------------------------
# Row dispatcher and metrics — motor/motor_recon_met2_real_data.py:113-162, 443-472
# ----------------------------------------------------------------------------------------------------------------


def t2_fit_voxel(M, Kernel, reg_method, Laplac, lambda_reg):
    """Per-voxel body of fitting_slice_T2 (motor...:126-155); M is the raw signal with M[0] > 0."""
    km = M[0]
    Mn = M / km
    if reg_method == "NNLS":
        x, _ = nnls(Kernel, Mn)
        reg = 0
    elif reg_method == "T2SPARC":
        reg = 1.8
        x = nnls_tik(Kernel, Mn, Laplac, reg)
    elif reg_method == "X2":
        x, _reg, k_est = nnls_x2(Kernel, Mn, Laplac, 1.02)
        reg = k_est
    elif reg_method == "L_curve":
        reg = nnls_lcurve_wrapper(Kernel, Mn, Laplac, lambda_reg)
        x = nnls_tik(Kernel, Mn, Laplac, reg)
    elif reg_method == "GCV":
        x, reg = nnls_gcv(Kernel, Mn, Laplac)
    elif reg_method == "BayesReg":
        x, reg = BayesReg_nnls(Kernel, Mn, Laplac)
    else:
        raise ValueError(reg_method)
    return x * km, np.dot(Kernel, x) * km, reg


def fitting_slice_T2(mask_1d, data_1d, FA_index_1d, nx, Dic_3D, lambda_reg, T2dim, nEchoes, reg_method, Laplac,
                     dist_x_prior=None):
    """Row worker — motor/motor_recon_met2_real_data.py:113-162."""
    f_sol = np.zeros((nx, T2dim))
    sig = np.zeros((nx, nEchoes))
    reg = np.zeros(nx)
    if np.count_nonzero(mask_1d) > 0:
        for v in range(nx):
            if _voxel_ok(mask_1d[v], data_1d[v, :]):
                M = np.ascontiguousarray(data_1d[v, :])
                Kernel = np.ascontiguousarray(Dic_3D[:, :, int(FA_index_1d[v])])
                if M[0] > 0:
                    f_sol[v, :], sig[v, :], reg[v] = t2_fit_voxel(M, Kernel, reg_method, Laplac, lambda_reg)
    return f_sol, sig, reg


def voxel_metrics(x_sol, T2s, ind_m, ind_t, ind_csf):
    """Step-4 metrics of one voxel — motor...:452-468.  Returns (MWF, IEWF, FWF, T2_M, T2_IE, TWC)."""
    logT2 = np.log(T2s)
    vt = np.sum(x_sol) + EPSILON
    x = x_sol / vt
    mwf = np.sum(x[ind_m])
    iewf = np.sum(x[ind_t])
    fwf = np.sum(x[ind_csf])
    t2m = np.exp(np.sum(x[ind_m] * logT2[ind_m]) / (np.sum(x[ind_m]) + EPSILON))
    t2ie = np.exp(np.sum(x[ind_t] * logT2[ind_t]) / (np.sum(x[ind_t]) + EPSILON))
    return mwf, iewf, fwf, t2m, t2ie, vt


# ----------------------------------------------------------------------------------------------------------------
# Whole-volume driver with the orchestrator's loop structure (motor...:349-373, 428-472), without I/O and plotting.
# ----------------------------------------------------------------------------------------------------------------


def _grids(reg_method, reg_matrix, FA_method, myelin_T2, n_echoes, tau, TR, npc=None, n_alphas=None):
    if npc is None:
        npc = 96 if reg_method == "T2SPARC" else 60
    T2s = np.logspace(math.log10(10.0), math.log10(2000.0), num=npc, endpoint=True, base=10.0)
    T1s = 1000.0 * np.ones_like(T2s)
    ind_m = T2s <= myelin_T2
    ind_t = (T2s > myelin_T2) & (T2s <= 200.0)
    ind_csf = T2s >= 200.0
    if FA_method == "spline":
        alpha_values = np.linspace(90.0, 180.0, 273 if n_alphas is None else n_alphas)
        alpha_spline = np.linspace(90.0, 180.0, 15)
    else:
        alpha_values = np.linspace(90.0, 180.0, 91 if n_alphas is None else n_alphas)
        alpha_spline = None
    lambda_reg = np.zeros(50)
    lambda_reg[1:] = np.logspace(math.log10(1e-8), math.log10(10.0), num=49, endpoint=True, base=10.0)
    if reg_matrix == "I":
        L = np.eye(npc)
    elif reg_matrix == "L1":
        L = np.eye(npc) - np.eye(npc, k=-1)
    elif reg_matrix == "L2":
        L = 2.0 * np.eye(npc) - np.eye(npc, k=-1) - np.eye(npc, k=1)
        L[0, 0] = 1.0
        L[-1, -1] = 1.0
    elif reg_matrix == "InvT2":
        T2s_mod = np.concatenate((np.array([T2s[0] - 1.0]), T2s[:-1]))
        d = T2s - T2s_mod
        d[0] = d[1]
        L = np.diag(1.0 / d)
    else:
        raise ValueError(reg_matrix)
    return dict(T2s=T2s, T1s=T1s, ind_m=ind_m, ind_t=ind_t, ind_csf=ind_csf, alpha_values=alpha_values,
                alpha_spline=alpha_spline, lambda_reg=lambda_reg, L=L, npc=npc)


def _fa_row(args):
    (FA_method, mask_1d, data_1d, nx, Dic, Dic_LR, alpha_values, alpha_spline) = args
    if FA_method == "spline":
        return fitting_slice_FA_spline_method(Dic_LR, Dic, data_1d, mask_1d, alpha_spline, nx, alpha_values)
    return fitting_slice_FA_brute_force(mask_1d, data_1d, nx, Dic, alpha_values)


def _t2_row(args):
    (mask_1d, data_1d, fa_idx_1d, nx, Dic, lambda_reg, npc, nte, reg_method, L) = args
    return fitting_slice_T2(mask_1d, data_1d, fa_idx_1d, nx, Dic, lambda_reg, npc, nte, reg_method, L, None)


def recon_volume(data, mask, TE_array, TR, reg_method, reg_matrix, FA_method, myelin_T2=40.0, num_cores=1,
                 npc=None, n_alphas=None, Dic_3D=None, Dic_3D_LR=None, data_fa=None, pool=None):
    """Steps 2-4 of motor_recon_met2 (motor...:204-277, 349-373, 428-472) on in-memory arrays.

    `num_cores` > 1 runs the row tasks in a multiprocessing pool, one batch of `ny` row tasks per slice like the
    reference's joblib loops (the pool is kept across slices instead of being re-created).  Returns a dict of arrays.
    """
    data = np.array(data, dtype=np.float64, copy=True)
    mask = np.asarray(mask).astype(np.int64)
    nx, ny, nz, nt = data.shape
    for c in range(nt):
        data[:, :, :, c] = data[:, :, :, c] * mask
    n_echoes = TE_array.shape[0]
    tau = TE_array[1] - TE_array[0]
    g = _grids(reg_method, reg_matrix, FA_method, myelin_T2, n_echoes, tau, TR, npc, n_alphas)
    npc = g["npc"]
    if Dic_3D is None:
        Dic_3D = create_Dic_3D(npc, g["T2s"], g["T1s"], n_echoes, tau, g["alpha_values"], TR)
    if FA_method == "spline" and Dic_3D_LR is None:
        Dic_3D_LR = create_Dic_3D(npc, g["T2s"], g["T1s"], n_echoes, tau, g["alpha_spline"], TR)
    data[data < 0.0] = 0.0
    data_smooth = data if data_fa is None else data_fa
    FA = np.zeros((nx, ny, nz))
    FA_index = np.zeros((nx, ny, nz))
    Ktotal = np.zeros((nx, ny, nz))
    f_sol_4D = np.zeros((nx, ny, nz, npc))
    s_sol_4D = np.zeros((nx, ny, nz, n_echoes))
    reg_param = np.zeros((nx, ny, nz))
    own_pool = None
    if num_cores != 1 and pool is None:
        import multiprocessing
        own_pool = pool = multiprocessing.Pool(None if num_cores in (-1, None) else num_cores)
    mapper = pool.map if pool is not None else (lambda f, it: [f(a) for a in it])
    try:
        mean_T2_dist = 0
        for z in range(nz):
            tasks = [(FA_method, mask[:, y, z], data_smooth[:, y, z, :], nx, Dic_3D, Dic_3D_LR, g["alpha_values"],
                      g["alpha_spline"]) for y in range(ny)]
            res = mapper(_fa_row, tasks)
            for y in range(ny):
                FA[:, y, z], FA_index[:, y, z], Ktotal[:, y, z] = res[y][0], res[y][1], res[y][2]
                mean_T2_dist = mean_T2_dist + res[y][3]
        for z in range(nz):
            tasks = [(mask[:, y, z], data[:, y, z, :], FA_index[:, y, z], nx, Dic_3D, g["lambda_reg"], npc, n_echoes,
                      reg_method, g["L"]) for y in range(ny)]
            res = mapper(_t2_row, tasks)
            for y in range(ny):
                f_sol_4D[:, y, z, :], s_sol_4D[:, y, z, :], reg_param[:, y, z] = res[y]
    finally:
        if own_pool is not None:
            own_pool.close()
            own_pool.join()
    maps = {k: np.zeros((nx, ny, nz)) for k in ("MWF", "IEWF", "FWF", "T2_M", "T2_IE")}
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                if mask[ix, iy, iz] > 0.0:
                    r = voxel_metrics(f_sol_4D[ix, iy, iz, :], g["T2s"], g["ind_m"], g["ind_t"], g["ind_csf"])
                    (maps["MWF"][ix, iy, iz], maps["IEWF"][ix, iy, iz], maps["FWF"][ix, iy, iz],
                     maps["T2_M"][ix, iy, iz], maps["T2_IE"][ix, iy, iz], Ktotal[ix, iy, iz]) = r
    out = dict(maps)
    out.update(TWC=Ktotal, FA=FA, FA_index=FA_index, fsol_4D=f_sol_4D, Est_Signal=s_sol_4D, reg_param=reg_param,
               mean_T2_dist=mean_T2_dist, T2s=g["T2s"], alpha_values=g["alpha_values"])
    return out


# ----------------------------------------------------------------------------------------------------------------
# Rows of SURVEY.md §8f either side of the voxel fit: NESMA denoiser, mean-spectrum diagnostics, ROI-based estimator.
# ----------------------------------------------------------------------------------------------------------------


def nesma_filter(data, mask, path_size=(6, 6, 6), threshold=2.5):
    """NESMA denoiser — motor/motor_recon_met2_real_data.py:305-333 (same loops, same NumPy expressions)."""
    data = np.asarray(data, dtype=np.float64)
    nx, ny, nz, nt = data.shape
    data_den = np.zeros_like(data)
    with np.errstate(all="ignore"):
        for voxelx in range(nx):
            min_x = np.max([voxelx - path_size[0], 0])
            max_x = np.min([voxelx + path_size[0], nx])
            for voxely in range(ny):
                min_y = np.max([voxely - path_size[1], 0])
                max_y = np.min([voxely + path_size[1], ny])
                for voxelz in range(nz):
                    if mask[voxelx, voxely, voxelz] == 1:
                        min_z = np.max([voxelz - path_size[2], 0])
                        max_z = np.min([voxelz + path_size[2], nz])
                        signal_path = data[min_x:max_x, min_y:max_y, min_z:max_z, :]
                        dim = signal_path.shape
                        signal_path2D = signal_path.reshape((np.prod(dim[0:3]), nt))
                        signal_xyz = data[voxelx, voxely, voxelz]
                        RE = 100 * np.sum(np.abs(signal_path2D - signal_xyz), axis=1) / np.sum(signal_xyz)
                        ind_valid = RE < threshold
                        if ind_valid.any():
                            data_den[voxelx, voxely, voxelz] = np.mean(signal_path2D[ind_valid, :], axis=0)
                        else:
                            data_den[voxelx, voxely, voxelz] = np.nan   # np.mean of an empty selection
    return data_den


def segment_mean(data, FA_index, Dic_3D, selector):
    """total_signal / total_Kernel of motor...:377-392 (selector = mask == 1) and of
    motor_recon_met2_real_data_ROI.py:408-423 (selector = ROIs == label): sequential sums in (x, y, z) order."""
    nx, ny, nz, nt = data.shape
    total_signal = 0
    total_Kernel = 0
    nv = 0
    for voxelx in range(nx):
        for voxely in range(ny):
            for voxelz in range(nz):
                if selector[voxelx, voxely, voxelz]:
                    total_signal = total_signal + data[voxelx, voxely, voxelz, :]
                    ind_xyz = int(FA_index[voxelx, voxely, voxelz])
                    total_Kernel = total_Kernel + Dic_3D[:, :, ind_xyz]
                    nv = nv + 1.0
    return total_signal / nv, total_Kernel / nv, nv


def mean_spectrum_diagnostics(data, mask, FA_index, Dic_3D, mean_T2_dist, factor=1.01):
    """The three curves of the 'Mean_spectrum_from_all_voxels' figure — motor...:375-403."""
    npc = Dic_3D.shape[1]
    mean_T2_dist = mean_T2_dist / np.sum(mean_T2_dist)
    total_signal, total_Kernel, nv = segment_mean(data, FA_index, Dic_3D, np.asarray(mask) == 1)
    fmean1, SSE = nnls(total_Kernel, total_signal)
    fmean2, reg_opt2, k_est = nnls_x2(total_Kernel, total_signal, np.eye(npc), factor)
    return dict(mean_T2_dist=mean_T2_dist, dist_T2_mean1=fmean1 / np.sum(fmean1), dist_T2_mean2=fmean2 / np.sum(fmean2),
                total_signal=total_signal, total_Kernel=total_Kernel, nv=nv, reg_opt2=reg_opt2, k_est=k_est)


def roi_estimates(data, mask, ROIs, FA_index, Dic_3D, Laplac, T2s, ind_m, ind_t, ind_csf, factor=1.01):
    """ROI-based estimator — motor/motor_recon_met2_real_data_ROI.py:166-183 (labels) and :405-445 (per-ROI X2 fit and
    metrics)."""
    ROIs = np.asarray(ROIs).astype(np.int64)
    values = np.unique(ROIs)
    values = np.delete(values, np.where(values == 0))
    ROIs = ROIs * np.asarray(mask).astype(np.int64)
    logT2 = np.log(T2s)
    out = dict(roi_values=values, fsol_ROIs=np.zeros((len(values), len(T2s))))
    for k in ("MWF", "IEWF", "FWF", "T2M", "T2IE", "TWC", "reg_opt", "k_est"):
        out[k] = np.zeros(len(values))
    for i, val in enumerate(values):
        total_signal, total_Kernel, nv = segment_mean(data, FA_index, Dic_3D, ROIs == val)
        x_sol, reg_opt2, k_est = nnls_x2(total_Kernel, total_signal, Laplac, factor)
        vt = np.sum(x_sol) + EPSILON
        x_sol = x_sol / vt
        out["fsol_ROIs"][i] = x_sol
        out["MWF"][i] = np.sum(x_sol[ind_m])
        out["IEWF"][i] = np.sum(x_sol[ind_t])
        out["FWF"][i] = np.sum(x_sol[ind_csf])
        out["T2M"][i] = np.exp(np.sum(x_sol[ind_m] * logT2[ind_m]) / (np.sum(x_sol[ind_m]) + EPSILON))
        out["T2IE"][i] = np.exp(np.sum(x_sol[ind_t] * logT2[ind_t]) / (np.sum(x_sol[ind_t]) + EPSILON))
        out["TWC"][i] = vt
        out["reg_opt"][i], out["k_est"][i] = reg_opt2, k_est
    return out


def nnls_gcv_grid(D, M, L, lambdas):
    """Extension for BASELINE.json configs[2] ("GCV over a 50-point lambda grid"; the reference itself uses Brent,
    algorithms.py:280): the reference's own objective obj_nnls_gcv (:285-296) evaluated on the grid, arg-min (np.argmin:
    first minimum), final solve with nnls_tik (:262).  Returns (f, lambda, costs)."""
    m, n = D.shape
    Maug = np.concatenate((M, np.zeros(n)))
    Im = np.eye(m)
    with np.errstate(all="ignore"):
        costs = np.array([obj_nnls_gcv(x, D, L, Maug, m, Im) for x in lambdas])
    reg = lambdas[int(np.argmin(costs))]
    return nnls_tik(D, M, L, reg), reg, costs
