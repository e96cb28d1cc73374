"""Prototype (CPU, numpy): GCV trace through a rank-revealing pivoted Cholesky of Mk = Dr^T Dr + x s 1 1^T followed by the
eigen-decomposition of the small r x r matrix C^T C, against the reference formulation (lstsq on the k x k matrix,
algorithms.py:285-296).  Measures what the reduced form costs in objective-level agreement."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import met2_oracle as O
from multicomponent_t2_toolbox_b200.grids import create_Laplacian_matrix
from multicomponent_t2_toolbox_b200.phantom import make_phantom

EPS = 2.220446049250313e-16

def trace_ref(D, f, x, L):
    sel = f > 0
    Dr = D[:, sel]; Lr = L[sel, sel]
    M = Dr.T @ Dr + x * (Lr.T @ Lr)
    A = Dr @ np.linalg.lstsq(M, Dr.T, rcond=None)[0]
    return np.trace(A)

def trace_eig(G, f, x, L):
    sel = np.nonzero(f > 0)[0]
    k = len(sel)
    s = (np.diag(L)[sel] ** 2).sum(); xs = x * s
    M = G[np.ix_(sel, sel)] + xs
    w, V = np.linalg.eigh(M)
    tau = EPS * k * w.max()
    e = V.sum(0)
    keep = w > tau
    return (1 - xs * e[keep] ** 2 / w[keep]).sum()

def trace_pivchol(G, f, x, L, tol=1e-16, rcap=20, stats=None):
    sel = np.nonzero(f > 0)[0]
    k = len(sel)
    if k == 0: return 0.0
    s = (np.diag(L)[sel] ** 2).sum(); xs = x * s
    d = G[sel, sel] + xs
    dmax = d.max()
    done = np.zeros(k, bool)
    C = np.zeros((k, rcap))
    r = 0
    while r < min(rcap, k):
        dd = np.where(done, -1.0, d)
        piv = int(np.argmax(dd))
        if not (dd[piv] > tol * dmax): break
        col = G[sel, sel[piv]] + xs - C[:, :r] @ C[piv, :r]
        c = col / np.sqrt(dd[piv])
        c[done] = 0.0
        c[piv] = np.sqrt(dd[piv])
        C[:, r] = c
        d = d - c * c
        done[piv] = True
        r += 1
    if stats is not None: stats.append((k, r))
    C = C[:, :r]
    H = C.T @ C
    g = C.sum(0)
    w, W = np.linalg.eigh(H)
    tau = EPS * k * w.max()
    e = g @ W
    keep = w > tau
    return (1 - xs * e[keep] ** 2 / w[keep] ** 2).sum()

def main():
    ph = make_phantom((16, 16, 4), seed=1)
    sig = ph["data"].reshape(-1, 32)[:96]
    T2s = np.logspace(1, np.log10(2000.0), 60)
    a273 = np.linspace(90.0, 180.0, 273)
    idxs = [100, 180, 230, 272]
    Dic = O.create_Dic_3D(60, T2s, 1000.0 * np.ones(60), 32, 10.0, a273[idxs], 1000.0)
    for rm in ("I", "L2"):
        L = create_Laplacian_matrix(60, 0 if rm == "I" else 2)
        for lam in (1e-5, 1e-3, 0.1, 3.8197):
            d1, d2 = [], []
            stats = []
            for v in range(96):
                D = np.ascontiguousarray(Dic[:, :, v % 4]); G = D.T @ D
                M = sig[v] / sig[v, 0]
                f, sser = O.nnls(np.concatenate((D, np.sqrt(lam) * L)), np.concatenate((M, np.zeros(60))))
                t0 = trace_ref(D, f, lam, L)
                t1 = trace_eig(G, f, lam, L)
                t2 = trace_pivchol(G, f, lam, L, stats=stats)
                # objective = log(num / ((m - tr)/m)^2): d obj = 2 * dlog(m - tr)
                d1.append(abs(2 * np.log((32 - t1) / (32 - t0))))
                d2.append(abs(2 * np.log((32 - t2) / (32 - t0))))
            d1, d2 = np.array(d1), np.array(d2)
            ks = np.array(stats)
            print(rm, lam, "eig: med %.1e max %.1e frac<1e-2 %.3f | pivchol: med %.1e max %.1e frac<1e-2 %.3f | k %.1f r %.1f rmax %d"
                  % (np.median(d1), d1.max(), (d1 < 1e-2).mean(), np.median(d2), d2.max(), (d2 < 1e-2).mean(),
                     ks[:, 0].mean(), ks[:, 1].mean(), ks[:, 1].max()))

main()
