"""Design prototype: warm-started Gram-domain LH inside the X2 Brent loop vs the oracle (parity of path-independent result)."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import met2_oracle as O
from proto_gram_nnls2 import TInv
from scipy.optimize import fminbound


def gram_nnls_ws(G, c, m_rows, P0=None, x0=None, stats=None):
    n = G.shape[0]; itmax = 3 * n
    x = np.zeros(n); P = []; inP = np.zeros(n, bool); F = TInv(); it = 0
    skip_outer = False
    if P0:
        for j in P0:
            res = F.try_append(G[P, j], G[j, j], c[j])
            if res is None:
                continue
            F.append(*res); P.append(j); inP[j] = True
        x[P] = x0[P]
        z = F.solve()
        skip_outer = True
    while True:
        if not skip_outer:
            p = len(P)
            if p >= n or p >= m_rows: break
            w = c - G[:, P] @ x[P] if p else c.copy()
            accepted = False; rejected = np.zeros(n, bool)
            while True:
                wz = np.where(inP | rejected, -np.inf, w); j = int(np.argmax(wz))
                if not (wz[j] > 0): break
                res = F.try_append(G[P, j], G[j, j], c[j])
                if res is None or not (res[2] > 0):
                    rejected[j] = True; continue
                F.append(*res); P.append(j); inP[j] = True; accepted = True; break
            if not accepted: break
            if stats is not None: stats['outer'] = stats.get('outer', 0) + 1
            z = F.solve()
        skip_outer = False
        done = False
        while True:
            it += 1
            if it > itmax: done = True; break
            if np.all(z > 0): break
            xP = x[P]; alpha = 2.0; jb = -1
            for ip in range(len(P)):
                if z[ip] <= 0:
                    t = -xP[ip] / (z[ip] - xP[ip])
                    if alpha > t: alpha = t; jb = ip
            if jb < 0: break
            x[P] = xP + alpha * (z - xP)
            ip = jb
            while True:
                jj = P[ip]; x[jj] = 0.0; F.remove(ip); inP[jj] = False; del P[ip]
                if stats is not None: stats['rem'] = stats.get('rem', 0) + 1
                nxt = [q for q, cj in enumerate(P) if x[cj] <= 0.0]
                if not nxt: break
                ip = nxt[0]
            z = F.solve()
        if done: break
        x[:] = 0.0
        x[P] = z
    return x, P


def x2_proto(D, M, L, warm, stats):
    n = D.shape[1]; G0 = D.T @ D; K = L.T @ L; c = D.T @ M
    f0, _ = gram_nnls_ws(G0, c, 32)
    SSE = np.sum((D @ f0 - M) ** 2)
    state = {'P': None, 'x': None}
    def solve(lam):
        x, P = gram_nnls_ws(G0 + lam * K, c, 92, state['P'] if warm else None, state['x'], stats)
        state['P'], state['x'] = list(P), x.copy()
        return x
    def obj(lam):
        f = solve(lam)
        return abs(np.sum((D @ f - M) ** 2) - 1.02 * SSE) / SSE
    reg = fminbound(obj, 0.0, 10.0, xtol=1e-5, maxfun=300)
    f = solve(reg)
    return f, reg, np.sum((D @ f - M) ** 2) / SSE


if __name__ == '__main__':
    from multicomponent_t2_toolbox_b200.phantom import make_phantom
    nv = int(sys.argv[1]); rm = sys.argv[2]
    ph = make_phantom((16, 16, 4), seed=7); sig = ph['data'].reshape(-1, 32)
    g = O._grids('X2', rm, 'spline', 40., 32, 10., 1000.)
    Dic = O.create_Dic_3D(60, g['T2s'], g['T1s'], 32, 10.0, np.array([105., 130., 155., 180.]), 1000.0)
    for warm in (False, True):
        bad = 0; mx = 0; mxk = 0; st = {}
        for v in range(nv):
            M = sig[v * 5] / sig[v * 5, 0]; D = np.ascontiguousarray(Dic[:, :, v % 4])
            fo, ro, ko = O.nnls_x2(D, M, g['L'], 1.02)
            f, r, k = x2_proto(D, M, g['L'], warm, st)
            if not np.array_equal(f > 0, fo > 0): bad += 1
            mx = max(mx, np.abs(f - fo).max() / np.abs(fo).max()); mxk = max(mxk, abs(k - ko), abs(r - ro) / ro)
        print(rm, 'warm' if warm else 'cold', 'support mismatches %d/%d' % (bad, nv), 'max rel %.2e' % mx, 'max dk/dlam %.2e' % mxk,
              'outer its/voxel %.0f rem/voxel %.0f' % (st.get('outer', 0) / nv, st.get('rem', 0) / nv))
