"""Design prototype (not product, not oracle): Gram-domain Lawson-Hanson variants vs SciPy's NNLS.

Measures support agreement and spectrum error of the arithmetic planned for the CUDA kernel:
  variant 'chol' : Cholesky factor R of G_PP with triangular solves
  variant 'hinv' : explicit inverse H = G_PP^-1 maintained by bordering / rank-1 downdates
  variant 'tinv' : explicit inverse triangular factor T = R^-1
"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo')
import met2_oracle as O


def gram_nnls(G, c, m_rows, variant='hinv', refine=0, itmax=None, stats=None):
    n = G.shape[0]
    if itmax is None:
        itmax = 3 * n
    x = np.zeros(n)
    P = []
    inP = np.zeros(n, bool)
    H = np.zeros((0, 0))
    z = np.zeros(0)
    it = 0
    w = c.copy()
    while True:
        p = len(P)
        if p >= n or p >= m_rows:
            break
        # gradient
        w = c - G[:, P] @ x[P] if p else c.copy()
        w[inP] = 0.0
        accepted = False
        while True:
            wz = np.where(inP, -np.inf, w)
            j = int(np.argmax(wz))
            if not (wz[j] > 0):
                break
            g = G[P, j]
            if variant == 'hinv':
                u = H @ g
                s = G[j, j] - g @ u
                if not (s > 0):
                    w[j] = 0; inP_tmp = True
                    # reject
                    wz[j] = -np.inf
                    w[j] = -0.0
                    w = np.where(np.arange(n) == j, 0.0, w)
                    # mark as rejected by zeroing
                    continue
                t = (c[j] - g @ z) / s        # new coefficient
                if not (t > 0):
                    w[j] = 0.0
                    continue
                Hn = np.empty((p + 1, p + 1))
                Hn[:p, :p] = H + np.outer(u, u) / s
                Hn[:p, p] = -u / s
                Hn[p, :p] = -u / s
                Hn[p, p] = 1.0 / s
                H = Hn
                z = np.concatenate((z - t * u, [t]))
            P.append(j); inP[j] = True
            accepted = True
            break
        if not accepted:
            break
        if stats is not None:
            stats['outer'] = stats.get('outer', 0) + 1
            stats['pmax'] = max(stats.get('pmax', 0), len(P))
        done = False
        while True:
            it += 1
            if it > itmax:
                done = True; break
            for _ in range(refine):
                res = c[P] - G[np.ix_(P, P)] @ z
                z = z + H @ res
            if np.all(z > 0):
                break
            xP = x[P]
            neg = z <= 0
            tt = np.where(neg, -xP / (z - xP), np.inf)
            alpha = 2.0; jb = -1
            for ip in range(len(P)):
                if neg[ip] and alpha > tt[ip]:
                    alpha = tt[ip]; jb = ip
            if jb < 0:
                break
            x[P] = xP + alpha * (z - xP)
            ip = jb
            while True:
                jj = P[ip]
                x[jj] = 0.0
                # remove ip from H
                h = H[:, ip].copy(); hkk = h[ip]
                keep = [q for q in range(len(P)) if q != ip]
                H = H[np.ix_(keep, keep)] - np.outer(h[keep], h[keep]) / hkk
                inP[jj] = False
                del P[ip]
                if stats is not None:
                    stats['rem'] = stats.get('rem', 0) + 1
                nxt = [q for q, cj in enumerate(P) if x[cj] <= 0.0]
                if not nxt:
                    break
                ip = nxt[0]
            z = H @ c[P]
        if done:
            break
        x[P] = z
    return x, P


if __name__ == '__main__':
    from multicomponent_t2_toolbox_b200.phantom import make_phantom
    refine = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    ph = make_phantom((16, 16, 4), seed=7)
    sig = ph['data'].reshape(-1, 32)
    T2s = np.logspace(1, np.log10(2000), 60); T1s = 1000 * np.ones(60)
    al = np.array([100., 125., 150., 180.])
    Dic = O.create_Dic_3D(60, T2s, T1s, 32, 10.0, al, 1000.0)
    for Lname in ['I', 'L2', 'InvT2']:
        L = O._grids('X2', Lname, 'spline', 40., 32, 10., 1000.)['L']
        K = L.T @ L
        for lam in [0, 1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1, 1, 3.8197]:
            if lam == 0 and Lname != 'I':
                continue
            bad = 0; mx = 0; cnt = 0; st = {}
            for v in range(nv):
                M = sig[v * 3] / sig[v * 3, 0]
                ia = v % 4
                D = np.ascontiguousarray(Dic[:, :, ia])
                if lam == 0:
                    A = D; b = M; mrows = 32
                else:
                    A = np.concatenate((D, np.sqrt(lam) * L)); b = np.concatenate((M, np.zeros(60))); mrows = 92
                x0, r0 = O.nnls(A, b)
                G = D.T @ D + lam * K
                c = D.T @ M
                x1, P = gram_nnls(G, c, mrows, refine=refine, stats=st)
                cnt += 1
                if not np.array_equal(x0 > 0, x1 > 0):
                    bad += 1
                mx = max(mx, np.abs(x0 - x1).max() / np.abs(x0).max())
            print(Lname, 'lam=%g' % lam, 'support mismatches %d/%d' % (bad, cnt), 'max rel %.2e' % mx,
                  'outer/solve %.1f pmax %d rem/solve %.1f' % (st.get('outer', 0) / cnt, st.get('pmax', 0), st.get('rem', 0) / cnt))
