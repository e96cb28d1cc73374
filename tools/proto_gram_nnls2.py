"""Design prototype #2: Gram-domain LH with (a) fresh Cholesky solves, (b) T = R^-1 updates, +/- CSNE refinement."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo')
import met2_oracle as O


class TInv:
    """Upper-triangular T with G_PP^-1 = T T^T, plus y = T^T c_P."""
    def __init__(self):
        self.T = np.zeros((0, 0)); self.y = np.zeros(0)

    def try_append(self, g, gjj, cj):
        r = self.T.T @ g
        rho2 = gjj - r @ r
        if not (rho2 > 0):
            return None
        rho = np.sqrt(rho2)
        ynew = (cj - r @ self.y) / rho
        return r, rho, ynew

    def append(self, r, rho, ynew):
        p = self.T.shape[0]
        Tn = np.zeros((p + 1, p + 1))
        Tn[:p, :p] = self.T
        Tn[:p, p] = -(self.T @ r) / rho
        Tn[p, p] = 1.0 / rho
        self.T = Tn
        self.y = np.concatenate((self.y, [ynew]))

    def remove(self, k):
        T = self.T; y = self.y
        p = T.shape[0]
        for q in range(k, p - 1):
            a, b = T[k, q], T[k, q + 1]
            nu = np.hypot(a, b)
            if nu == 0:
                c, s = 1.0, 0.0
            else:
                c, s = b / nu, a / nu
            colq = T[:, q].copy(); colq1 = T[:, q + 1].copy()
            T[:, q] = c * colq - s * colq1
            T[:, q + 1] = s * colq + c * colq1
            yq, yq1 = y[q], y[q + 1]
            y[q] = c * yq - s * yq1
            y[q + 1] = s * yq + c * yq1
        keep = [i for i in range(p) if i != k]
        self.T = T[np.ix_(keep, list(range(p - 1)))]
        self.y = y[:p - 1]

    def solve(self):
        return self.T @ self.y


def gram_nnls(G, c, m_rows, D=None, M=None, lamK=None, variant='tinv', refine=0, itmax=None, stats=None):
    n = G.shape[0]
    itmax = 3 * n if itmax is None else itmax
    x = np.zeros(n); P = []; inP = np.zeros(n, bool)
    F = TInv()
    it = 0

    def solve_z():
        if variant == 'tinv':
            z = F.solve()
        else:
            Gpp = G[np.ix_(P, P)]
            Lc = np.linalg.cholesky(Gpp)
            z = np.linalg.solve(Lc.T, np.linalg.solve(Lc, c[P]))
        for _ in range(refine):
            # corrected semi-normal equations: residual in D-space
            r = M - D[:, P] @ z
            gres = D[:, P].T @ r
            if lamK is not None:
                gres = gres - lamK[np.ix_(P, P)] @ z
            if variant == 'tinv':
                z = z + F.T @ (F.T.T @ gres)
            else:
                z = z + np.linalg.solve(Lc.T, np.linalg.solve(Lc, gres))
        return z

    while True:
        p = len(P)
        if p >= n or p >= m_rows:
            break
        if refine and p:
            r = M - D[:, P] @ x[P]
            w = D.T @ r
            if lamK is not None:
                w = w - lamK[:, P] @ x[P]
        else:
            w = c - G[:, P] @ x[P] if p else c.copy()
        w[inP] = 0.0
        accepted = False
        rejected = np.zeros(n, bool)
        while True:
            wz = np.where(inP | rejected, -np.inf, w)
            j = int(np.argmax(wz))
            if not (wz[j] > 0):
                break
            res = F.try_append(G[P, j], G[j, j], c[j])
            if res is None or not (res[2] > 0):
                rejected[j] = True
                if stats is not None: stats['rej'] = stats.get('rej', 0) + 1
                continue
            F.append(*res)
            P.append(j); inP[j] = True
            accepted = True
            break
        if not accepted:
            break
        if stats is not None:
            stats['outer'] = stats.get('outer', 0) + 1
        z = solve_z()
        done = False
        while True:
            it += 1
            if it > itmax:
                done = True; break
            if np.all(z > 0):
                break
            xP = x[P]
            alpha = 2.0; jb = -1
            for ip in range(len(P)):
                if z[ip] <= 0:
                    t = -xP[ip] / (z[ip] - xP[ip])
                    if alpha > t:
                        alpha = t; jb = ip
            if jb < 0:
                break
            x[P] = xP + alpha * (z - xP)
            ip = jb
            while True:
                jj = P[ip]
                x[jj] = 0.0
                F.remove(ip)
                inP[jj] = False
                del P[ip]
                nxt = [q for q, cj in enumerate(P) if x[cj] <= 0.0]
                if not nxt:
                    break
                ip = nxt[0]
            z = solve_z()
        if done:
            break
        x[P] = z
    return x, P


if __name__ == '__main__':
    from multicomponent_t2_toolbox_b200.phantom import make_phantom
    variant = sys.argv[1]; refine = int(sys.argv[2]); nv = int(sys.argv[3])
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
                D = np.ascontiguousarray(Dic[:, :, v % 4])
                if lam == 0:
                    A = D; b = M; mrows = 32; lamK = None
                else:
                    A = np.concatenate((D, np.sqrt(lam) * L)); b = np.concatenate((M, np.zeros(60))); mrows = 92; lamK = lam * K
                x0, r0 = O.nnls(A, b)
                G = D.T @ D + lam * K
                c = D.T @ M
                x1, P = gram_nnls(G, c, mrows, D, M, lamK, variant=variant, refine=refine, stats=st)
                cnt += 1
                if not np.array_equal(x0 > 0, x1 > 0):
                    bad += 1
                mx = max(mx, np.abs(x0 - x1).max() / np.abs(x0).max())
            print(Lname, 'lam=%g' % lam, 'support mismatches %d/%d' % (bad, cnt), 'max rel %.2e' % mx,
                  'outer/solve %.1f rej %d' % (st.get('outer', 0) / cnt, st.get('rej', 0)))
