// met2_device.cuh — per-voxel building blocks shared by the FA and T2 kernels (one warp = one voxel).
#pragma once
#include "met2_nnls.cuh"

namespace met2 {

// SciPy's bounded Brent (scipy/optimize/_optimize.py:_minimize_scalar_bounded), reached from
// algorithms.py:219,280, bayesian_interpolation.py:101 (fminbound) and fa_estimation.py:55 (minimize_scalar 'Bounded').
// Branch-for-branch restatement turned inside out into a resumable state machine so that the
// caller's objective (an NNLS solve) has a single inlined call site:
//     x = B.start(lo, hi, xatol, maxfun);  do { f = objective(x); } while (B.feed(f, x));  result = B.xf
// Warp-uniform: every lane carries the same state.
struct Brent {
    double a, b, fulc, nfc, xf, rat, e, fx, ffulc, fnfc, xm, tol1, tol2, xatol, x;
    int num, maxfun;

    __device__ __forceinline__ double start(double x1, double x2, double xatol_, int maxfun_) {
        const double golden_mean = 0.3819660112501051;   // 0.5 * (3 - sqrt(5))
        a = x1;
        b = x2;
        xatol = xatol_;
        maxfun = maxfun_;
        fulc = a + golden_mean * (b - a);
        nfc = fulc;
        xf = fulc;
        rat = 0.0;
        e = 0.0;
        x = xf;
        num = 0;
        return x;
    }

    // Consume fu = f(x) for the abscissa last handed out.  Returns true and sets xnext if another evaluation is needed.
    __device__ __forceinline__ bool feed(double fu, double& xnext) {
        const double sqrt_eps = 1.4832396974191326e-08;  // sqrt(2.2e-16)
        const double golden_mean = 0.3819660112501051;
        if (num == 0) {
            fx = fu;
            num = 1;
            ffulc = fx;
            fnfc = fx;
        } else {
            ++num;
            if (fu <= fx) {
                if (x >= xf) a = xf; else b = xf;
                fulc = nfc; ffulc = fnfc;
                nfc = xf; fnfc = fx;
                xf = x; fx = fu;
            } else {
                if (x < xf) a = x; else b = x;
                if ((fu <= fnfc) || (nfc == xf)) {
                    fulc = nfc; ffulc = fnfc;
                    nfc = x; fnfc = fu;
                } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
                    fulc = x; ffulc = fu;
                }
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num > 1 && num >= maxfun) return false;
        if (!(fabs(xf - xm) > (tol2 - 0.5 * (b - a)))) return false;
        bool golden = true;
        if (fabs(e) > tol1) {
            golden = false;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                rat = (p + 0.0) / q;
                double xt = xf + rat;
                if (((xt - a) < tol2) || ((b - xt) < tol2)) {
                    double d = xm - xf;
                    double si = (d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : ((d == 0.0) ? 1.0 : d));
                    rat = tol1 * si;
                }
            } else {
                golden = true;
            }
        }
        if (golden) {
            e = (xf >= xm) ? (a - xf) : (b - xf);
            rat = golden_mean * e;
        }
        double si = (rat > 0.0) ? 1.0 : ((rat < 0.0) ? -1.0 : ((rat == 0.0) ? 1.0 : rat));
        x = xf + si * fmax(fabs(rat), tol1);
        xnext = x;
        return true;
    }
};

// c = D^T M into S[W.cc..] (column space).  D is [m][n] row-major in global memory; S[oM + 0..m) is the signal.
template <int NS>
__device__ __forceinline__ void compute_c(const Slots<NS>& W, const double* __restrict__ D, int oM, int m, int n,
                                          int lane) {
    double a0[NS], a1[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) a0[s] = a1[s] = 0.0;
    const int col0 = NS * lane;
    const int nvalid = n - col0;
    const bool vec = ((n & 1) == 0);
    int e = 0;
    if (nvalid > 0) {
#pragma unroll (Unroll<NS>::v)
        for (; e + 1 < m; e += 2) {
            const double m0 = S[oM + e], m1 = S[oM + e + 1];
            double v0[NS], v1[NS];
            ldg_vec<NS>(D + e * n + col0, vec, nvalid, v0);
            ldg_vec<NS>(D + (e + 1) * n + col0, vec, nvalid, v1);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                a0[s] = fma(v0[s], m0, a0[s]);
                a1[s] = fma(v1[s], m1, a1[s]);
            }
        }
        if (e < m) {
            const double m0 = S[oM + e];
            double v0[NS];
            ldg_vec<NS>(D + e * n + col0, vec, nvalid, v0);
#pragma unroll
            for (int s = 0; s < NS; ++s) a0[s] = fma(v0[s], m0, a0[s]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int col = NS * lane + s;
        if (col < n) S[W.cc + col] = a0[s] + a1[s];
    }
    __syncwarp();
}

// Shared-memory variant: D ([m][n], row stride n, n even when NS is even) staged at S[oD..].
template <int NS>
__device__ __forceinline__ void compute_c_sh(const Slots<NS>& W, int oD, int oM, int m, int n, int lane) {
    double a0[NS], a1[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) a0[s] = a1[s] = 0.0;
    const int col0 = NS * lane;
    if (col0 < n) {
        const bool vec = ((n & 1) == 0) && (col0 + NS <= n);
        int e = 0;
#pragma unroll (Unroll<NS>::v)
        for (; e + 1 < m; e += 2) {
            const double m0 = S[oM + e], m1 = S[oM + e + 1];
            double v0[NS], v1[NS];
            if (vec) {
                lds_vec<NS>(oD + e * n + col0, v0);
                lds_vec<NS>(oD + (e + 1) * n + col0, v1);
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    v0[s] = (col0 + s < n) ? S[oD + e * n + col0 + s] : 0.0;
                    v1[s] = (col0 + s < n) ? S[oD + (e + 1) * n + col0 + s] : 0.0;
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                a0[s] = fma(v0[s], m0, a0[s]);
                a1[s] = fma(v1[s], m1, a1[s]);
            }
        }
        if (e < m) {
            const double m0 = S[oM + e];
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (col0 + s < n) a0[s] = fma(S[oD + e * n + col0 + s], m0, a0[s]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int col = NS * lane + s;
        if (col < n) S[W.cc + col] = a0[s] + a1[s];
    }
    __syncwarp();
}

// Shared-memory variant of fit_and_sse: Dt ([n][m]) staged at S[oDt..].
template <int NS, int ME>
__device__ __forceinline__ double fit_and_sse_sh(const Slots<NS>& W, int oDt, int oM, int m, int p, int lane,
                                                 double (&fit)[ME]) {
    double f1[ME];
#pragma unroll
    for (int u = 0; u < ME; ++u) fit[u] = f1[u] = 0.0;
    int k = 0;
#pragma unroll (Unroll<NS>::v)
    for (; k + 1 < p; k += 2) {
        const int d0 = oDt + SI(W.ix, k) * m, d1 = oDt + SI(W.ix, k + 1) * m;
        const double x0 = S[W.xs + k], x1 = S[W.xs + k + 1];
#pragma unroll
        for (int u = 0; u < ME; ++u) {
            int e = lane + 32 * u;
            if (e < m) {
                fit[u] = fma(S[d0 + e], x0, fit[u]);
                f1[u] = fma(S[d1 + e], x1, f1[u]);
            }
        }
    }
    if (k < p) {
        const int d0 = oDt + SI(W.ix, k) * m;
        const double x0 = S[W.xs + k];
#pragma unroll
        for (int u = 0; u < ME; ++u) {
            int e = lane + 32 * u;
            if (e < m) fit[u] = fma(S[d0 + e], x0, fit[u]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < ME; ++u) {
        int e = lane + 32 * u;
        fit[u] += f1[u];
        if (e < m) {
            double dd = fit[u] - S[oM + e];
            s = fma(dd, dd, s);
        }
    }
    return warp_sum(s);
}

// fit = D x for the current positive set (lane-slot u owns echo e = lane + 32 u) and SSE = sum (fit - M)^2.
// Dt is [n][m] row-major (the transposed dictionary) in global memory.
template <int NS, int ME>
__device__ __forceinline__ double fit_and_sse(const Slots<NS>& W, const double* __restrict__ Dt, int oM, int m, int p,
                                              int lane, double (&fit)[ME]) {
    double f1[ME];
#pragma unroll
    for (int u = 0; u < ME; ++u) fit[u] = f1[u] = 0.0;
    int k = 0;
#pragma unroll (Unroll<NS>::v)
    for (; k + 1 < p; k += 2) {
        const double* d0 = Dt + SI(W.ix, k) * m;
        const double* d1 = Dt + SI(W.ix, k + 1) * m;
        const double x0 = S[W.xs + k], x1 = S[W.xs + k + 1];
#pragma unroll
        for (int u = 0; u < ME; ++u) {
            int e = lane + 32 * u;
            if (e < m) {
                fit[u] = fma(__ldg(d0 + e), x0, fit[u]);
                f1[u] = fma(__ldg(d1 + e), x1, f1[u]);
            }
        }
    }
    if (k < p) {
        const double* d0 = Dt + SI(W.ix, k) * m;
        const double x0 = S[W.xs + k];
#pragma unroll
        for (int u = 0; u < ME; ++u) {
            int e = lane + 32 * u;
            if (e < m) fit[u] = fma(__ldg(d0 + e), x0, fit[u]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < ME; ++u) {
        int e = lane + 32 * u;
        fit[u] += f1[u];
        if (e < m) {
            double d = fit[u] - S[oM + e];
            s = fma(d, d, s);
        }
    }
    return warp_sum(s);
}

// One step of iterative refinement of a least-squares solution on its final positive set, with the residual evaluated
// in D-space ("corrected semi-normal equations"):  r = M - D_P x,  g = D_P^T r - lam (K x)_P,  x += T T^T g.
// The Gram-domain solve squares the condition number of D_P (up to ~1e5 for 4-6 EPG columns), which leaves ~1e-6
// (plain) / ~1e-9 (Tikhonov) relative error in the worst voxels; the reference's QR-based Lawson-Hanson does not have
// that loss.  One refinement step brings the solution to the D-space accuracy (~1e-12).  Used for plain NNLS outputs
// (the 1e-6 spectrum tolerance had no margin without it) and for every BayesReg evaluation: the evidence depends on f
// to first order (sum log(1 + erf(U f / sqrt 2))) and its curve is flat enough for a 1e-9 error of f to move Brent's
// lambda by more than 1e-6 in a quarter of the config-4 voxels.  The step is skipped if it would make a
// coefficient non-positive (the active set is the solver's decision, not the refinement's).
// reg: Tikhonov term lam * K with K in 5-band form at S[oKb + d*n + c] = K[c+d-2][c]; S[W.xc..] must hold x in column
// space (zeros outside P), as nnls_gram leaves it.  Scratch: S[W.rs..] (m <= 32 NS doubles), S[W.gs..].
template <int NS, int ME>
__device__ __forceinline__ void refine_on_support(const Slots<NS>& W, const double* __restrict__ Dt, int oM, int m, int p,
                                                  int lane, bool reg, double lam, int oKb, int n) {
    static_assert(ME <= NS, "the residual is staged in a position-space slot");
    double fit[ME];
    (void)fit_and_sse<NS, ME>(W, Dt, oM, m, p, lane, fit);
    const int oR = W.rs;
#pragma unroll
    for (int u = 0; u < ME; ++u) {
        const int e = lane + 32 * u;
        if (e < m) S[oR + e] = S[oM + e] - fit[u];
    }
    __syncwarp();
    double gv[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int i = lane + 32 * t;
        gv[t] = 0.0;
        if (i < p) {
            const int col = SI(W.ix, i);
            const double* d = Dt + col * m;
            double a0 = 0.0, a1 = 0.0;
            int e = 0;
#pragma unroll 1
            for (; e + 1 < m; e += 2) {
                a0 = fma(__ldg(d + e), S[oR + e], a0);
                a1 = fma(__ldg(d + e + 1), S[oR + e + 1], a1);
            }
            if (e < m) a0 = fma(__ldg(d + e), S[oR + e], a0);
            double gi = a0 + a1;
            if (reg) {
                double kx = 0.0;
#pragma unroll
                for (int dd = 0; dd < 5; ++dd) {
                    const int c2 = col + dd - 2;
                    if (c2 >= 0 && c2 < n) kx = fma(S[oKb + dd * n + col], S[W.xc + c2], kx);
                }
                gi = fma(-lam, kx, gi);
            }
            gv[t] = gi;
        }
    }
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int i = lane + 32 * t;
        if (i < p) S[W.gs + i] = gv[t];
    }
    __syncwarp();
    double y[NS], dz[NS];
    tmul_transposed<NS>(W.T, W.gs, p, lane, y);
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int i = lane + 32 * t;
        if (i < p) S[W.rs + i] = y[t];
    }
    __syncwarp();
    tmul<NS>(W.T, W.rs, p, lane, dz);
    bool ok = true;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int i = lane + 32 * t;
        if (i < p && !(S[W.xs + i] + dz[t] > 0.0)) ok = false;
    }
    if (__all_sync(FULL_MASK, ok)) {
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int i = lane + 32 * t;
            if (i < p) {
                const double xn = S[W.xs + i] + dz[t];
                S[W.xs + i] = xn;
                S[W.xc + SI(W.ix, i)] = xn;
            }
        }
    }
    __syncwarp();
}

// sum_r ((L x)_r)^2 with L in 5-band row form S[oLb + d*n + r] = L[r][r+d-2]; x in column space (S[W.xc..]).
template <int NS>
__device__ __forceinline__ double reg_norm2(const Slots<NS>& W, int oLb, int n, int lane) {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int r = NS * lane + t;
        if (r < n) {
            double acc = 0.0;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                int c2 = r + d - 2;
                if (c2 >= 0 && c2 < n) acc = fma(S[oLb + d * n + r], S[W.xc + c2], acc);
            }
            s = fma(acc, acc, s);
        }
    }
    return warp_sum(s);
}

// Load the raw signal of voxel v into S[oM..]; returns 0 if it is to be fitted, else the status bits
// (fa_estimation.py:100 / motor...:124: mask > 0 and sum(M) > 0; NaN/Inf: algorithms.py:56 raises).
template <int ME>
__device__ __forceinline__ unsigned load_signal(const double* __restrict__ sig, long long v, int m, int oM, int lane) {
    double s = 0.0;
    bool bad = false;
#pragma unroll
    for (int u = 0; u < ME; ++u) {
        int e = lane + 32 * u;
        if (e < m) {
            double xv = sig[v * m + e];
            S[oM + e] = xv;
            s += xv;
            if (!isfinite(xv)) bad = true;
        }
    }
    s = warp_sum(s);
    bad = __any_sync(FULL_MASK, bad);
    __syncwarp();
    if (bad) return 2u | 1u;       // MET2_ST_NONFINITE | MET2_ST_SKIPPED
    if (!(s > 0.0)) return 1u;     // MET2_ST_SKIPPED
    return 0u;
}

}  // namespace met2
