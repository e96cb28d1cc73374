// met2_device.cuh — per-voxel building blocks shared by the FA and T2 kernels (one warp = one voxel).
#pragma once
#include "met2_nnls.cuh"

namespace met2 {

// SciPy's bounded Brent (scipy/optimize/_optimize.py:_minimize_scalar_bounded), reached from
// algorithms.py:219,280, bayesian_interpolation.py:101 (fminbound) and fa_estimation.py:55 (minimize_scalar 'Bounded').
// Branch-for-branch restatement; the objective is evaluated warp-uniformly.
template <class F>
__device__ __forceinline__ double brent_bounded(F&& func, double x1, double x2, double xatol, int maxfun, double& fval,
                                                int& nfev) {
    const double sqrt_eps = 1.4832396974191326e-08;  // sqrt(2.2e-16)
    const double golden_mean = 0.3819660112501051;   // 0.5 * (3 - sqrt(5))
    double a = x1, b = x2;
    double fulc = a + golden_mean * (b - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = func(x);
    int num = 1;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
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
                x = xf + rat;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
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
        double fu = func(x);
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
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= maxfun) break;
    }
    fval = fx;
    nfev = num;
    return xf;
}

// c = D^T M into W.cc (column space).  D is [m][n] row-major; ms[0..m) is the signal in shared memory.
template <int NS>
__device__ __forceinline__ void compute_c(const NnlsWork<NS>& W, const double* __restrict__ D, const double* ms, int m,
                                          int n, int lane) {
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.0;
    for (int e = 0; e < m; ++e) {
        double me = ms[e];
        const double* drow = D + e * n;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int col = lane + 32 * s;
            if (col < n) acc[s] = fma(drow[col], me, acc[s]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int col = lane + 32 * s;
        if (col < n) W.cc[col] = acc[s];
    }
    __syncwarp();
}

// fit = D x for the current positive set (lane-slot u owns echo e = lane + 32 u) and SSE = sum (fit - M)^2.
// Dt is [n][m] row-major (the transposed dictionary).
template <int NS, int ME>
__device__ __forceinline__ double fit_and_sse(const NnlsWork<NS>& W, const double* __restrict__ Dt, const double* ms,
                                              int m, int p, int lane, double (&fit)[ME]) {
#pragma unroll
    for (int u = 0; u < ME; ++u) fit[u] = 0.0;
    for (int k = 0; k < p; ++k) {
        const double* drow = Dt + W.idx[k] * m;
        double xk = W.xs[k];
#pragma unroll
        for (int u = 0; u < ME; ++u) {
            int e = lane + 32 * u;
            if (e < m) fit[u] = fma(drow[e], xk, fit[u]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < ME; ++u) {
        int e = lane + 32 * u;
        if (e < m) {
            double d = fit[u] - ms[e];
            s = fma(d, d, s);
        }
    }
    return warp_sum(s);
}

// sum_r ((L x)_r)^2 with L in 5-band row form lb[d*n + r] = L[r][r+d-2]; x in column space (W.xc).
template <int NS>
__device__ __forceinline__ double reg_norm2(const NnlsWork<NS>& W, const double* __restrict__ lb, int n, int lane) {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int r = lane + 32 * t;
        if (r < n) {
            double acc = 0.0;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                int c2 = r + d - 2;
                if (c2 >= 0 && c2 < n) acc = fma(lb[d * n + r], W.xc[c2], acc);
            }
            s = fma(acc, acc, s);
        }
    }
    return warp_sum(s);
}

}  // namespace met2
