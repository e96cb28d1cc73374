// met2_nnls.cuh — warp-per-voxel Lawson-Hanson NNLS in the Gram domain (FP64, sm_100a).
//
// Takes over scipy's Lawson-Hanson routine reached from intravoxel_algorithms/algorithms.py:55-82 (nnls) and, with the
// Tikhonov term, algorithms.py:262-269 (A = [D; sqrt(lambda) L]).  Same pivot rule and control flow as nnls.f:
//   outer: w = A^T(b - A x); j = argmax_{j in Z} w_j (first max), stop if w_j <= 0; accept j unless the column is
//          numerically dependent on the positive set or its new coefficient is <= 0 (then w_j := 0 and re-pick);
//   inner: iter++ (stop at itmax = 3n, current x kept); if all z_P > 0 leave; alpha = min_{z_i<=0} x_i/(x_i - z_i)
//          (first min); x += alpha (z - x); move the blocking index and every x_i <= 0 out of P; re-solve.
// What differs is the arithmetic: instead of Householder/Givens on the m x n matrix we keep, per warp, the inverse
// Cholesky factor T (upper triangular, packed) of the positive-set Gram matrix, (G + lambda K)_PP^-1 = T T^T:
//   append column j :  r = T^T g_Pj ; rho^2 = G_jj - r.r ; new column of T = [-T r ; 1] / rho ; y_new = (c_j - r.y)/rho
//   remove position k: Givens rotations on column pairs of T that empty row k (parameters from a prefix sum of squares)
//   solve            : z = T y   with y = T^T c_P
// Every step is a (triangular) matrix-vector product spread over the 32 lanes; no serial substitution chains.
// Measured against SciPy on the design-prototype (tools/proto_gram_nnls2.py): identical supports, <= 1e-9 relative.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace met2 {

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

__host__ __device__ __forceinline__ size_t align_up256(size_t x) { return (x + 255) & ~(size_t)255; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ void warp_sum2(double& a, double& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ta = __shfl_xor_sync(FULL_MASK, a, o);
        double tb = __shfl_xor_sync(FULL_MASK, b, o);
        a += ta;
        b += tb;
    }
}

// Index of the largest strictly positive value over the warp (lowest index on ties); -1 if no lane has v > 0.
// Positive doubles order like their bit patterns, so two 32-bit REDUX max passes find the maximum.
__device__ __forceinline__ int warp_argmax_pos(double v, int index, double& vmax) {
    unsigned long long key = (v > 0.0) ? (unsigned long long)__double_as_longlong(v) : 0ull;
    unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
    unsigned mlo = __reduce_max_sync(FULL_MASK, (hi == mhi) ? lo : 0u);
    if ((mhi | mlo) == 0u) return -1;
    bool win = (key != 0ull) && (hi == mhi) && (lo == mlo);
    unsigned best = __reduce_min_sync(FULL_MASK, win ? (unsigned)index : 0xffffffffu);
    vmax = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    return (int)best;
}

// Index of the smallest value among lanes with index >= 0 and v >= 0 (lowest index on ties); -1 if none.
__device__ __forceinline__ int warp_argmin_nonneg(double v, int index, double& vmin) {
    bool valid = (index >= 0) && (v >= 0.0) && (v < 1.0e308);
    unsigned long long key = valid ? (unsigned long long)__double_as_longlong(v + 0.0) : ~0ull;
    unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    unsigned mhi = __reduce_min_sync(FULL_MASK, hi);
    unsigned mlo = __reduce_min_sync(FULL_MASK, (hi == mhi) ? lo : 0xffffffffu);
    if (mhi == 0xffffffffu && mlo == 0xffffffffu) return -1;
    bool win = valid && (hi == mhi) && (lo == mlo);
    unsigned best = __reduce_min_sync(FULL_MASK, win ? (unsigned)index : 0xffffffffu);
    vmin = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    return (int)best;
}

// Per-warp shared-memory workspace.  NS = slots per lane; all vectors have 32*NS entries.
template <int NS>
struct NnlsWork {
    double* T;    // packed upper triangular, column-major: T(k, i) at T[tri(i) + k], k <= i < pmax
    double* gs;   // position-space scratch (gathered Gram column / c_P / rotation cosines)
    double* rs;   // position-space scratch (r, y / rotation sines)
    double* xs;   // current feasible x in position space
    double* cc;   // c = A^T b in column space
    double* xc;   // x in column space (zeros outside P); kept current only when the Tikhonov term is present
    int* idx;     // position -> column
    static constexpr int LEN = 32 * NS;
    __host__ __device__ static size_t bytes(int pmax) {
        return sizeof(double) * (size_t)((pmax * (pmax + 1)) / 2 + 5 * LEN) + sizeof(int) * (size_t)LEN;
    }
    __device__ void carve(unsigned char* base, int pmax) {
        T = reinterpret_cast<double*>(base);
        gs = T + (pmax * (pmax + 1)) / 2;
        rs = gs + LEN;
        xs = rs + LEN;
        cc = xs + LEN;
        xc = cc + LEN;
        idx = reinterpret_cast<int*>(xc + LEN);
    }
};

// out[t] (position i = lane + 32 t) = sum_{k <= i} T(k, i) * v[k]   — column dot products, T^T v
template <int NS>
__device__ __forceinline__ void tmul_transposed(const double* __restrict__ T, const double* __restrict__ v, int p,
                                                int lane, double (&out)[NS]) {
    int base[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        out[t] = 0.0;
        base[t] = tri(lane + 32 * t);
    }
    for (int k = 0; k < p; ++k) {
        double vk = v[k];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            int i = lane + 32 * t;
            if (i < p && k <= i) out[t] = fma(T[base[t] + k], vk, out[t]);
        }
    }
}

// out[t] (position k = lane + 32 t) = sum_{i >= k, i < p} T(k, i) * v[i]   — row dot products, T v
template <int NS>
__device__ __forceinline__ void tmul(const double* __restrict__ T, const double* __restrict__ v, int p, int lane,
                                     double (&out)[NS]) {
#pragma unroll
    for (int t = 0; t < NS; ++t) out[t] = 0.0;
    for (int i = 0; i < p; ++i) {
        double vi = v[i];
        int ti = tri(i);
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            int k = lane + 32 * t;
            if (k <= i) out[t] = fma(T[ti + k], vi, out[t]);
        }
    }
}

// Inclusive prefix sum over positions (lane + 32 t ordering).
template <int NS>
__device__ __forceinline__ void warp_scan_positions(double (&v)[NS], int lane) {
    double carry = 0.0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        double s = v[t];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u = __shfl_up_sync(FULL_MASK, s, o);
            if (lane >= o) s += u;
        }
        s += carry;
        v[t] = s;
        carry = __shfl_sync(FULL_MASK, s, 31);
    }
}

// Delete position k of the positive set (p -> p-1): re-triangularise T, shift idx and x.
template <int NS>
__device__ __forceinline__ void remove_position(const NnlsWork<NS>& W, int k, int& p, int lane, double (&x)[NS]) {
    double* T = W.T;
    // rotation parameters from the prefix sums of squares of row k
    double tau[NS], S[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int q = lane + 32 * t;
        tau[t] = (q >= k && q < p) ? T[tri(q) + k] : 0.0;
        S[t] = tau[t] * tau[t];
    }
    warp_scan_positions<NS>(S, lane);
    // nu_q = sqrt(S_q); step q (k <= q <= p-2) needs c_q = tau_{q+1}/nu_{q+1}, s_q = nu_q/nu_{q+1}
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int q = lane + 32 * t;
        if (q >= k && q < p) {
            double nu = sqrt(S[t]);
            W.rs[q] = nu;          // nu_q
            W.gs[q] = tau[t];      // tau_q
        }
    }
    __syncwarp();
    double cq[NS], sq[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int q = lane + 32 * t;
        cq[t] = 0.0;
        sq[t] = 1.0;
        if (q >= k && q + 1 < p) {
            double nu1 = W.rs[q + 1];
            double inv = 1.0 / nu1;
            cq[t] = W.gs[q + 1] * inv;
            sq[t] = W.rs[q] * inv;
        }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int q = lane + 32 * t;
        if (q >= k && q + 1 < p) {
            W.gs[q] = cq[t];
            W.rs[q] = sq[t];
        }
    }
    // shift x through xs (xs is rewritten with the shifted vector below)
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int i = lane + 32 * t;
        if (i < p) W.xs[i] = x[t];
    }
    __syncwarp();
    int idx_next[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int i = lane + 32 * t;
        idx_next[t] = (i >= k && i + 1 < p) ? W.idx[i + 1] : -1;
        if (i >= k) x[t] = (i + 1 < p) ? W.xs[i + 1] : 0.0;
    }
    // row sweep: lane-slot owns old row rr (rr != k); carry starts as T(rr, k)
    double carry[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int rr = lane + 32 * t;
        carry[t] = (rr < k) ? T[tri(k) + rr] : 0.0;
    }
    for (int q = k; q + 1 < p; ++q) {
        double c = W.gs[q], s = W.rs[q];
        int tq1 = tri(q + 1), tq = tri(q);
        double nv[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            int rr = lane + 32 * t;
            double b = (rr <= q + 1 && rr != k) ? T[tq1 + rr] : 0.0;
            nv[t] = s * b - c * carry[t];
            carry[t] = fma(s, carry[t], c * b);
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            int rr = lane + 32 * t;
            if (rr != k && rr <= q + 1) {
                int rn = rr - (rr > k ? 1 : 0);
                T[tq + rn] = nv[t];
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int i = lane + 32 * t;
        if (idx_next[t] >= 0) W.idx[i] = idx_next[t];
    }
    --p;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        int i = lane + 32 * t;
        if (i < p) W.xs[i] = x[t];
    }
    __syncwarp();
}

// Gram-domain Lawson-Hanson.  On entry W.cc[0..n) holds c = A^T b (visible to the whole warp).
// G: n x n row-major Gram matrix of the unregularised dictionary (shared or global memory).
// REG: add lam * K, K given in 5-band form kb[d*n + c] = K[c+d-2][c].
// On exit W.idx[0..p) / W.xs[0..p) hold the positive set and its coefficients; returns p.  status gets bit 0 on itmax.
template <int NS, bool REG>
__device__ __noinline__ int nnls_gram(const NnlsWork<NS>& W, const double* __restrict__ G,
                                      const double* __restrict__ kb, double lam, int n, int mrows, int lane,
                                      int& status) {
    const int itmax = 3 * n;
    double creg[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int col = lane + 32 * s;
        creg[s] = (col < n) ? W.cc[col] : 0.0;
        if (REG && col < n) W.xc[col] = 0.0;
    }
    unsigned inP = 0u;
    double x[NS], y[NS], z[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) x[t] = y[t] = z[t] = 0.0;
    int p = 0, iter = 0;
    __syncwarp();
    while (true) {
        if (p >= n || p >= mrows) break;
        // ---- dual vector w = c - (G + lam K) x on the zero set
        double w[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) w[s] = creg[s];
        if (p > 0) {
            for (int k = 0; k < p; ++k) {
                const double* grow = G + W.idx[k] * n;
                double xk = W.xs[k];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    int col = lane + 32 * s;
                    if (col < n) w[s] = fma(-xk, grow[col], w[s]);
                }
            }
            if (REG) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    int col = lane + 32 * s;
                    if (col < n) {
                        double acc = 0.0;
#pragma unroll
                        for (int d = 0; d < 5; ++d) {
                            int c2 = col + d - 2;
                            if (c2 >= 0 && c2 < n) acc = fma(kb[d * n + col], W.xc[c2], acc);
                        }
                        w[s] = fma(-lam, acc, w[s]);
                    }
                }
            }
        }
        // ---- pick the entering column
        unsigned rejected = 0u;
        bool accepted = false;
        while (true) {
            double bv = 0.0;
            int bj = -1;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                int col = lane + 32 * s;
                if (col < n && !((inP >> s) & 1u) && !((rejected >> s) & 1u) && w[s] > bv) {
                    bv = w[s];
                    bj = col;
                }
            }
            double wmax;
            int j = warp_argmax_pos(bv, bj, wmax);
            if (j < 0) break;
            double gjj = G[j * n + j];
            if (REG) gjj = fma(lam, kb[2 * n + j], gjj);
            const double* gcol = G + j * n;   // G is symmetric: column j == row j
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p) {
                    int r = W.idx[i];
                    double gv = gcol[r];
                    if (REG) {
                        int d = r - j + 2;
                        if (d >= 0 && d <= 4) gv = fma(lam, kb[d * n + j], gv);
                    }
                    W.gs[i] = gv;
                }
            }
            __syncwarp();
            double r[NS];
            tmul_transposed<NS>(W.T, W.gs, p, lane, r);
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p) {
                    s1 = fma(r[t], r[t], s1);
                    s2 = fma(r[t], y[t], s2);
                }
            }
            warp_sum2(s1, s2);
            double rho2 = gjj - s1;
            double cj = W.cc[j];
            double rho = sqrt(rho2);
            double rinv = 1.0 / rho;
            double ynew = (cj - s2) * rinv;
            // dependence test of nnls.f (unorm + |a_new|*0.01 > unorm) and its "ztest > 0" in Gram-domain form
            bool ok = (rho2 > 0.0) && (sqrt(s1) + rho * 0.01 > sqrt(s1)) && (ynew > 0.0);
            if (!ok) {
                if ((j & 31) == lane) rejected |= 1u << (j >> 5);
                __syncwarp();
                continue;
            }
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p) W.rs[i] = r[t];
            }
            __syncwarp();
            double acc[NS];
            tmul<NS>(W.T, W.rs, p, lane, acc);
            const int tp = tri(p);
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int k = lane + 32 * t;
                if (k < p) {
                    double tk = -acc[t] * rinv;
                    W.T[tp + k] = tk;
                    z[t] = fma(ynew, tk, z[t]);
                } else if (k == p) {
                    W.T[tp + p] = rinv;
                    z[t] = ynew * rinv;
                    y[t] = ynew;
                    W.idx[p] = j;
                    W.xs[p] = 0.0;   // x of the entering column is 0 until the solve is accepted
                }
            }
            if ((j & 31) == lane) inP |= 1u << (j >> 5);
            ++p;
            accepted = true;
            __syncwarp();
            break;
        }
        if (!accepted) break;
        // ---- secondary loop
        bool stop = false;
        while (true) {
            ++iter;
            if (iter > itmax) {
                status |= 1;
                stop = true;
                break;
            }
            bool neg = false;
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p && z[t] <= 0.0) neg = true;
            }
            if (!__any_sync(FULL_MASK, neg)) break;
            double bt = 2.0;
            int bi = -1;
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p && z[t] <= 0.0) {
                    double tt = x[t] / (x[t] - z[t]);
                    if (tt < bt) {
                        bt = tt;
                        bi = i;
                    }
                }
            }
            double alpha;
            int jb = warp_argmin_nonneg(bt, bi, alpha);
            if (jb < 0) break;
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p) x[t] = x[t] + alpha * (z[t] - x[t]);
            }
            int k = jb;
            while (true) {
                int colk = W.idx[k];
                if ((colk & 31) == lane) inP &= ~(1u << (colk >> 5));
                if (REG && lane == 0) W.xc[colk] = 0.0;
                __syncwarp();
                remove_position<NS>(W, k, p, lane, x);
                int q = 0x7fffffff;
#pragma unroll
                for (int t = NS - 1; t >= 0; --t) {
                    int i = lane + 32 * t;
                    if (i < p && x[t] <= 0.0) q = i;
                }
                q = (int)__reduce_min_sync(FULL_MASK, (unsigned)q);
                if (q == 0x7fffffff) break;
                k = q;
            }
            // fresh y = T^T c_P and z = T y
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p) W.gs[i] = W.cc[W.idx[i]];
            }
            __syncwarp();
            tmul_transposed<NS>(W.T, W.gs, p, lane, y);
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                int i = lane + 32 * t;
                if (i < p) W.rs[i] = y[t];
                else y[t] = 0.0;
            }
            __syncwarp();
            tmul<NS>(W.T, W.rs, p, lane, z);
            __syncwarp();
        }
        if (stop) break;
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            int i = lane + 32 * t;
            if (i < p) {
                x[t] = z[t];
                W.xs[i] = z[t];
                if (REG) W.xc[W.idx[i]] = z[t];
            }
        }
        __syncwarp();
    }
    // xs mirrors x on every path: after an accepted solve it was just written, after an itmax stop
    // remove_position left the interpolated x there.
    if (REG) {
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            int i = lane + 32 * t;
            if (i < p) W.xc[W.idx[i]] = x[t];
        }
    }
    __syncwarp();
    return p;
}

}  // namespace met2
