// met2_basis.cu — reduced echo basis of the EPG dictionary (run once per reconstruction, after met2_epg_dictionary).
#include <cmath>

#include "met2_host.h"
#include "met2_nnls.cuh"

namespace met2 {

// ---------------------------------------------------------------------------------------------- reduced echo basis
// The EPG decay curves of one flip angle are numerically low-rank: the singular values of D_a (nTE x nT2) fall by ~5x
// per index and reach the rounding level of D's own entries (1e-17 sigma_1) at index ~20 whatever nTE is.  So
//     D_a = U_a C_a + E,   U_a: nTE x R orthonormal,   C_a = U_a^T D_a: R x nT2,   |E| ~ 1e-16 |D_a|   (R = 24)
// and every least-squares problem against D_a is, to the rounding of D_a, a problem in R dimensions:
//     |D x - b|^2 = |C x - U^T b|^2 + |b - U U^T b|^2.
// The echo-space Tikhonov kernels (met2_t2_echo.cu) work on (C, U^T b) — factors of 24 x 24 instead of 32 x 32 (or
// 48 x 48 at nTE = 48).  ANY orthonormal basis of a subspace that contains range(D_a) to rounding will do, so this
// kernel takes the cheapest rank-revealing one: Gram-Schmidt with column pivoting on D_a (the column of largest
// residual norm next), every new direction re-orthogonalised twice against the earlier ones, C = U^T D recomputed from
// the dictionary itself at the end, and the residual max_j |d_j - U C_j| / max_j |d_j| reported per angle so that the
// caller can verify the reduction instead of trusting it (batched.Dictionary: <= 1e-15).
// One warp per angle; lane = dictionary column (nT2 <= 128: four column slots), no collective inside the loops except
// one arg-max per step.  Shared memory: W (residual dictionary, nTE x ldw) and Q (R x nTE).
__global__ void __launch_bounds__(32) echo_basis_kernel(const double* __restrict__ dic, int m, int n, int R,
                                                        double* __restrict__ basis, double* __restrict__ coef,
                                                        double* __restrict__ tail) {
    const int lane = threadIdx.x, a = blockIdx.x;
    const int ldw = n | 1;                          // odd stride: conflict-free column walks
    const int oW = 0, oQ = m * ldw, oT = oQ + R * m;   // oT: m doubles of scratch (the pivot column)
    const double* D = dic + (size_t)a * m * n;
    for (int i = lane; i < m * n; i += 32) {
        const int r = i / n;
        S[oW + r * ldw + (i - r * n)] = D[i];
    }
    __syncwarp();
    double dmax2 = 0.0;
    for (int k = 0; k < R; ++k) {
        // ---- pivot: column of largest residual norm (first maximum)
        double bv = 0.0;
        int bj = -1;
        for (int j = lane; j < n; j += 32) {
            double s2 = 0.0;
            for (int r = 0; r < m; ++r) s2 = fma(S[oW + r * ldw + j], S[oW + r * ldw + j], s2);
            if (s2 > bv) {
                bv = s2;
                bj = j;
            }
        }
        const int piv = warp_argmax_pos(bv, bj);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bv = fmax(bv, __shfl_xor_sync(FULL_MASK, bv, o));
        const double bv_piv = bv;
        if (k == 0) dmax2 = bv;
        // residual below the rounding of D's own entries (3e-16 of the largest column; exactly zero if nTE < R): the
        // dictionary is captured, the remaining directions are zero (they contribute nothing to U C or U^T b)
        if (piv < 0 || bv_piv < 1e-31 * dmax2) {
            for (int r = lane; r < m; r += 32) S[oQ + k * m + r] = 0.0;
            __syncwarp();
            continue;
        }
        for (int r = lane; r < m; r += 32) S[oT + r] = S[oW + r * ldw + piv];
        __syncwarp();
        // ---- orthogonalise the pivot column against the earlier directions (twice), normalise
        // (deflation already made it orthogonal once; repeated until the norm stops collapsing — "twice is enough"
        // unless the column is numerically inside the span, Kahan / Parlett)
        double nprev = bv_piv;
        for (int round = 0; round < 5; ++round) {
            // lane i < k owns direction i: its coefficient, then everyone subtracts
            for (int i0 = 0; i0 < k; i0 += 32) {
                const int i = i0 + lane;
                double c = 0.0;
                if (i < k)
                    for (int r = 0; r < m; ++r) c = fma(S[oQ + i * m + r], S[oT + r], c);
                __syncwarp();
                for (int ii = 0; ii < 32 && i0 + ii < k; ++ii) {
                    const double ci = __shfl_sync(FULL_MASK, c, ii);
                    for (int r = lane; r < m; r += 32) S[oT + r] = fma(-ci, S[oQ + (i0 + ii) * m + r], S[oT + r]);
                }
                __syncwarp();
            }
            double n2 = 0.0;
            for (int r = lane; r < m; r += 32) n2 = fma(S[oT + r], S[oT + r], n2);
            n2 = warp_sum(n2);
            const bool settled = (n2 > 0.25 * nprev);
            nprev = n2;
            if (settled && round > 0) break;
        }
        const double nn = nprev;
        const double inv = (nn > 0.0) ? 1.0 / sqrt(nn) : 0.0;
        for (int r = lane; r < m; r += 32) S[oQ + k * m + r] = S[oT + r] * inv;
        __syncwarp();
        // ---- deflate: w_j -= (q . w_j) q for every column
        for (int j = lane; j < n; j += 32) {
            double c = 0.0;
            for (int r = 0; r < m; ++r) c = fma(S[oQ + k * m + r], S[oW + r * ldw + j], c);
            for (int r = 0; r < m; ++r) S[oW + r * ldw + j] = fma(-c, S[oQ + k * m + r], S[oW + r * ldw + j]);
        }
        __syncwarp();
    }
    // ---- C = U^T D from the dictionary itself, and the true residual of the reduction
    double worst = 0.0;
    for (int j = lane; j < n; j += 32) {
        for (int k = 0; k < R; ++k) {
            double c = 0.0;
            for (int r = 0; r < m; ++r) c = fma(S[oQ + k * m + r], D[r * n + j], c);
            coef[((size_t)a * n + j) * R + k] = c;
        }
        double res2 = 0.0;
        for (int r = 0; r < m; ++r) {
            double v = D[r * n + j];
            for (int k = 0; k < R; ++k) v = fma(-S[oQ + k * m + r], coef[((size_t)a * n + j) * R + k], v);
            res2 = fma(v, v, res2);
        }
        worst = fmax(worst, res2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(FULL_MASK, worst, o));
    for (int i = lane; i < m * R; i += 32) {
        const int r = i / R, k = i - r * R;
        basis[(size_t)a * m * R + i] = S[oQ + k * m + r];
    }
    if (lane == 0 && tail) tail[a] = (dmax2 > 0.0) ? sqrt(worst / dmax2) : 0.0;
}

}  // namespace met2

using namespace met2;

extern "C" int met2_echo_basis(const double* dic, int nA, int nTE, int nT2, int R, double* basis, double* coef,
                               double* tail, void* stream) {
    if (!dic || !basis || !coef || nA <= 0 || nT2 <= 0 || nT2 > MET2_MAX_NT2 || nTE <= 0 || nTE > MET2_MAX_NTE || R < 1 ||
        R > MET2_MAX_NTE)
        return set_error(MET2_ERR_ARG, "met2_echo_basis: bad argument (nA=%d nTE=%d nT2=%d R=%d)", nA, nTE, nT2, R);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(double) * ((size_t)nTE * (nT2 | 1) + (size_t)R * nTE + nTE);
    cudaError_t e = cudaFuncSetAttribute(echo_basis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(MET2_ERR_CUDA, "echo_basis attr: %s", cudaGetErrorString(e));
    MET2_LAUNCH(nA, 32, smem, st, echo_basis_kernel)(dic, nTE, nT2, R, basis, coef, tail);
    count_launch();
    return check_launch("echo_basis_kernel");
}
