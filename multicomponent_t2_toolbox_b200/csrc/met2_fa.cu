// met2_fa.cu — flip-angle estimation for a batch of voxels (Step 2 of the reference orchestrator).
//
// Takes over flip_angle_algorithms/fa_estimation.py: compute_optimal_FA (:74-90, brute force),
// fitting_slice_FA_spline_method (:35-70) and the joblib loop that drives them (motor/motor_recon_met2_real_data.py:349-373).
//   kernel 1 (fa_search_kernel): residual norm of the plain NNLS fit for every search angle and voxel.  Angles are the
//       OUTER loop of each CTA: the ~59 KB of tables of one angle (D, D^T, G) are staged in shared memory once per
//       angle and CTA (the version that read them through L1/L2 spent its time on L2 latency), and every voxel's fit
//       is warm-started from its fit at the previous angle.
//   kernel 2 (fa_select_kernel): brute force -> the running arg-min of kernel 1; spline -> not-a-knot cubic through the
//       15 knot residuals, SciPy's bounded Brent on it, snap to the fine grid; then the NNLS at the chosen angle for
//       km = sum(f) and the per-warp partial sums of f (mean_T2_dist).
//   kernel 3: fixed-order reduction of those partial sums (deterministic for a given launch geometry).
#include <cmath>

#include "met2_device.cuh"
#include "met2_host.h"

namespace met2 {

struct FaArgs {
    const double* sig;
    long long V;
    met2_fa_cfg cfg;
    const double *dic, *dicT, *G, *alphas;        // fine grid tables
    const double *dic_s, *dicT_s, *G_s, *knots;   // search tables (== fine for brute force)
    int nS;                                        // number of search angles
    int* fa_index;
    double *fa_deg, *km, *fsol_sum;
    unsigned* status;
    // workspace
    double* resid;      // spline: [V][nS]; brute force: best residual [V]
    double* wsp;        // [nK][nK] spline weights
    double* partial;    // [total warps][nT2]
    int* ws_p;          // [V] size of the positive set carried from the previous search angle (0 = none)
    int* ws_ix;         // [V][FA_CARRY] its columns
    double* ws_x;       // [V][FA_CARRY] its coefficients
    int pmax;
};

constexpr int FA_CARRY = 16;   // supports of the plain fits have 3-8 columns; larger ones restart cold

constexpr int FA_WARPS = 8;

// per-warp shared memory in doubles: NNLS slots + signal ms[64]
template <int NS>
__host__ __device__ __forceinline__ int fa_warp_doubles(int pmax) {
    return (Slots<NS>::doubles(pmax) + 64 + 31) & ~31;
}

constexpr int FA_SEARCH_WARPS = 16;

__host__ __device__ __forceinline__ int fa_ldg(int n) { return (n + 1) & ~1; }
// staged tables of one angle, in doubles: D [m][n], Dt [n][m], G [n][ldg]
__host__ __device__ __forceinline__ int fa_table_doubles(int n, int m) { return (2 * m * n + n * fa_ldg(n) + 31) & ~31; }

template <int NS, int ME>
__global__ void __launch_bounds__(FA_SEARCH_WARPS * 32, 1) fa_search_kernel(FaArgs A) {
    // a plain solve has at most min(nT2, nTE) <= 32 ME positive columns: position-space loops of ME slots, not NS
    constexpr int FA_PS = (ME < NS) ? ME : NS;
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = A.cfg.nT2, m = A.cfg.nTE;
    const int ldg = fa_ldg(n);
    const int oD = 0, oDt = m * n, oG = 2 * m * n;
    Slots<NS> W;
    const int wbase = fa_table_doubles(n, m) + warp * fa_warp_doubles<NS>(A.pmax);
    W.carve(wbase, A.pmax);
    const int oM = wbase + Slots<NS>::doubles(A.pmax);
    const long long chunk = (A.V + gridDim.x - 1) / gridDim.x;
    const long long v0 = (long long)blockIdx.x * chunk;
    const long long v1 = (v0 + chunk < A.V) ? v0 + chunk : A.V;
    const bool brute = (A.cfg.method == MET2_FA_BRUTE_FORCE);
    for (int a = 0; a < A.nS; ++a) {
        {
            const double* D = A.dic_s + (size_t)a * m * n;
            const double* Dt = A.dicT_s + (size_t)a * n * m;
            const double* G = A.G_s + (size_t)a * n * n;
            for (int i = threadIdx.x; i < m * n; i += blockDim.x) {
                S[oD + i] = __ldg(D + i);
                S[oDt + i] = __ldg(Dt + i);
            }
            for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
                const int r = i / n;
                S[oG + r * ldg + (i - r * n)] = __ldg(G + i);
            }
        }
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        // the warps pull the CTA's voxels one at a time (a static deal left them waiting for the slowest warp at the
        // barrier that ends every angle: 1.3 stalled warps per issue in profiles/r02_fa_*_ncu_summary.txt)
        while (true) {
            int iv = 0;
            if (lane == 0) iv = atomicAdd(&s_next, 1);
            iv = __shfl_sync(FULL_MASK, iv, 0);
            const long long v = v0 + iv;
            if (v >= v1) break;
            unsigned st = load_signal<ME>(A.sig, v, m, oM, lane);
            if (st) continue;
            compute_c_sh<NS>(W, oD, oM, m, n, lane);
            // warm start from the fit at the previous search angle (same voxel, neighbouring dictionary): the residual
            // norm of the NNLS optimum does not depend on the starting point
            int p0 = 0;
            if (a > 0) {
                p0 = A.ws_p[v];
                if (lane < p0) {
                    SI(W.ix, lane) = A.ws_ix[v * FA_CARRY + lane];
                    S[W.xs + lane] = A.ws_x[v * FA_CARRY + lane];
                }
                __syncwarp();
            }
            int nst = 0;
            set_dspace<NS>(W, A.dicT_s + (size_t)a * n * m, oM, lane);
            int p = nnls_gram<NS, true, FA_PS>(W, oG, nullptr, ldg, 0, false, 0.0, n, m, lane, nst, p0);
            if (a + 1 < A.nS) {
                const bool keep = (p <= FA_CARRY);
                if (keep && lane < p) {
                    A.ws_ix[v * FA_CARRY + lane] = SI(W.ix, lane);
                    A.ws_x[v * FA_CARRY + lane] = S[W.xs + lane];
                }
                if (lane == 0) A.ws_p[v] = keep ? p : 0;
            }
            double fit[ME];
            double sse = fit_and_sse_sh<NS, ME>(W, oDt, oM, m, p, lane, fit);
            double rnorm = sqrt(sse);
            if (lane == 0) {
                if (brute) {
                    // np.argmin: first minimum (fa_estimation.py:83)
                    if (a == 0 || rnorm < A.resid[v]) {
                        A.resid[v] = rnorm;
                        A.fa_index[v] = a;
                    }
                } else {
                    A.resid[v * A.nS + a] = rnorm;
                }
                if (nst) atomicOr(&A.status[v], MET2_ST_ITMAX);
            }
            __syncwarp();
        }
        __syncthreads();   // all warps are done with this angle's tables
    }
}

// Not-a-knot cubic spline as a linear map from knot values to knot second derivatives: Msec = Wsp * y.
// (scipy.interpolate.interp1d(kind='cubic') -> make_interp_spline(k=3), default not-a-knot; fa_estimation.py:54.)
// Gauss-Jordan with partial pivoting on the augmented K x 2K system, one warp: lane j owns columns j and j + 32 of every
// row, so each entry sees exactly the operations (and the order) of a serial elimination.  (Round 1 ran it on ONE
// thread with the matrix in a 16.6 KB local-memory frame: 112 us in front of every FA call.)
__global__ void __launch_bounds__(32, 1) spline_weights_kernel(const double* __restrict__ knots, int K,
                                                               double* __restrict__ wsp) {
    constexpr int LD = 2 * MET2_MAX_KNOTS + 1;
    __shared__ double Aug[MET2_MAX_KNOTS * LD];
    __shared__ double h[MET2_MAX_KNOTS];
    const int lane = threadIdx.x;
    const int W2 = 2 * K;
    for (int i = 0; i < K; ++i)
        for (int j = lane; j < W2; j += 32) Aug[i * LD + j] = 0.0;
    if (lane + 1 < K) h[lane] = knots[lane + 1] - knots[lane];
    __syncwarp();
    if (lane == 0) {
        Aug[0] = -1.0 / h[0];
        Aug[1] = 1.0 / h[0] + 1.0 / h[1];
        Aug[2] = -1.0 / h[1];
        Aug[(K - 1) * LD + K - 3] = -1.0 / h[K - 3];
        Aug[(K - 1) * LD + K - 2] = 1.0 / h[K - 3] + 1.0 / h[K - 2];
        Aug[(K - 1) * LD + K - 1] = -1.0 / h[K - 2];
    }
    if (lane >= 1 && lane + 1 < K) {
        const int i = lane;
        Aug[i * LD + i - 1] = h[i - 1] / 6.0;
        Aug[i * LD + i] = (h[i - 1] + h[i]) / 3.0;
        Aug[i * LD + i + 1] = h[i] / 6.0;
        Aug[i * LD + K + i - 1] = 1.0 / h[i - 1];
        Aug[i * LD + K + i] = -1.0 / h[i - 1] - 1.0 / h[i];
        Aug[i * LD + K + i + 1] = 1.0 / h[i];
    }
    __syncwarp();
    for (int c = 0; c < K; ++c) {
        int piv = c;
        double best = fabs(Aug[c * LD + c]);
        for (int r = c + 1; r < K; ++r)
            if (fabs(Aug[r * LD + c]) > best) {
                best = fabs(Aug[r * LD + c]);
                piv = r;
            }
        __syncwarp();
        if (piv != c) {
            for (int j = lane; j < W2; j += 32) {
                double t = Aug[c * LD + j];
                Aug[c * LD + j] = Aug[piv * LD + j];
                Aug[piv * LD + j] = t;
            }
            __syncwarp();
        }
        const double inv = 1.0 / Aug[c * LD + c];
        __syncwarp();
        for (int j = lane; j < W2; j += 32) Aug[c * LD + j] *= inv;
        __syncwarp();
        const double f_own = (lane < K) ? Aug[lane * LD + c] : 0.0;   // column c of every row, before it is eliminated
        __syncwarp();
        for (int r = 0; r < K; ++r) {
            const double f = __shfl_sync(0xffffffffu, f_own, r);
            if (r == c || f == 0.0) continue;
            for (int j = lane; j < W2; j += 32) Aug[r * LD + j] -= f * Aug[c * LD + j];
        }
        __syncwarp();
    }
    for (int i = 0; i < K; ++i)
        for (int j = lane; j < K; j += 32) wsp[i * K + j] = Aug[i * LD + K + j];
}

// knots at S[oX..], values at S[oY..], second derivatives at S[oM2..]
__device__ __forceinline__ double spline_eval(int oX, int oY, int oM2, int K, double x) {
    int i = 0;
    while (i + 2 < K && x >= S[oX + i + 1]) ++i;
    double h = S[oX + i + 1] - S[oX + i];
    double a = (S[oX + i + 1] - x) / h;
    double b = (x - S[oX + i]) / h;
    return a * S[oY + i] + b * S[oY + i + 1] +
           ((a * a * a - a) * S[oM2 + i] + (b * b * b - b) * S[oM2 + i + 1]) * (h * h) / 6.0;
}

template <int NS, int ME>
__global__ void __launch_bounds__(FA_WARPS * 32, 3) fa_select_kernel(FaArgs A) {
    constexpr int FA_PS = (ME < NS) ? ME : NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = A.cfg.nT2, m = A.cfg.nTE, nA = A.cfg.nA;
    Slots<NS> W;
    const int wbase = warp * fa_warp_doubles<NS>(A.pmax);
    W.carve(wbase, A.pmax);
    const int oM = wbase + Slots<NS>::doubles(A.pmax);
    const bool brute = (A.cfg.method == MET2_FA_BRUTE_FORCE);
    const int K = A.cfg.nKnots;
    double fs[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) fs[s] = 0.0;
    const long long gw = (long long)blockIdx.x * FA_WARPS + warp;
    const long long nw = (long long)gridDim.x * FA_WARPS;
    for (long long v = gw; v < A.V; v += nw) {
        unsigned st = load_signal<ME>(A.sig, v, m, oM, lane);
        if (st) {
            if (lane == 0) {
                A.fa_index[v] = 0;
                A.fa_deg[v] = 0.0;
                A.km[v] = 0.0;
                A.status[v] |= st;
            }
            continue;
        }
        int index;
        if (brute) {
            index = A.fa_index[v];
        } else {
            // knot values and second derivatives into shared scratch: gs = y, rs = Msec, xs = knots
            if (lane < K) {
                S[W.gs + lane] = A.resid[v * K + lane];
                S[W.xs + lane] = A.knots[lane];
            }
            __syncwarp();
            if (lane < K) {
                double acc = 0.0;
                for (int j = 0; j < K; ++j) acc += A.wsp[lane * K + j] * S[W.gs + j];
                S[W.rs + lane] = acc;
            }
            __syncwarp();
            Brent B;
            double xq = B.start(A.cfg.brent_lo, A.cfg.brent_hi, A.cfg.brent_xatol, A.cfg.brent_maxfun);
            while (B.feed(spline_eval(W.xs, W.gs, W.rs, K, xq), xq)) {
            }
            const double xmin = B.xf;
            // indexFA = argmin |alpha_values - res.x| (first minimum; fa_estimation.py:57)
            double bd = 1.0e300;
            int bi = -1;
            for (int i = lane; i < nA; i += 32) {
                double d = fabs(A.alphas[i] - xmin);
                if (d < bd) {
                    bd = d;
                    bi = i;
                }
            }
            double dmin;
            index = warp_argmin_nonneg(bd, bi, dmin);
            if (index < 0) index = 0;
            __syncwarp();
        }
        double kmv = 0.0;
        if (A.cfg.final_solve) {
            const double* D = A.dic + (size_t)index * m * n;
            const double* G = A.G + (size_t)index * n * n;
            compute_c<NS>(W, D, oM, m, n, lane);
            int nst = 0;
            set_dspace<NS>(W, A.dicT + (size_t)index * n * m, oM, lane);
            (void)nnls_gram<NS, false, FA_PS>(W, 0, G, n, 0, false, 0.0, n, m, lane, nst);
            if (nst && lane == 0) A.status[v] |= MET2_ST_ITMAX;
            // nnls_gram leaves the solution in column space in S[W.xc..]; km = sum(f)
            double part = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                int col = NS * lane + s;
                double xv = (col < n) ? S[W.xc + col] : 0.0;
                fs[s] += xv;
                part += xv;
            }
            kmv = warp_sum(part);
            __syncwarp();
        }
        if (lane == 0) {
            A.fa_index[v] = index;
            A.fa_deg[v] = A.alphas[index];
            A.km[v] = kmv;
        }
    }
    if (A.partial) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int col = NS * lane + s;
            if (col < n) A.partial[gw * n + col] = fs[s];
        }
    }
}

// One CTA per T2 bin: thread t adds rows t, t + 256, ... and a shared-memory tree joins the 256 partial sums — a fixed
// order for a given launch geometry.  (Round 1: one thread per bin walked all ~9 500 rows, 178 us for 60 sums.)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partial, long long nrows, int n,
                                                              double* __restrict__ out) {
    __shared__ double part[256];
    const int c = blockIdx.x;
    double acc = 0.0;
    for (long long r = threadIdx.x; r < nrows; r += 256) acc += partial[r * n + c];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = part[0];
}

struct FaGeom {
    int grid;          // select kernel
    size_t smem;
    int pmax;
    int grid_search;   // search kernel: one CTA of up to FA_SEARCH_WARPS warps per SM, tables staged in shared memory
    size_t smem_search;
    int warps_search;
};

template <int NS>
static FaGeom fa_geometry(const met2_fa_cfg* cfg) {
    FaGeom g;
    g.pmax = cfg->nT2 < cfg->nTE ? cfg->nT2 : cfg->nTE;
    g.smem = sizeof(double) * (size_t)fa_warp_doubles<NS>(g.pmax) * FA_WARPS;
    int per_sm = (int)((size_t)(200 * 1024) / (g.smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    g.grid = sms * per_sm;
    g.grid_search = sms;
    {
        const size_t tables = sizeof(double) * (size_t)fa_table_doubles(cfg->nT2, cfg->nTE);
        const size_t per_warp = sizeof(double) * (size_t)fa_warp_doubles<NS>(g.pmax);
        const size_t budget = 227 * 1024 - 1024;
        int w = tables < budget ? (int)((budget - tables) / per_warp) : 0;
        if (w > FA_SEARCH_WARPS) w = FA_SEARCH_WARPS;
        if (w < 1) w = 1;
        g.warps_search = w;
        g.smem_search = tables + per_warp * w;
    }
    return g;
}

static FaGeom fa_geometry_any(const met2_fa_cfg* cfg) {
    int ns = (cfg->nT2 + 31) / 32;
    if (ns <= 2) return fa_geometry<2>(cfg);
    if (ns == 3) return fa_geometry<3>(cfg);
    return fa_geometry<4>(cfg);
}

template <int NS, int ME>
static int fa_launch(const FaArgs& A, const FaGeom& g, cudaStream_t st) {
    cudaError_t e;
    e = cudaFuncSetAttribute(fa_search_kernel<NS, ME>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)g.smem_search);
    if (e != cudaSuccess) return set_error(MET2_ERR_CUDA, "fa_search attr: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(fa_select_kernel<NS, ME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return set_error(MET2_ERR_CUDA, "fa_select attr: %s", cudaGetErrorString(e));
    MET2_LAUNCH(g.grid_search, g.warps_search * 32, g.smem_search, st, fa_search_kernel<NS, ME>)(A);
    count_launch();
    int rc = check_launch("fa_search_kernel");
    if (rc) return rc;
    if (A.cfg.method == MET2_FA_SPLINE) {
        MET2_LAUNCH(1, 32, 0, st, spline_weights_kernel)(A.knots, A.cfg.nKnots, A.wsp);
        count_launch();
        rc = check_launch("spline_weights_kernel");
        if (rc) return rc;
    }
    MET2_LAUNCH(g.grid, FA_WARPS * 32, g.smem, st, fa_select_kernel<NS, ME>)(A);
    count_launch();
    rc = check_launch("fa_select_kernel");
    if (rc) return rc;
    if (A.fsol_sum) {
        MET2_LAUNCH(A.cfg.nT2, 256, 0, st, reduce_partials_kernel)(A.partial, (long long)g.grid * FA_WARPS, A.cfg.nT2,
                                                                   A.fsol_sum);
        count_launch();
        rc = check_launch("reduce_partials_kernel");
    }
    return rc;
}

}  // namespace met2

using namespace met2;

static int fa_check_cfg(const met2_fa_cfg* cfg) {
    if (!cfg) return set_error(MET2_ERR_ARG, "met2_fa: cfg is NULL");
    if (cfg->nT2 <= 0 || cfg->nT2 > MET2_MAX_NT2 || cfg->nTE <= 0 || cfg->nTE > MET2_MAX_NTE || cfg->nA <= 0)
        return set_error(MET2_ERR_ARG, "met2_fa: unsupported sizes nT2=%d nTE=%d nA=%d", cfg->nT2, cfg->nTE, cfg->nA);
    if (cfg->method != MET2_FA_BRUTE_FORCE && cfg->method != MET2_FA_SPLINE)
        return set_error(MET2_ERR_ARG, "met2_fa: unknown method %d", cfg->method);
    if (cfg->method == MET2_FA_SPLINE && (cfg->nKnots < 4 || cfg->nKnots > MET2_MAX_KNOTS))
        return set_error(MET2_ERR_ARG, "met2_fa: spline needs 4..%d knots, got %d", MET2_MAX_KNOTS, cfg->nKnots);
    return MET2_OK;
}

extern "C" int64_t met2_fa_workspace_bytes(int64_t V, const met2_fa_cfg* cfg) {
    if (fa_check_cfg(cfg) || V < 0) return -1;
    FaGeom g = fa_geometry_any(cfg);
    int nS = cfg->method == MET2_FA_SPLINE ? cfg->nKnots : 1;
    size_t b = align256(sizeof(double) * (size_t)V * nS);
    b += align256(sizeof(double) * MET2_MAX_KNOTS * MET2_MAX_KNOTS);
    b += align256(sizeof(double) * (size_t)g.grid * FA_WARPS * cfg->nT2);
    b += align256(sizeof(int) * (size_t)V) + align256(sizeof(int) * (size_t)V * FA_CARRY) +
         align256(sizeof(double) * (size_t)V * FA_CARRY);
    return (int64_t)b + 256;
}

extern "C" int met2_fa_fit(const double* sig, int64_t V, const met2_fa_cfg* cfg, const double* dic, const double* dicT,
                           const double* G, const double* alphas, const double* dic_s, const double* dicT_s,
                           const double* G_s, const double* knots, int32_t* fa_index, double* fa_deg, double* km,
                           double* fsol_sum, uint32_t* status, void* workspace, void* stream) {
    int rc = fa_check_cfg(cfg);
    if (rc) return rc;
    if (V < 0 || !sig || !dic || !dicT || !G || !alphas || !fa_index || !fa_deg || !km || !status || !workspace)
        return set_error(MET2_ERR_ARG, "met2_fa_fit: NULL argument");
    const bool spline = cfg->method == MET2_FA_SPLINE;
    if (spline && (!dic_s || !dicT_s || !G_s || !knots))
        return set_error(MET2_ERR_ARG, "met2_fa_fit: spline method needs the coarse dictionary and knots");
    if (V == 0) return MET2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FaGeom g = fa_geometry_any(cfg);
    FaArgs A;
    A.sig = sig;
    A.V = V;
    A.cfg = *cfg;
    A.dic = dic; A.dicT = dicT; A.G = G; A.alphas = alphas;
    if (spline) {
        A.dic_s = dic_s; A.dicT_s = dicT_s; A.G_s = G_s; A.knots = knots;
        A.nS = cfg->nKnots;
    } else {
        A.dic_s = dic; A.dicT_s = dicT; A.G_s = G; A.knots = nullptr;
        A.nS = cfg->nA;
    }
    A.fa_index = fa_index; A.fa_deg = fa_deg; A.km = km; A.fsol_sum = fsol_sum; A.status = status;
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int nSr = spline ? cfg->nKnots : 1;
    A.resid = reinterpret_cast<double*>(w);
    w += align256(sizeof(double) * (size_t)V * nSr);
    A.wsp = reinterpret_cast<double*>(w);
    w += align256(sizeof(double) * MET2_MAX_KNOTS * MET2_MAX_KNOTS);
    A.partial = fsol_sum ? reinterpret_cast<double*>(w) : nullptr;
    w += align256(sizeof(double) * (size_t)g.grid * FA_WARPS * cfg->nT2);
    A.ws_p = reinterpret_cast<int*>(w);
    w += align256(sizeof(int) * (size_t)V);
    A.ws_ix = reinterpret_cast<int*>(w);
    w += align256(sizeof(int) * (size_t)V * FA_CARRY);
    A.ws_x = reinterpret_cast<double*>(w);
    A.pmax = g.pmax;
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(uint32_t) * (size_t)V, st);
    if (e != cudaSuccess) return set_error(MET2_ERR_CUDA, "memset status: %s", cudaGetErrorString(e));
    int ns = (cfg->nT2 + 31) / 32;
    int me = (cfg->nTE + 31) / 32;
    if (ns <= 2 && me == 1) return fa_launch<2, 1>(A, g, st);
    if (ns <= 2 && me == 2) return fa_launch<2, 2>(A, g, st);
    if (ns == 3 && me == 1) return fa_launch<3, 1>(A, g, st);
    if (ns == 3 && me == 2) return fa_launch<3, 2>(A, g, st);
    if (ns == 4 && me == 1) return fa_launch<4, 1>(A, g, st);
    if (ns == 4 && me == 2) return fa_launch<4, 2>(A, g, st);
    return set_error(MET2_ERR_UNSUPPORTED, "met2_fa_fit: unsupported template sizes");
}
