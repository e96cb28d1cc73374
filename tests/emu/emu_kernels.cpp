// emu_kernels.cpp — TEST INFRASTRUCTURE ONLY: the T2 fit kernels of csrc/ compiled by g++ against the SIMT emulator
// (simt_emu.h) and exposed through a tiny C interface for tests/test_emu_kernels.py.  Host arrays in, host arrays out;
// one emulated thread block walks every tile.  Nothing here is reachable from the product library.
#define MET2_HOST_EMU 1
#include "simt_emu.h"

namespace simt {
Block* g_block = nullptr;
emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
long long n_fma = 0, n_smem = 0, n_syncwarp = 0;
}  // namespace simt

#include "met2_t2_echo.cu"
#include "met2_fa.cu"

namespace met2 {
constexpr size_t S_DOUBLES = 40960;   // 320 KB: more than any kernel's dynamic shared memory
alignas(16) static double s_storage[S_DOUBLES];
simt::Shared S = {s_storage, (long)S_DOUBLES};
int set_error(int code, const char*, ...) { return code; }
int check_launch(const char*) { return 0; }
void count_launch(int) {}
int sm_count() { return 1; }
}  // namespace met2

using namespace met2;

// Host stand-in for t2_hist / t2_tiles / t2_scatter: one tile per populated flip-angle index (split at T2_TILE_MAX).
static void make_tiles(const int* fa_index, long long V, int nA, std::vector<int>& perm, std::vector<int>& tile_fa,
                       std::vector<int>& tile_start, std::vector<int>& tile_cnt) {
    std::vector<std::vector<int>> bins(nA);
    for (long long v = 0; v < V; ++v) {
        int f = fa_index[v];
        if (f < 0 || f >= nA) f = 0;
        bins[f].push_back((int)v);
    }
    for (int a = 0; a < nA; ++a) {
        for (size_t s = 0; s < bins[a].size(); s += T2_TILE_MAX) {
            tile_fa.push_back(a);
            tile_start.push_back((int)perm.size() + 0);
            const size_t c = std::min(bins[a].size() - s, (size_t)T2_TILE_MAX);
            tile_cnt.push_back((int)c);
            for (size_t i = 0; i < c; ++i) perm.push_back(bins[a][s + i]);
        }
    }
}

struct Prepared {
    std::vector<int> perm, tile_fa, tile_start, tile_cnt;
    int counters[4];
};

static T2Args make_args(Prepared& P, const double* sig, const int* fa_index, long long V, const met2_t2_cfg* cfg,
                        const double* dic, const double* dicT, const double* G, const double* kband,
                        const double* lambdas, const double* logT2, const unsigned char* comp, double* fsol, double* est,
                        double* reg, double* maps, unsigned* status) {
    make_tiles(fa_index, V, cfg->nA, P.perm, P.tile_fa, P.tile_start, P.tile_cnt);
    P.counters[0] = (int)P.tile_fa.size();
    P.counters[1] = 0;
    T2Args A;
    memset(&A, 0, sizeof(A));
    A.sig = sig; A.fa_index = fa_index; A.V = V; A.cfg = *cfg;
    A.dic = dic; A.dicT = dicT; A.G = G; A.kband = kband; A.lambdas = lambdas; A.logT2 = logT2; A.comp = comp;
    A.fsol = fsol; A.est = est; A.reg = reg; A.maps = maps; A.status = status;
    A.perm = P.perm.data(); A.tile_fa = P.tile_fa.data(); A.tile_start = P.tile_start.data();
    A.tile_cnt = P.tile_cnt.data(); A.counters = P.counters;
    return A;
}

// The production kernel t2_fit_kernel<NS, ME, METHOD> without the shared full-set factor tables (A.tfull = NULL: those
// only shortcut factorisations) — same code path as the GPU otherwise.  Instantiated for the reference's sizes
// (nT2 <= 64, nTE <= 32: every method), T2SPARC's 96 bins (NS = 3) and BASELINE.json's config 4 (100 bins, 48 echoes:
// NS = 4, ME = 2; NNLS, X2, BayesReg).
template <int NS, int ME, int METHOD>
static long long run_t2(T2Args& A, int warps) {
    const size_t need = (size_t)(t2_table_doubles(A.cfg.nT2) + warps * t2_warp_doubles<NS>(A.pmax));
    if (need > S_DOUBLES) return -2;
    S.size = (long)need;
    return simt::run_block(warps * 32, 0, 1, [&]() { t2_fit_kernel<NS, ME, METHOD>(A); });
}

extern "C" {

// work counters since the last call: [fma, shared-memory accesses, __syncwarp] per thread (divide by 32 for warp level)
void emu_counters(long long* out) {
    out[0] = simt::n_fma;
    out[1] = simt::n_smem;
    out[2] = simt::n_syncwarp;
    simt::n_fma = simt::n_smem = simt::n_syncwarp = 0;
}

// The experimental echo-space X2 kernel (csrc/met2_t2_echo.cu).  Returns the number of warp collectives executed.
long long emu_t2_echo_x2(const double* sig, const int* fa_index, long long V, const met2_t2_cfg* cfg, const double* dic,
                         const double* dicT, const double* G, const double* kband, const double* logT2,
                         const unsigned char* comp, double* fsol, double* est, double* reg, double* maps,
                         unsigned* status, int warps) {
    if (!t2_echo_eligible(cfg)) return -1;
    Prepared P;
    T2Args A = make_args(P, sig, fa_index, V, cfg, dic, dicT, G, kband, nullptr, logT2, comp, fsol, est, reg, maps, status);
    const size_t need = (size_t)(echo_table_doubles(cfg->nT2) + warps * echo_warp_doubles());
    if (need > S_DOUBLES) return -2;
    S.size = (long)need;
    return simt::run_block(warps * 32, 0, 1, [&]() { t2_echo_x2_kernel(A); });
}

long long emu_t2_fit(const double* sig, const int* fa_index, long long V, const met2_t2_cfg* cfg, const double* dic,
                     const double* dicT, const double* G, const double* kband, const double* lambdas, const double* logT2,
                     const unsigned char* comp, double* fsol, double* est, double* reg, double* maps, unsigned* status,
                     int warps) {
    Prepared P;
    T2Args A = make_args(P, sig, fa_index, V, cfg, dic, dicT, G, kband, lambdas, logT2, comp, fsol, est, reg, maps, status);
    const bool plain = (cfg->method == MET2_REG_NNLS);
    A.pmax = plain ? std::min(cfg->nT2, cfg->nTE) : cfg->nT2;
    if (cfg->method == MET2_REG_GCV)
        while (tri(A.pmax) < gcv_region_doubles(cfg->nT2)) ++A.pmax;
    A.warps = warps;
    const int ns = (cfg->nT2 + 31) / 32, me = (cfg->nTE + 31) / 32;
    if (ns <= 2 && me == 1) {
        switch (cfg->method) {
            case MET2_REG_NNLS: return run_t2<2, 1, MET2_REG_NNLS>(A, warps);
            case MET2_REG_T2SPARC: return run_t2<2, 1, MET2_REG_T2SPARC>(A, warps);
            case MET2_REG_X2: return run_t2<2, 1, MET2_REG_X2>(A, warps);
            case MET2_REG_LCURVE: return run_t2<2, 1, MET2_REG_LCURVE>(A, warps);
            case MET2_REG_GCV: return run_t2<2, 1, MET2_REG_GCV>(A, warps);
            case MET2_REG_BAYESREG: return run_t2<2, 1, MET2_REG_BAYESREG>(A, warps);
            default: return -3;
        }
    }
    if (ns == 3 && me == 1 && cfg->method == MET2_REG_T2SPARC) return run_t2<3, 1, MET2_REG_T2SPARC>(A, warps);
    if (ns == 4 && me == 2) {
        switch (cfg->method) {
            case MET2_REG_NNLS: return run_t2<4, 2, MET2_REG_NNLS>(A, warps);
            case MET2_REG_X2: return run_t2<4, 2, MET2_REG_X2>(A, warps);
            case MET2_REG_BAYESREG: return run_t2<4, 2, MET2_REG_BAYESREG>(A, warps);
            default: return -3;
        }
    }
    return -1;
}

// The flip-angle stage (csrc/met2_fa.cu): fa_search_kernel -> [spline_weights_kernel] -> fa_select_kernel ->
// reduce_partials_kernel, nT2 <= 64 and nTE <= 32, one emulated block each (the search kernel walks all voxels with
// `warps` warps; the select kernel has its fixed FA_WARPS warps).  dic_s/dicT_s/G_s/knots: the coarse (spline) tables
// or NULL for brute force.
long long emu_fa_fit(const double* sig, long long V, const met2_fa_cfg* cfg, const double* dic, const double* dicT,
                     const double* G, const double* alphas, const double* dic_s, const double* dicT_s, const double* G_s,
                     const double* knots, int* fa_index, double* fa_deg, double* km, double* fsol_sum, unsigned* status,
                     int warps) {
    if (cfg->nT2 > 64 || cfg->nTE > 32) return -1;
    const bool spline = cfg->method == MET2_FA_SPLINE;
    FaArgs A;
    memset(&A, 0, sizeof(A));
    A.sig = sig; A.V = V; A.cfg = *cfg;
    A.dic = dic; A.dicT = dicT; A.G = G; A.alphas = alphas;
    if (spline) {
        A.dic_s = dic_s; A.dicT_s = dicT_s; A.G_s = G_s; A.knots = knots; A.nS = cfg->nKnots;
    } else {
        A.dic_s = dic; A.dicT_s = dicT; A.G_s = G; A.knots = nullptr; A.nS = cfg->nA;
    }
    A.fa_index = fa_index; A.fa_deg = fa_deg; A.km = km; A.fsol_sum = fsol_sum; A.status = status;
    const int nSr = spline ? cfg->nKnots : 1;
    std::vector<double> resid((size_t)V * nSr), wsp(MET2_MAX_KNOTS * MET2_MAX_KNOTS), partial((size_t)FA_WARPS * cfg->nT2);
    std::vector<int> ws_p(V), ws_ix((size_t)V * FA_CARRY);
    std::vector<double> ws_x((size_t)V * FA_CARRY);
    A.resid = resid.data(); A.wsp = wsp.data(); A.partial = fsol_sum ? partial.data() : nullptr;
    A.ws_p = ws_p.data(); A.ws_ix = ws_ix.data(); A.ws_x = ws_x.data();
    A.pmax = std::min(cfg->nT2, cfg->nTE);
    for (long long v = 0; v < V; ++v) status[v] = 0;
    long long c = 0;
    S.size = fa_table_doubles(cfg->nT2, cfg->nTE) + warps * fa_warp_doubles<2>(A.pmax);
    if ((size_t)S.size > S_DOUBLES) return -2;
    c += simt::run_block(warps * 32, 0, 1, [&]() { fa_search_kernel<2, 1>(A); });
    if (spline) c += simt::run_block(32, 0, 1, [&]() { spline_weights_kernel(A.knots, A.cfg.nKnots, A.wsp); });
    S.size = FA_WARPS * fa_warp_doubles<2>(A.pmax);
    c += simt::run_block(FA_WARPS * 32, 0, 1, [&]() { fa_select_kernel<2, 1>(A); });
    if (fsol_sum) {
        const int nb = (cfg->nT2 + 127) / 128;
        for (int b = 0; b < nb; ++b)
            c += simt::run_block(128, b, nb, [&]() { reduce_partials_kernel(A.partial, (long long)FA_WARPS, A.cfg.nT2, A.fsol_sum); });
    }
    return c;
}

}  // extern "C"
