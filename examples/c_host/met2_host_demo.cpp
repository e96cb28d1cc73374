// met2_host_demo.cpp — a host program that drives the MET2 path through the C ABI alone (include/met2.h + the CUDA
// runtime; no Python, no torch): EPG dictionary -> Gram tables -> brute-force flip-angle search -> NNLS / X2 spectrum
// fit + maps, on signals synthesised from the dictionary itself, with self-checks.  Exit code 0 = all checks passed.
//
//   nvcc -std=c++17 -I include examples/c_host/met2_host_demo.cpp -L multicomponent_t2_toolbox_b200 -lmet2 \
//        -Xlinker -rpath -Xlinker $PWD/multicomponent_t2_toolbox_b200 -o examples/c_host/met2_host_demo
//
// This is the binding a C/C++ caller of the reference's Steps 2-4 (motor/motor_recon_met2_real_data.py:349-472) would
// write; the Python package does exactly the same calls through ctypes (multicomponent_t2_toolbox_b200/_lib.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "met2.h"

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)
#define MK(x)                                                                                  \
    do {                                                                                       \
        int rc_ = (x);                                                                         \
        if (rc_ != MET2_OK) {                                                                  \
            std::fprintf(stderr, "met2 error %d: %s at %s:%d\n", rc_, met2_last_error(), __FILE__, __LINE__); \
            return 3;                                                                          \
        }                                                                                      \
    } while (0)

template <class T>
static T* dev_alloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T) + 256) != cudaSuccess) return nullptr;
    return static_cast<T*>(p);
}

int main() {
    const int nTE = 32, nT2 = 60, nA = 91, V = 257;
    const double tau = 10.0, TR = 1000.0;
    std::printf("met2 C-ABI version %d\n", met2_version());
    // ---- grids (motor...:204-245): 60 log-spaced T2 in [10, 2000] ms, T1 = 1000 ms, 91 angles 90..180
    std::vector<double> T2s(nT2), T1s(nT2, 1000.0), alphas(nA), logT2(nT2);
    std::vector<unsigned char> comp(nT2);
    for (int i = 0; i < nT2; ++i) {
        T2s[i] = std::pow(10.0, 1.0 + (std::log10(2000.0) - 1.0) * i / (nT2 - 1));
        logT2[i] = std::log(T2s[i]);
        comp[i] = (unsigned char)((T2s[i] <= 40.0 ? 1 : 0) | ((T2s[i] > 40.0 && T2s[i] <= 200.0) ? 2 : 0) |
                                  (T2s[i] >= 200.0 ? 4 : 0));
    }
    for (int a = 0; a < nA; ++a) alphas[a] = 90.0 + a;
    double *d_T2 = dev_alloc<double>(nT2), *d_T1 = dev_alloc<double>(nT2), *d_al = dev_alloc<double>(nA);
    double *d_logT2 = dev_alloc<double>(nT2);
    unsigned char* d_comp = dev_alloc<unsigned char>(nT2);
    double *d_dic = dev_alloc<double>((size_t)nA * nTE * nT2), *d_dicT = dev_alloc<double>((size_t)nA * nTE * nT2);
    double* d_G = dev_alloc<double>((size_t)nA * nT2 * nT2);
    CK(cudaMemcpy(d_T2, T2s.data(), nT2 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_T1, T1s.data(), nT2 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_al, alphas.data(), nA * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_logT2, logT2.data(), nT2 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_comp, comp.data(), nT2, cudaMemcpyHostToDevice));
    // ---- dictionary + Gram tables (epg/epg.py:155)
    MK(met2_epg_dictionary(d_al, nA, d_T2, d_T1, nT2, nTE, tau, TR, d_dic, d_dicT, nullptr));
    MK(met2_gram_tables(d_dic, nA, nTE, nT2, nullptr, d_G, nullptr, nullptr, nullptr));
    // identity regularisation matrix -> band form of K = L^T L
    std::vector<double> L((size_t)nT2 * nT2, 0.0);
    for (int i = 0; i < nT2; ++i) L[(size_t)i * nT2 + i] = 1.0;
    double *d_L = dev_alloc<double>((size_t)nT2 * nT2), *d_kband = dev_alloc<double>(10 * nT2);
    int* d_band_err = dev_alloc<int>(1);
    CK(cudaMemcpy(d_L, L.data(), L.size() * 8, cudaMemcpyHostToDevice));
    MK(met2_gram_tables(nullptr, 0, nTE, nT2, d_L, nullptr, d_kband, d_band_err, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<double> dic((size_t)nA * nTE * nT2);
    CK(cudaMemcpy(dic.data(), d_dic, dic.size() * 8, cudaMemcpyDeviceToHost));
    // 180-degree column is a mono-exponential: (1 - exp(-TR/T1)) exp(-TE/T2)
    int fails = 0;
    for (int e = 0; e < nTE; ++e) {
        const double want = (1.0 - std::exp(-TR / 1000.0)) * std::exp(-(e + 1) * tau / T2s[0]);
        const double got = dic[((size_t)(nA - 1) * nTE + e) * nT2 + 0];
        if (std::fabs(got - want) > 1e-12 * want + 1e-300) ++fails;
    }
    std::printf("dictionary: 180-degree column check, %d mismatches\n", fails);
    // ---- signals: voxel v = 1000 * (0.15 D[:, j1, a] + 0.8 D[:, j2, a] + 0.05 D[:, 59, a]) at angle a(v); voxel 0 empty
    std::vector<double> sig((size_t)V * nTE, 0.0);
    std::vector<int> a_true(V), j1(V), j2(V);
    for (int v = 1; v < V; ++v) {
        a_true[v] = 10 + (v * 7) % 80;
        j1[v] = 4 + v % 8;
        j2[v] = 24 + v % 6;
        for (int e = 0; e < nTE; ++e) {
            const double* row = &dic[((size_t)a_true[v] * nTE + e) * nT2];
            // + a deterministic 0.05 % ripple so that the plain-NNLS residual is not exactly zero (X2 divides by it)
            sig[(size_t)v * nTE + e] = 1000.0 * (0.15 * row[j1[v]] + 0.8 * row[j2[v]] + 0.05 * row[nT2 - 1]) *
                                       (1.0 + 5e-4 * std::sin(1.7 * v + 2.3 * e));
        }
    }
    double* d_sig = dev_alloc<double>((size_t)V * nTE);
    CK(cudaMemcpy(d_sig, sig.data(), sig.size() * 8, cudaMemcpyHostToDevice));
    // ---- Step 2: brute-force FA search (fa_estimation.py:74-112)
    met2_fa_cfg fcfg = {};
    fcfg.method = MET2_FA_BRUTE_FORCE; fcfg.nTE = nTE; fcfg.nT2 = nT2; fcfg.nA = nA; fcfg.nKnots = 0; fcfg.final_solve = 1;
    fcfg.brent_lo = 90.0; fcfg.brent_hi = 180.0; fcfg.brent_xatol = 1e-5; fcfg.brent_maxfun = 500;
    const int64_t fa_ws = met2_fa_workspace_bytes(V, &fcfg);
    void* d_ws_fa = dev_alloc<unsigned char>((size_t)fa_ws);
    int32_t* d_idx = dev_alloc<int32_t>(V);
    double *d_fa = dev_alloc<double>(V), *d_km = dev_alloc<double>(V), *d_fsum = dev_alloc<double>(nT2);
    uint32_t* d_st = dev_alloc<uint32_t>(V);
    CK(cudaMemset(d_fsum, 0, nT2 * 8));
    MK(met2_fa_fit(d_sig, V, &fcfg, d_dic, d_dicT, d_G, d_al, d_dic, d_dicT, d_G, nullptr, d_idx, d_fa, d_km, d_fsum, d_st,
                   d_ws_fa, nullptr));
    // ---- Steps 3 + 4: X2 (factor 1.02) spectrum fit and maps (motor...:113-162, 443-472)
    met2_t2_cfg tcfg = {};
    tcfg.method = MET2_REG_X2; tcfg.nTE = nTE; tcfg.nT2 = nT2; tcfg.nA = nA; tcfg.nLambda = 0; tcfg.maxfun = 300;
    tcfg.factor = 1.02; tcfg.lambda_fixed = 1.8; tcfg.brent_lo = 0.0; tcfg.brent_hi = 10.0; tcfg.brent_xatol = 1e-5;
    const int64_t t2_ws = met2_t2_workspace_bytes(V, &tcfg);
    void* d_ws_t2 = dev_alloc<unsigned char>((size_t)t2_ws);
    double *d_fsol = dev_alloc<double>((size_t)V * nT2), *d_est = dev_alloc<double>((size_t)V * nTE);
    double *d_reg = dev_alloc<double>(V), *d_maps = dev_alloc<double>((size_t)V * 6);
    uint32_t* d_st2 = dev_alloc<uint32_t>(V);
    MK(met2_t2_fit(d_sig, d_idx, V, &tcfg, d_dic, d_dicT, d_G, d_kband, nullptr, d_logT2, d_comp, d_fsol, d_est, d_reg,
                   d_maps, d_st2, d_ws_t2, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> idx(V);
    std::vector<uint32_t> st(V), st2(V);
    std::vector<double> est((size_t)V * nTE), maps((size_t)V * 6), reg(V), fsol((size_t)V * nT2);
    CK(cudaMemcpy(idx.data(), d_idx, V * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), d_st, V * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st2.data(), d_st2, V * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(est.data(), d_est, est.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(maps.data(), d_maps, maps.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(reg.data(), d_reg, V * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fsol.data(), d_fsol, fsol.size() * 8, cudaMemcpyDeviceToHost));
    // ---- the same fit in the reduced echo space (met2_echo_basis + met2_t2_fit_echo, MET2_T2_FLAG_ECHO_SPACE): the rank is
    //      chosen from the measured residual of the reduction, as batched.Dictionary.echo_basis does
    int bad_echo = 0;
    {
        int R = met2_echo_rank(1);
        double *d_basis = dev_alloc<double>((size_t)nA * nTE * MET2_ECHO_RANK), *d_coef = dev_alloc<double>((size_t)nA * nT2 * MET2_ECHO_RANK);
        double* d_tail = dev_alloc<double>(nA);
        std::vector<double> tail(nA);
        double tmax = 0.0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            MK(met2_echo_basis(d_dic, nA, nTE, nT2, R, d_basis, d_coef, d_tail, nullptr));
            CK(cudaMemcpy(tail.data(), d_tail, nA * 8, cudaMemcpyDeviceToHost));
            tmax = 0.0;
            for (int a = 0; a < nA; ++a) tmax = std::max(tmax, tail[a]);
            if (tmax <= (R == met2_echo_rank(1) ? 4e-12 : 1e-15)) break;
            R = met2_echo_rank(0);
        }
        met2_t2_cfg ecfg = tcfg;
        ecfg.flags |= MET2_T2_FLAG_ECHO_SPACE;
        ecfg.echo_rank = R;
        double *d_fsol2 = dev_alloc<double>((size_t)V * nT2), *d_est2 = dev_alloc<double>((size_t)V * nTE);
        double *d_reg2 = dev_alloc<double>(V), *d_maps2 = dev_alloc<double>((size_t)V * 6);
        uint32_t* d_st3 = dev_alloc<uint32_t>(V);
        MK(met2_t2_fit_echo(d_sig, d_idx, V, &ecfg, d_dic, d_dicT, d_G, d_kband, nullptr, d_logT2, d_comp, d_basis, d_coef, d_fsol2,
                            d_est2, d_reg2, d_maps2, d_st3, d_ws_t2, nullptr));
        CK(cudaDeviceSynchronize());
        std::vector<double> fsol2((size_t)V * nT2);
        std::vector<uint32_t> st3(V);
        CK(cudaMemcpy(fsol2.data(), d_fsol2, fsol2.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(st3.data(), d_st3, V * 4, cudaMemcpyDeviceToHost));
        double worst = 0.0;
        for (int v = 1; v < V; ++v) {
            double fmax = 0.0, dmax = 0.0;
            bool same = (st3[v] == 0);
            for (int j = 0; j < nT2; ++j) {
                const double a = fsol[(size_t)v * nT2 + j], b = fsol2[(size_t)v * nT2 + j];
                fmax = std::max(fmax, std::fabs(a));
                dmax = std::max(dmax, std::fabs(a - b));
                if ((a > 0.0) != (b > 0.0)) same = false;
            }
            worst = std::max(worst, dmax / fmax);
            if (!same || dmax > 1e-6 * fmax) ++bad_echo;
        }
        std::printf("echo space (rank %d, residual of the reduction %.1e): %d/%d voxels differ from the Gram-domain fit, "
                    "largest relative spectrum difference %.1e\n", R, tmax, bad_echo, V - 1, worst);
    }
    // ---- checks
    if (!(st[0] & MET2_ST_SKIPPED) || !(st2[0] & MET2_ST_SKIPPED)) { std::printf("empty voxel not skipped\n"); ++fails; }
    for (int j = 0; j < nT2; ++j) if (fsol[j] != 0.0) { std::printf("empty voxel has a spectrum\n"); ++fails; break; }
    int bad_idx = 0, bad_fit = 0, bad_map = 0;
    for (int v = 1; v < V; ++v) {
        if (st[v] || st2[v]) ++fails;
        if (std::abs(idx[v] - a_true[v]) > 1) ++bad_idx;       // the generating angle (1-degree grid) up to the ripple
        double num = 0.0, den = 0.0, tot = 0.0;
        for (int e = 0; e < nTE; ++e) {
            const double d = est[(size_t)v * nTE + e] - sig[(size_t)v * nTE + e];
            num += d * d;
            den += sig[(size_t)v * nTE + e] * sig[(size_t)v * nTE + e];
        }
        if (std::sqrt(num / den) > 5e-3) ++bad_fit;            // X2 lets the residual grow to 1.02x the (ripple-sized) optimum
        const double* m = &maps[(size_t)v * 6];
        for (int j = 0; j < nT2; ++j) tot += fsol[(size_t)v * nT2 + j];
        if (std::fabs(m[0] + m[1] + m[2] - 1.0) > 0.06 || std::fabs(m[5] - (tot + 1e-16)) > 1e-9 * tot ||
            !(m[0] > 0.05 && m[0] < 0.3))
            ++bad_map;                                         // fractions sum to ~1 (bins at 40/200 ms count once/twice), TWC = sum f
    }
    std::printf("FA index: %d/%d wrong; fit: %d voxels with relative residual > 5e-3; maps: %d inconsistent\n", bad_idx,
                V - 1, bad_fit, bad_map);
    std::printf("voxel 1: FA %d (true %d), MWF %.4f, k_est %.4f; kernel launches so far: %lld\n", idx[1], a_true[1],
                maps[6], reg[1], (long long)met2_launch_count());
    fails += bad_idx + bad_fit + bad_map + bad_echo;
    std::printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
