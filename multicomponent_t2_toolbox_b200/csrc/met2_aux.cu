// met2_aux.cu — the two "next" rows of SURVEY.md §8f that sit either side of the voxel fit:
//
//  (1) met2_segment_means: per-segment mean signal and mean kernel (mean of the dictionary slices selected by each
//      voxel's FA index), the input of the reference's mean-spectrum diagnostics
//      (motor/motor_recon_met2_real_data.py:377-403: one segment = mask == 1) and of its ROI-based estimator
//      (motor/motor_recon_met2_real_data_ROI.py:405-445: one segment per ROI label).  The reference walks the volume in
//      a triple Python loop per ROI; here one pass accumulates every segment at once.
//
//  (2) met2_nesma_filter: the NESMA denoiser of motor...:305-333 — for every voxel with mask == 1, the mean of the
//      signals in its [-6, +6) neighbourhood whose relative L1 distance to the voxel's own signal is below 2.5 %.
//      The reference is a pure-Python O(V * 1728 * nTE) loop; here one thread owns one voxel, keeps its own signal and
//      the running mean in registers and walks the window over an echo-major copy of the volume so that a warp's loads
//      are contiguous.  Summation orders follow NumPy's (pairwise 8-way for the contiguous reductions, sequential over
//      the window for the mean), so the result is bitwise equal to the NumPy restatement wherever the 2.5 % decisions
//      agree (tests/test_gpu_dropin.py).
#include "met2_host.h"

namespace met2 {

// ------------------------------------------------------------------------------------------------ segment means
// One warp per run of SEG_RUN consecutive voxels; lane = echo (+32 for the second slot).  Signals of consecutive
// voxels of the same segment are summed in registers and flushed with one atomicAdd per echo when the segment changes.
constexpr int SEG_RUN = 128;

__global__ void seg_accumulate_kernel(const double* __restrict__ sig, const int* __restrict__ fa_index,
                                      const int* __restrict__ label, long long V, int nTE, int nA, int nSeg,
                                      double* __restrict__ sum_signal, int* __restrict__ hist, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long v0 = warp * SEG_RUN;
    if (v0 >= V) return;
    const long long v1 = (v0 + SEG_RUN < V) ? v0 + SEG_RUN : V;
    int cur = -1, run = 0;
    double a0 = 0.0, a1 = 0.0;
    auto flush = [&]() {
        if (cur >= 0) {
            if (lane < nTE) atomicAdd(sum_signal + (size_t)cur * nTE + lane, a0);
            if (lane + 32 < nTE) atomicAdd(sum_signal + (size_t)cur * nTE + lane + 32, a1);
            if (lane == 0) atomicAdd(counts + cur, run);
        }
        a0 = a1 = 0.0;
        run = 0;
    };
    for (long long v = v0; v < v1; ++v) {
        int s = label[v];
        if (s < 0 || s >= nSeg) s = -1;
        if (s != cur) {
            flush();
            cur = s;
        }
        if (s >= 0) {
            if (lane < nTE) a0 += sig[v * nTE + lane];
            if (lane + 32 < nTE) a1 += sig[v * nTE + lane + 32];
            ++run;
            if (lane == 0) {
                const int fa = fa_index[v];
                if (fa >= 0 && fa < nA) atomicAdd(hist + (size_t)s * nA + fa, 1);
            }
        }
    }
    flush();
}

// mean_signal = sum / count; mean_kernel[s] = sum_a hist[s][a] * dic[a] / count  (one thread per output element)
__global__ void seg_finish_kernel(const double* __restrict__ dic, const int* __restrict__ hist,
                                  const int* __restrict__ counts, int nTE, int nT2, int nA, int nSeg,
                                  double* __restrict__ mean_signal, double* __restrict__ mean_kernel) {
    const int per = nTE * nT2;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)nSeg * per) return;
    const int s = (int)(e / per), r = (int)(e - (long long)s * per);
    const double nv = (double)counts[s];
    double acc = 0.0;
    for (int a = 0; a < nA; ++a) {
        const int h = hist[(size_t)s * nA + a];
        if (h) acc = fma((double)h, dic[(size_t)a * per + r], acc);
    }
    mean_kernel[e] = acc / nv;
    if (r < nTE) mean_signal[(size_t)s * nTE + r] = mean_signal[(size_t)s * nTE + r] / nv;
}

// ------------------------------------------------------------------------------------------------ NESMA
// [V3][nt] -> [nt][V3] through a 32 x 32 shared-memory tile
__global__ void to_echo_major_kernel(const double* __restrict__ in, double* __restrict__ out, long long V3, int nt) {
    __shared__ double tile[32][33];
    const long long vb = (long long)blockIdx.x * 32;
    const int tb = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long v = vb + r;
        const int t = tb + threadIdx.x;
        tile[r][threadIdx.x] = (v < V3 && t < nt) ? in[v * nt + t] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = tb + r;
        const long long v = vb + threadIdx.x;
        if (v < V3 && t < nt) out[(size_t)t * V3 + v] = tile[threadIdx.x][r];
    }
}

// NumPy's pairwise sum of a contiguous run of n < 128 doubles (numpy/_core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum): eight interleaved accumulators over blocks of 8, combined as a balanced tree, then the tail.
template <int NT>
__device__ __forceinline__ double numpy_sum(const double (&a)[NT], int n) {
    if (n < 8) {
        double r = 0.0;   // numpy starts from the identity for short runs: res = 0.; res += a[i]
#pragma unroll
        for (int i = 0; i < NT; ++i)
            if (i < n) r += a[i];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    const int nb = n - (n % 8);
#pragma unroll
    for (int i = 8; i < NT; i += 8) {
        if (i < nb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        }
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int i = 8; i < NT; ++i)
        if (i >= nb && i < n) res += a[i];
    return res;
}

template <int NT>
__global__ void __launch_bounds__(128) nesma_kernel(const double* __restrict__ T, const int* __restrict__ mask, int nx,
                                                    int ny, int nz, int nt, int hw, double thr,
                                                    double* __restrict__ out) {
    const long long V3 = (long long)nx * ny * nz;
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V3) return;
    double* o = out + v * nt;
    if (mask[v] != 1) {   // data_den stays zero outside the mask (motor...:306,316)
        for (int t = 0; t < nt; ++t) o[t] = 0.0;
        return;
    }
    const int z = (int)(v % nz);
    const int y = (int)((v / nz) % ny);
    const int x = (int)(v / ((long long)nz * ny));
    double own[NT], acc[NT], d[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        own[t] = (t < nt) ? T[(size_t)t * V3 + v] : 0.0;
        acc[t] = 0.0;
    }
    const double ssum = numpy_sum<NT>(own, nt);
    // window [c - hw, c + hw) clipped to the volume: np.max([c - 6, 0]) : np.min([c + 6, n]) (motor...:311-320)
    const int x0 = max(x - hw, 0), x1 = min(x + hw, nx);
    const int y0 = max(y - hw, 0), y1 = min(y + hw, ny);
    const int z0 = max(z - hw, 0), z1 = min(z + hw, nz);
    int nvalid = 0;
    for (int xx = x0; xx < x1; ++xx)
        for (int yy = y0; yy < y1; ++yy) {
            const long long rowq = ((long long)xx * ny + yy) * nz;
            for (int zz = z0; zz < z1; ++zz) {
                const double* q = T + rowq + zz;
#pragma unroll
                for (int t = 0; t < NT; ++t) d[t] = (t < nt) ? fabs(q[(size_t)t * V3] - own[t]) : 0.0;
                const double re = 100.0 * numpy_sum<NT>(d, nt) / ssum;
                if (re < thr) {
                    ++nvalid;
#pragma unroll
                    for (int t = 0; t < NT; ++t)
                        if (t < nt) acc[t] += q[(size_t)t * V3];
                }
            }
        }
    const double nvd = (double)nvalid;   // np.mean over an empty selection is NaN (0/0)
#pragma unroll
    for (int t = 0; t < NT; ++t)
        if (t < nt) o[t] = acc[t] / nvd;
}

}  // namespace met2

using namespace met2;

extern "C" int64_t met2_segment_workspace_bytes(int nSeg, int nA) {
    if (nSeg < 0 || nA < 0) return 0;
    return (int64_t)align256((size_t)nSeg * nA * sizeof(int));
}

extern "C" int met2_segment_means(const double* sig, const int32_t* fa_index, const int32_t* label, int64_t V, int nTE,
                                  int nT2, int nA, int nSeg, const double* dic, double* mean_signal, double* mean_kernel,
                                  int32_t* counts, void* workspace, void* stream) {
    if (V < 0 || nSeg <= 0 || nTE <= 0 || nTE > MET2_MAX_NTE || nT2 <= 0 || nA <= 0 || !dic || !mean_signal ||
        !mean_kernel || !counts || !workspace || (V > 0 && (!sig || !fa_index || !label)))
        return set_error(MET2_ERR_ARG, "met2_segment_means: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int* hist = (int*)workspace;
    cudaMemsetAsync(hist, 0, (size_t)nSeg * nA * sizeof(int), st);
    cudaMemsetAsync(counts, 0, (size_t)nSeg * sizeof(int), st);
    cudaMemsetAsync(mean_signal, 0, (size_t)nSeg * nTE * sizeof(double), st);
    if (V > 0) {
        const long long warps = (V + SEG_RUN - 1) / SEG_RUN;
        const int tb = 128;
        const unsigned nb = (unsigned)((warps * 32 + tb - 1) / tb);
        MET2_LAUNCH(nb, tb, 0, st, seg_accumulate_kernel)(sig, fa_index, label, V, nTE, nA, nSeg, mean_signal, hist, counts);
        count_launch();
    }
    const long long total = (long long)nSeg * nTE * nT2;
    MET2_LAUNCH((unsigned)((total + 255) / 256), 256, 0, st, seg_finish_kernel)(dic, hist, counts, nTE, nT2, nA, nSeg, mean_signal,
                                                                        mean_kernel);
    count_launch();
    return check_launch("met2_segment_means");
}

extern "C" int met2_nesma_filter(const double* vol, const int32_t* mask, int nx, int ny, int nz, int nt, int half_window,
                                 double threshold_percent, double* out, double* tmp, void* stream) {
    if (!vol || !mask || !out || !tmp || nx <= 0 || ny <= 0 || nz <= 0 || nt <= 0 || nt > MET2_MAX_NTE || half_window < 0)
        return set_error(MET2_ERR_ARG, "met2_nesma_filter: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long V3 = (long long)nx * ny * nz;
    dim3 tgrid((unsigned)((V3 + 31) / 32), (unsigned)((nt + 31) / 32));
    MET2_LAUNCH(tgrid, dim3(32, 8), 0, st, to_echo_major_kernel)(vol, tmp, V3, nt);
    const unsigned nb = (unsigned)((V3 + 127) / 128);
    if (nt <= 32)
        MET2_LAUNCH(nb, 128, 0, st, nesma_kernel<32>)(tmp, mask, nx, ny, nz, nt, half_window, threshold_percent, out);
    else
        MET2_LAUNCH(nb, 128, 0, st, nesma_kernel<64>)(tmp, mask, nx, ny, nz, nt, half_window, threshold_percent, out);
    count_launch(2);
    return check_launch("met2_nesma_filter");
}
