// met2_smooth.cu — separable FP64 Gaussian smoothing of every echo volume for the FA stage.
//
// Takes over `filt.gaussian_filter(data[:, :, :, c], sig_g, 0)` with sig_g = 2 (motor/motor_recon_met2_real_data.py:336-346):
// scipy.ndimage applies correlate1d along axes 0, 1, 2 with mode 'reflect' (d c b a | a b c d | d c b a), radius
// int(4 sigma + 0.5) = 8, and for a symmetric kernel accumulates  tmp = x[c] w[r];  for ll = -r..-1:
// tmp += (x[c+ll] + x[c-ll]) * w[ll+r].  The kernel follows that order with separate multiply and add (the library is
// built with -fmad=false), so the result is bitwise equal to SciPy's (tests/test_gpu_dropin.py).
#include "met2_host.h"

namespace met2 {

__device__ __forceinline__ int reflect_index(int i, int n) {
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i;
}

// data is [n0][n1][n2][nt] C-order; filter along axis (0, 1 or 2); one thread per element.
__global__ void gauss1d_kernel(const double* __restrict__ in, double* __restrict__ out, int n0, int n1, int n2, int nt,
                               int axis, const double* __restrict__ w, int r) {
    const long long total = (long long)n0 * n1 * n2 * nt;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    long long rem = e;
    const int t = (int)(rem % nt); rem /= nt;
    const int i2 = (int)(rem % n2); rem /= n2;
    const int i1 = (int)(rem % n1); rem /= n1;
    const int i0 = (int)rem;
    (void)t;
    const long long s2 = nt, s1 = (long long)n2 * nt, s0 = (long long)n1 * n2 * nt;
    int c, n;
    long long stride;
    if (axis == 0) { c = i0; n = n0; stride = s0; }
    else if (axis == 1) { c = i1; n = n1; stride = s1; }
    else { c = i2; n = n2; stride = s2; }
    const long long base = e - (long long)c * stride;
    double tmp = in[base + (long long)c * stride] * w[r];
    for (int ll = -r; ll < 0; ++ll) {
        const double a = in[base + (long long)reflect_index(c + ll, n) * stride];
        const double b = in[base + (long long)reflect_index(c - ll, n) * stride];
        tmp = tmp + (a + b) * w[ll + r];
    }
    out[e] = tmp;
}

}  // namespace met2

using namespace met2;

extern "C" int met2_gaussian_smooth(const double* vol, int nx, int ny, int nz, int nt, const double* weights, int radius,
                                    double* out, double* tmp, void* stream) {
    if (!vol || !weights || !out || !tmp || nx <= 0 || ny <= 0 || nz <= 0 || nt <= 0 || radius < 0)
        return set_error(MET2_ERR_ARG, "met2_gaussian_smooth: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)nx * ny * nz * nt;
    const int tb = 256;
    const unsigned nb = (unsigned)((total + tb - 1) / tb);
    MET2_LAUNCH(nb, tb, 0, st, gauss1d_kernel)(vol, out, nx, ny, nz, nt, 0, weights, radius);
    MET2_LAUNCH(nb, tb, 0, st, gauss1d_kernel)(out, tmp, nx, ny, nz, nt, 1, weights, radius);
    MET2_LAUNCH(nb, tb, 0, st, gauss1d_kernel)(tmp, out, nx, ny, nz, nt, 2, weights, radius);
    count_launch(3);
    return check_launch("gauss1d_kernel");
}
