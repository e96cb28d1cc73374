// Probe: does mma.sync m8n8k4 f64 run on sm_100a, what is its fragment layout and dependent latency?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void probe(double* out, long long* cyc, int iters) {
    int lane = threadIdx.x;
    // A[8x4]: A[r][c] = 10*r + c ; B[4x8]: B[r][c] = (r==c%4) ? 1 : 0 plus 100*(c>=4)...
    int ar = lane / 4, ac = lane % 4;      // A: row = groupID, col = threadID_in_group
    int br = lane % 4, bc = lane / 4;      // B: row = threadID_in_group, col = groupID
    double a = 10.0 * ar + ac;
    double b = (br == (bc % 4)) ? 1.0 : 0.0;
    if (bc >= 4) b *= 2.0;
    double d0 = 0, d1 = 0;
    dmma(d0, d1, a, b, 0.0, 0.0);
    // D[r][c]: r = lane/4, c = 2*(lane%4) + {0,1}
    out[lane * 2] = d0; out[lane * 2 + 1] = d1;
    long long t0 = clock64();
    double c0 = d0, c1 = d1;
    for (int i = 0; i < iters; ++i) { dmma(c0, c1, a, b, c0, c1); }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[64 + lane] = c0 + c1;
    // throughput: 8 independent chains
    double e[8][2];
    for (int q = 0; q < 8; ++q) { e[q][0] = q; e[q][1] = -q; }
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) dmma(e[q][0], e[q][1], a, b, e[q][0], e[q][1]);
    }
    t1 = clock64();
    if (lane == 0) cyc[1] = t1 - t0;
    double s = 0; for (int q = 0; q < 8; ++q) s += e[q][0] + e[q][1];
    out[96 + lane] = s;
}
int main() {
    double* out; long long* cyc;
    cudaMallocManaged(&out, 256 * sizeof(double)); cudaMallocManaged(&cyc, 16);
    probe<<<1, 32>>>(out, cyc, 1024);
    cudaError_t e = cudaDeviceSynchronize();
    printf("err=%s\n", cudaGetErrorString(e));
    // expected D = A*B: D[r][c] = sum_k A[r][k]*B[k][c] = A[r][c%4] * (c>=4?2:1)
    int bad = 0;
    for (int lane = 0; lane < 32; ++lane) for (int h = 0; h < 2; ++h) {
        int r = lane / 4, c = 2 * (lane % 4) + h;
        double exp = (10.0 * r + (c % 4)) * (c >= 4 ? 2.0 : 1.0);
        if (out[lane * 2 + h] != exp) { bad++; if (bad < 5) printf("mismatch lane %d h %d got %g exp %g\n", lane, h, out[lane*2+h], exp); }
    }
    printf("layout mismatches=%d  dep latency=%.1f cyc  8-chain issue interval=%.1f cyc/mma\n", bad, cyc[0] / 1024.0, cyc[1] / (1024.0 * 8));
    return 0;
}
