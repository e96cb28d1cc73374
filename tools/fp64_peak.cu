// FP64 DFMA peak micro-benchmark for the roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent-chain latency: one warp, one chain
__global__ void dfma_latency(double* out, int iters, double a, double b, long long* cyc) {
    double acc = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b);
        acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void shfl_latency(double* out, int iters, long long* cyc) {
    double acc = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4); acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void ddiv_latency(double* out, int iters, double a, long long* cyc) {
    double acc = threadIdx.x + 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        acc = a / acc + 1.0; acc = a / acc + 1.0; acc = a / acc + 1.0; acc = a / acc + 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 32 * 256 * 4);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16;
    const int ILP = 8;
    for (int blocks_per_sm : {1, 2, 4, 8}) {
        int grid = p.multiProcessorCount * blocks_per_sm;
        dfma_kernel<ILP><<<grid, 256>>>(out, 1024, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            dfma_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double flops = 2.0 * grid * 256.0 * ILP * iters;
        printf(" \"dfma_tflops_%dblk\": %.3f,\n", blocks_per_sm, flops / (best * 1e-3) / 1e12);
    }
    dfma_latency<<<1, 32>>>(out, 4096, 1.0000001, 1e-9, cyc); cudaDeviceSynchronize();
    printf(" \"dfma_dep_latency_cyc\": %.2f,\n", (double)*cyc / (4096.0 * 8));
    shfl_latency<<<1, 32>>>(out, 4096, cyc); cudaDeviceSynchronize();
    printf(" \"shfl_dadd_dep_latency_cyc\": %.2f,\n", (double)*cyc / (4096.0 * 4));
    ddiv_latency<<<1, 32>>>(out, 4096, 3.0, cyc); cudaDeviceSynchronize();
    printf(" \"ddiv_dadd_dep_latency_cyc\": %.2f,\n", (double)*cyc / (4096.0 * 4));
    printf(" \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
