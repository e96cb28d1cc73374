// simt_emu.h — TEST INFRASTRUCTURE ONLY.  A single-threaded SIMT emulator that lets g++ compile and run the device code
// of multicomponent_t2_toolbox_b200/csrc/*.cuh|*.cu on the CPU, so that the kernels' control flow, indexing and warp
// collectives can be exercised in the `-m "not gpu"` tests (this container has no GPU).  It is never linked into
// libmet2.so: the csrc headers include it only under MET2_HOST_EMU, which only tests/emu/*.cpp define.
//
// Model: ONE thread block; every CUDA thread is a ucontext fiber; the scheduler runs the fibers round-robin and a fiber
// yields whenever it waits at a barrier.  Warp collectives (__shfl*_sync, __reduce_*_sync, votes, the FP64 MMA) are
// "deposit -> warp barrier -> read -> warp barrier" on a per-warp exchange buffer, which is exact for the code under
// test: every collective is called with the full mask from warp-uniform control flow.  Shared memory is one global
// array; `__shared__` locals become statics (one block at a time).  Arithmetic is IEEE double with explicit std::fma,
// compiled with -ffp-contract=off like the library's -fmad=false.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <execinfo.h>
#include <ucontext.h>

#include <cmath>
#include <functional>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

typedef void* cudaStream_t;
struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef dim3 emu_dim3;

// the few CUDA runtime calls the host side of csrc/ makes
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    int multiProcessorCount;
};
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
    memset(p, v, n);
    return cudaSuccess;
}
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) {
    memcpy(dst, src, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) {
    *d = 0;
    return cudaSuccess;
}
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
    const char* ev = getenv("SIMT_EMU_SMS");
    *v = ev ? atoi(ev) : 2;   // "SMs" of the emulated device: sizes the persistent grids
    return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    const char* ev = getenv("SIMT_EMU_SMS");
    p->multiProcessorCount = ev ? atoi(ev) : 2;   // "SMs" of the emulated device: sizes the persistent grids
    return cudaSuccess;
}

namespace simt {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;   // from the process-wide pool below (allocated once, never zeroed, reused by every block)
    int tid = 0;
    bool done = false;
};
constexpr size_t FIBER_STACK = 512 * 1024;
inline char* fiber_stack(int t) {
    static std::vector<char*> pool;
    while ((int)pool.size() <= t) pool.push_back(static_cast<char*>(malloc(FIBER_STACK)));
    return pool[t];
}

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = 0;
    int nthreads = 0;
    // barriers: generation counters
    int cta_arrived = 0, cta_gen = 0;
    std::vector<int> warp_arrived, warp_gen;
    std::vector<uint64_t> xbuf;   // [warp][32][2] exchange slots
    std::function<void()> body;
    long long collectives = 0;
};

extern Block* g_block;
extern emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

// Work counters (per THREAD events; divide by 32 for warp-level instruction estimates): explicit fma() calls, accesses to
// the shared-memory array S, warp barriers.  A cost model for comparing kernel variants without a GPU.
extern long long n_fma, n_smem, n_syncwarp;
struct Shared;
extern Shared g_shared;   // the block's dynamic shared memory (csrc: `S`)

// The dynamic shared memory: S[i] and S + off like the device array, with access counting.
struct Shared {
    double* data;
    long size;    // doubles the kernel under test may touch (its dynamic shared-memory size): checked on every access
    void check(long i) const {
        if (i < 0 || i >= size) {
            fprintf(stderr, "simt_emu: shared-memory access out of bounds: S[%ld], size %ld doubles (thread %u)\n", i, size,
                    g_threadIdx.x);
            void* bt[16];
            backtrace_symbols_fd(bt, backtrace(bt, 16), 2);   // resolve with addr2line -e tests/emu/_build/libmet2_emu.so
            abort();
        }
    }
    double& operator[](long i) {
        ++n_smem;
        check(i);
        return data[i];
    }
    double* operator+(long off) {
        ++n_smem;
        check(off);
        return data + off;
    }
};

inline int lane_id() { return g_block->cur & 31; }
inline int warp_id() { return g_block->cur >> 5; }

inline void yield() {
    Block* b = g_block;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

inline void warp_barrier() {
    Block* b = g_block;
    const int w = warp_id();
    const int gen = b->warp_gen[w];
    if (++b->warp_arrived[w] == 32) {
        b->warp_arrived[w] = 0;
        ++b->warp_gen[w];
        return;
    }
    while (b->warp_gen[w] == gen) yield();
}

inline void cta_barrier() {
    Block* b = g_block;
    const int gen = b->cta_gen;
    if (++b->cta_arrived == b->nthreads) {
        b->cta_arrived = 0;
        ++b->cta_gen;
        return;
    }
    while (b->cta_gen == gen) yield();
}

inline uint64_t* slot(int lane, int k = 0) { return &g_block->xbuf[((size_t)warp_id() * 32 + lane) * 2 + k]; }

template <class T>
inline uint64_t to_bits(T v) {
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T from_bits(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}

// every lane deposits v; returns the value deposited by lane `src`
template <class T>
inline T exchange(T v, int src) {
    ++g_block->collectives;
    *slot(lane_id()) = to_bits(v);
    warp_barrier();
    T r = from_bits<T>(*slot(src & 31));
    warp_barrier();
    return r;
}

template <class T, class F>
inline T reduce_all(T v, F f) {
    ++g_block->collectives;
    *slot(lane_id()) = to_bits(v);
    warp_barrier();
    T r = from_bits<T>(*slot(0));
    for (int l = 1; l < 32; ++l) r = f(r, from_bits<T>(*slot(l)));
    warp_barrier();
    return r;
}

static void fiber_entry() {
    Block* b = g_block;
    b->body();
    b->fibers[b->cur].done = true;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

// Run `body` as one thread block of `nthreads` threads (a multiple of 32).
inline long long run_block(int nthreads, int block_index, int grid, std::function<void()> body, dim3 bdim = dim3(0)) {
    Block blk;
    g_block = &blk;
    blk.nthreads = nthreads;
    blk.body = body;
    blk.fibers.resize(nthreads);
    if (nthreads % 32) {
        fprintf(stderr, "simt_emu: block size %d is not a multiple of 32\n", nthreads);
        abort();
    }
    blk.warp_arrived.assign(nthreads / 32, 0);
    blk.warp_gen.assign(nthreads / 32, 0);
    blk.xbuf.assign((size_t)nthreads * 2, 0);
    g_blockIdx = dim3((unsigned)block_index, 0, 0);
    g_blockDim = bdim.x ? bdim : dim3((unsigned)nthreads, 1, 1);
    g_gridDim = dim3((unsigned)grid, 1, 1);
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = blk.fibers[t];
        f.tid = t;
        f.stack = fiber_stack(t);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = FIBER_STACK;
        f.ctx.uc_link = &blk.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    // Scheduling order of the fibers inside a round.  A kernel whose results depend on it has a missing barrier (a lane
    // reading what another lane of the same phase writes): SIMT_EMU_ORDER=reverse runs the lanes 31..0, =shuffle draws a
    // new pseudo-random permutation every round.  The tests run every kernel under all three orders.
    std::vector<int> order(nthreads);
    for (int t = 0; t < nthreads; ++t) order[t] = t;
    const char* ord = getenv("SIMT_EMU_ORDER");
    const bool reverse = ord && !strcmp(ord, "reverse"), shuffle = ord && !strcmp(ord, "shuffle");
    if (reverse)
        for (int t = 0; t < nthreads; ++t) order[t] = nthreads - 1 - t;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    int live = nthreads;
    long long idle_rounds = 0;
    while (live > 0) {
        int progressed = 0;
        if (shuffle) {
            for (int t = nthreads - 1; t > 0; --t) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const int u = (int)((rng >> 33) % (uint64_t)(t + 1));
                const int tmp = order[t];
                order[t] = order[u];
                order[u] = tmp;
            }
        }
        for (int oi = 0; oi < nthreads; ++oi) {
            const int t = order[oi];
            Fiber& f = blk.fibers[t];
            if (f.done) continue;
            blk.cur = t;
            g_threadIdx = dim3((unsigned)t % g_blockDim.x, ((unsigned)t / g_blockDim.x) % g_blockDim.y,
                               (unsigned)t / (g_blockDim.x * g_blockDim.y));
            swapcontext(&blk.sched, &f.ctx);
            ++progressed;
            if (f.done) --live;
        }
        if (!progressed) break;
        if (++idle_rounds > 2000000000LL) {
            fprintf(stderr, "simt_emu: scheduler ran away\n");
            abort();
        }
    }
    g_block = nullptr;
    return blk.collectives;
}

extern long long n_collectives, n_launches;

// `MET2_LAUNCH(grid, block, smem, stream, kernel)(args...)` of csrc/met2_host.h: runs the blocks of a 1-D grid one after
// the other, each as run_block; the dynamic shared memory is bounds-checked against the size the host code asked for.
template <class... P>
struct Launch {
    dim3 grid, block;
    size_t smem;
    void (*kern)(P...);
    template <class... Q>
    void operator()(Q&&... args) const {
        if (grid.y != 1 || grid.z != 1) {
            fprintf(stderr, "simt_emu: only 1-D grids are emulated\n");
            abort();
        }
        const int nthreads = (int)(block.x * block.y * block.z);
        g_shared.size = (long)(smem / sizeof(double));
        ++n_launches;
        for (unsigned b = 0; b < grid.x; ++b) {
            // a block starts with its dynamic shared memory POISONED (NaN): a kernel that reads a shared location
            // before writing it shows up as NaN / wrong outputs instead of silently using the previous block's data
            for (long i = 0; i < g_shared.size; ++i) g_shared.data[i] = NAN;
            n_collectives += run_block(nthreads, (int)b, (int)grid.x, [&]() { kern(args...); }, block);
        }
    }
};
template <class... P>
inline Launch<P...> make_launch(dim3 grid, dim3 block, size_t smem, void (*kern)(P...)) {
    return Launch<P...>{grid, block, smem, kern};
}

}  // namespace simt

#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)

// ------------------------------------------------------------------------------------------------ intrinsics
inline void __syncwarp(unsigned = 0xffffffffu) {
    ++simt::n_syncwarp;
    simt::warp_barrier();
}
inline void __syncthreads() { simt::cta_barrier(); }

template <class T>
inline T __ldg(const T* p) { return *p; }

template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange<T>(v, src); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) { return simt::exchange<T>(v, simt::lane_id() ^ o); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) {
    const int l = simt::lane_id();
    return simt::exchange<T>(v, (l - o >= 0) ? l - o : l);
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    return simt::reduce_all<unsigned>(v, [](unsigned a, unsigned b) { return a > b ? a : b; });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    return simt::reduce_all<unsigned>(v, [](unsigned a, unsigned b) { return a < b ? a : b; });
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    return simt::reduce_all<unsigned>(v, [](unsigned a, unsigned b) { return a + b; });
}
inline int __reduce_add_sync(unsigned, int v) {
    return simt::reduce_all<int>(v, [](int a, int b) { return a + b; });
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    ++simt::g_block->collectives;
    *simt::slot(simt::lane_id()) = pred ? 1u : 0u;
    simt::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (*simt::slot(l)) r |= 1u << l;
    simt::warp_barrier();
    return r;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// CUDA's global-scope integer min / max
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline long long __double_as_longlong(double d) { return simt::from_bits<long long>(simt::to_bits(d)); }
inline double __longlong_as_double(long long v) { return simt::from_bits<double>(simt::to_bits(v)); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
template <class T>
inline T atomicAdd(T* p, T v) {   // one fiber runs at a time: plain read-modify-write is atomic here
    T o = *p;
    *p = o + v;
    return o;
}
inline unsigned atomicOr(unsigned* p, unsigned v) {
    unsigned o = *p;
    *p = o | v;
    return o;
}
inline int atomicOr(int* p, int v) {
    int o = *p;
    *p = o | v;
    return o;
}
using std::isfinite;
namespace met2 {   // unqualified fma() inside namespace met2 resolves here: counted, then the correctly rounded std::fma
inline double fma(double a, double b, double c) {
    ++simt::n_fma;
    return std::fma(a, b, c);
}
}  // namespace met2

// FP64 MMA m8n8k4: D(8x8) += A(8x4) B(4x8); fragments A[lane/4][lane%4], B[lane%4][lane/4], C/D[lane/4][2*(lane%4)+{0,1}]
inline void emu_dmma884(double& d0, double& d1, double a, double b) {
    ++simt::g_block->collectives;
    const int lane = simt::lane_id();
    *simt::slot(lane, 0) = simt::to_bits(a);
    *simt::slot(lane, 1) = simt::to_bits(b);
    simt::warp_barrier();
    const int g = lane >> 2, q = lane & 3;
    for (int k = 0; k < 4; ++k) {
        const double av = simt::from_bits<double>(*simt::slot(4 * g + k, 0));
        const double b0 = simt::from_bits<double>(*simt::slot(4 * (2 * q) + k, 1));
        const double b1 = simt::from_bits<double>(*simt::slot(4 * (2 * q + 1) + k, 1));
        d0 = std::fma(av, b0, d0);
        d1 = std::fma(av, b1, d1);
    }
    simt::warp_barrier();
}
