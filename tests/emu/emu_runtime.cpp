// emu_runtime.cpp — TEST INFRASTRUCTURE ONLY: the globals of the SIMT emulator (simt_emu.h) and a counter read-out.
// tests/emu/emu.py compiles every csrc/*.cu with g++ -DMET2_HOST_EMU and links them with this file into
// tests/emu/_build/libmet2_emu.so, which exports the same C ABI as libmet2.so (include/met2.h) but takes HOST pointers
// and runs the kernels' own source on the CPU.  Nothing in the product package loads it.
#define MET2_HOST_EMU 1
#include "simt_emu.h"

namespace simt {
Block* g_block = nullptr;
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
long long n_fma = 0, n_smem = 0, n_syncwarp = 0, n_collectives = 0, n_launches = 0;
constexpr size_t S_DOUBLES = 40960;   // 320 KB: more than any kernel's dynamic shared memory
alignas(16) static double s_storage[S_DOUBLES];
Shared g_shared = {s_storage, (long)S_DOUBLES};
}  // namespace simt

extern "C" {
// work counters since the last call, per THREAD (divide by 32 for warp level):
// [fma, shared-memory accesses, __syncwarp, warp collectives, kernel launches]
void emu_counters(long long* out) {
    out[0] = simt::n_fma;
    out[1] = simt::n_smem;
    out[2] = simt::n_syncwarp;
    out[3] = simt::n_collectives;
    out[4] = simt::n_launches;
    simt::n_fma = simt::n_smem = simt::n_syncwarp = simt::n_collectives = simt::n_launches = 0;
}
}
