// met2_host.h — host-side plumbing shared by the translation units of libmet2.so (error state, launch accounting).
#pragma once
#ifdef MET2_HOST_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/met2.h"

// Kernel launch: `MET2_LAUNCH(grid, block, dynamic_smem_bytes, stream, kernel<...>)(args...)` is the CUDA launch
// `kernel<...><<<grid, block, smem, stream>>>(args...)`; under MET2_HOST_EMU (tests/emu, CPU) the same host code drives
// the SIMT emulator instead.  The kernel name comes last so that template argument lists may contain commas.
#ifdef MET2_HOST_EMU
#define MET2_LAUNCH(grid, block, smem, stream, ...) simt::make_launch(dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#else
#define MET2_LAUNCH(grid, block, smem, stream, ...) __VA_ARGS__<<<(grid), (block), (smem), (stream)>>>
#endif

namespace met2 {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);
int sm_count();
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace met2
