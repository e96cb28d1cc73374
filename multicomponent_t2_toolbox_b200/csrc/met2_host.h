// met2_host.h — host-side plumbing shared by the translation units of libmet2.so (error state, launch accounting).
#pragma once
#ifdef MET2_HOST_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/met2.h"

namespace met2 {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);
int sm_count();
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace met2
