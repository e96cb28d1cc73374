// met2_t2_echo_r24.cu — the reduced-echo-space kernels (met2_t2_echo_impl.cuh) at rank 24.
#define MET2_ECHO_RD 24
#define MET2_ECHO_NS echo24
#define MET2_ECHO_LAUNCH t2_launch_echo_r24
#include "met2_t2_echo_impl.cuh"
