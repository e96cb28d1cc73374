// met2_t2_echo_reg_r24.cu — the reduced-echo-space L-curve / BayesReg kernel (met2_t2_echo_reg_impl.cuh) at rank 24.
#define MET2_ECHO_RD 24
#define MET2_ECHO_NS echo24
#define MET2_ECHO_PART 2
#define MET2_ECHO_LAUNCH t2_launch_echo_reg_r24
#include "met2_t2_echo_impl.cuh"
