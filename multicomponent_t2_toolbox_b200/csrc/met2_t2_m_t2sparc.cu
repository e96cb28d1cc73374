// met2_t2_m_t2sparc.cu — instantiates the T2 fit kernel for reg_method = t2sparc (all nT2 / nTE size classes).
#define MET2_UNROLL_WIDE_NS 3   // 96 T2 bins: three warps per SM, unrolled mat-vec loops (met2_nnls.cuh)
#include "met2_t2_impl.cuh"

namespace met2 {
int t2_launch_t2sparc(const T2Args& A, const T2Geom& g, cudaStream_t st) { return t2_launch_method<MET2_REG_T2SPARC>(A, g, st); }
}  // namespace met2
