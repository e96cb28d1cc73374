// met2_t2_m_lcurve.cu — instantiates the T2 fit kernel for reg_method = lcurve (all nT2 / nTE size classes).
#include "met2_t2_impl.cuh"

namespace met2 {
int t2_launch_lcurve(const T2Args& A, const T2Geom& g, cudaStream_t st) { return t2_launch_method<MET2_REG_LCURVE>(A, g, st); }
}  // namespace met2
