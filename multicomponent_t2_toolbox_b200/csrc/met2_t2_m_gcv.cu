// met2_t2_m_gcv.cu — instantiates the T2 fit kernel for reg_method = GCV (all nT2 / nTE size classes).
#include "met2_t2_impl.cuh"

namespace met2 {
int t2_launch_gcv(const T2Args& A, const T2Geom& g, cudaStream_t st) { return t2_launch_method<MET2_REG_GCV>(A, g, st); }
}  // namespace met2
