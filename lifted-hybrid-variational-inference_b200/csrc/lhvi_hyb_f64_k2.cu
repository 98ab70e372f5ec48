// Instantiation unit of the kernels for records with hidden discrete arguments / quadrature degree 10:
// double, K=2 (see lhvi_hyb_impl.cuh).
#include "lhvi_hyb_impl.cuh"

namespace lhvi {
int hyb_f64_k2(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    return launch_hyb_k<double, 2>(m, g, row0, s);
}
}  // namespace lhvi
