// Instantiation unit: float, K=1, T=3 (see lhvi_spec_impl.cuh).
#include "lhvi_spec_impl.cuh"

namespace lhvi {
int spec_f32_k1(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    return launch_kt<float, 1, 3>(m, g, row0, s);
}
}  // namespace lhvi
