// Instantiation unit of the persistent iteration kernel: float, K=3, T=3 (see lhvi_iter_impl.cuh).
#include "lhvi_iter_impl.cuh"

namespace lhvi {
int iter_f32_k3(const lhvi_model* m, const lhvi_group* groups, int n_groups, const lhvi_exchange* x,
                const lhvi_optim* o, int n_iter, int probe_only, cudaStream_t s) {
    return launch_iterate_kt<float, 3, 3>(m, groups, n_groups, x, o, n_iter, probe_only, s);
}
}  // namespace lhvi
