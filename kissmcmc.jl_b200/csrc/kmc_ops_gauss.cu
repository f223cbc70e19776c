// Kernel instantiations of the fused FP64 "gaussian" plugin (test/runtests.jl:53,61), d in {1..6, 8, 10, 12, 16}.
#define KMC_OPS_IMPL
#include "kmc_ops.cuh"

namespace kmc_host {
bool ops_gaussian(int d, Ops &o) { return ops_for_dim<kmc::Gaussian>(d, o); }
}  // namespace kmc_host
