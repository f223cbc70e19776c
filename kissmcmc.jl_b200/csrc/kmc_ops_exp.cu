// Kernel instantiations of the "exponential" plugin (README.md:15), d in {1..6, 8, 10, 12, 16}.
#define KMC_OPS_IMPL
#include "kmc_ops.cuh"

namespace kmc_host {
bool ops_exponential(int d, Ops &o) { return ops_for_dim<kmc::Exponential>(d, o); }
}  // namespace kmc_host
