// Kernel instantiations of the "rosenbrock" (test/runtests.jl:68, d = 2) and "lognormal" (:56, d = 1) plugins.
#define KMC_OPS_IMPL
#include "kmc_ops.cuh"

namespace kmc_host {
bool ops_rosenbrock(int d, Ops &o) {
    if (d != 2) return false;
    o = make_ops<kmc::Rosenbrock, 2>();
    return true;
}
bool ops_lognormal(int d, Ops &o) {
    if (d != 1) return false;
    o = make_ops<kmc::LogNormal, 1>();
    return true;
}
}  // namespace kmc_host
