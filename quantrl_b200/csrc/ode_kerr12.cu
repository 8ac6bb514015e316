// instantiation unit of the deterministic-interval kernels (see ode_impl.cuh)
#include "ode_impl.cuh"
namespace qrl {
QRL_ODE_INSTANTIATE(ode_kerr1, SysKerrChain<1>)
QRL_ODE_INSTANTIATE(ode_kerr2, SysKerrChain<2>)
}  // namespace qrl
