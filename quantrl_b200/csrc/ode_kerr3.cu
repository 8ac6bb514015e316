// instantiation unit of the deterministic-interval kernels (see ode_impl.cuh)
#include "ode_impl.cuh"
namespace qrl {
QRL_ODE_INSTANTIATE(ode_kerr3, SysKerrChain<3>)
}  // namespace qrl
