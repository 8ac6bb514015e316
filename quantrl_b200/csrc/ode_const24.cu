// instantiation unit of the deterministic-interval kernels (see ode_impl.cuh)
#include "ode_impl.cuh"
namespace qrl {
QRL_ODE_INSTANTIATE(ode_const2, SysConstAD<2>)
QRL_ODE_INSTANTIATE(ode_const4, SysConstAD<4>)
}  // namespace qrl
