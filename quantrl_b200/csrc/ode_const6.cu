// instantiation unit of the deterministic-interval kernels (see ode_impl.cuh)
#include "ode_impl.cuh"
namespace qrl {
QRL_ODE_INSTANTIATE(ode_const6, SysConstAD<6>)
}  // namespace qrl
