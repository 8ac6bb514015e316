// placeholder until the RK4/DOPRI5 kernels land (next commit)
#include "common.cuh"
extern "C" int qrl_ode_step(const qrl_ode_args* args, void* stream) {
    (void)args; (void)stream;
    return qrl::fail(QRL_E_UNSUPPORTED, "%s", "qrl_ode_step: not built yet");
}
