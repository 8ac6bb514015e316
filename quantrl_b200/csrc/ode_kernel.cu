// Dispatcher of qrl_ode_step: validates the arguments and forwards to the per-system instantiation units
// (ode_optomech.cu, ode_kerr*.cu, ode_const*.cu; kernels in ode_impl.cuh).
#include "common.cuh"

namespace qrl {
int ode_optomech(const qrl_ode_args& a, cudaStream_t st);
int ode_kerr1(const qrl_ode_args& a, cudaStream_t st);
int ode_kerr2(const qrl_ode_args& a, cudaStream_t st);
int ode_kerr3(const qrl_ode_args& a, cudaStream_t st);
int ode_kerr4(const qrl_ode_args& a, cudaStream_t st);
int ode_const2(const qrl_ode_args& a, cudaStream_t st);
int ode_const4(const qrl_ode_args& a, cudaStream_t st);
int ode_const6(const qrl_ode_args& a, cudaStream_t st);
int ode_const8(const qrl_ode_args& a, cudaStream_t st);
}  // namespace qrl

extern "C" int qrl_ode_step(const qrl_ode_args* args, void* stream) {
    using namespace qrl;
    QRL_REQUIRE(args != nullptr, "qrl_ode_step: args is NULL");
    const qrl_ode_args& a = *args;
    QRL_REQUIRE(a.n_envs >= 0 && a.K >= 0, "qrl_ode_step: n_envs and K must be non-negative");
    QRL_REQUIRE(a.method >= QRL_ODE_RK4 && a.method <= QRL_ODE_MIDPOINT, "qrl_ode_step: unknown method");
    const bool fixed_step = a.method == QRL_ODE_RK4 || a.method == QRL_ODE_EULER || a.method == QRL_ODE_MIDPOINT;
    QRL_REQUIRE(!fixed_step || a.substeps >= 1, "qrl_ode_step: substeps must be >= 1");
    QRL_REQUIRE(fixed_step || (a.atol > 0.0 && a.rtol > 0.0), "qrl_ode_step: atol and rtol must be positive");
    QRL_REQUIRE(a.t_ssz > 0.0, "qrl_ode_step: t_ssz must be positive");
    if (a.n_envs == 0) return 0;
    QRL_REQUIRE(a.y0 != nullptr, "qrl_ode_step: y0 is NULL");
    QRL_REQUIRE(a.t_idx >= 0 && a.env_offset >= 0, "qrl_ode_step: negative t_idx / env_offset");
    QRL_REQUIRE(a.reward_kind == QRL_REWARD_NONE ||
                    (a.reward_kind >= QRL_REWARD_ENTAN_LN_OBS && a.reward_kind <= QRL_REWARD_SYNC_P_STATE && a.num_quads == 4),
                "qrl_ode_step: reward_kind must be NONE, ENTAN_LN_*, SYNC_C_* or SYNC_P_* (num_quads == 4)");
    QRL_REQUIRE(a.reward_kind < QRL_REWARD_SYNC_P_OBS || a.num_modes >= 2,
                "qrl_ode_step: SYNC_P_* needs the amplitudes of two modes (num_modes >= 2)");
    QRL_REQUIRE(a.action_scale == nullptr || (a.n_actions >= 1 && a.n_actions <= 4), "qrl_ode_step: action_scale needs 1..4 actions");
    cudaStream_t st = (cudaStream_t)stream;
    const int m = a.num_modes, q = a.num_quads;
    switch (a.system) {
        case QRL_SYS_OPTOMECH:
            QRL_REQUIRE(m == 2 && q == 4, "qrl_ode_step: optomech needs num_modes=2, num_quads=4");
            QRL_REQUIRE(a.params && a.actions && a.n_actions >= 1, "qrl_ode_step: optomech needs params and >= 1 action");
            return ode_optomech(a, st);
        case QRL_SYS_KERR_CHAIN:
            QRL_REQUIRE(q == 2 * m, "qrl_ode_step: kerr_chain needs num_quads = 2 * num_modes");
            QRL_REQUIRE(a.params && a.actions && a.n_actions >= 1, "qrl_ode_step: kerr_chain needs params and >= 1 action");
            if (m == 1) return ode_kerr1(a, st);
            if (m == 2) return ode_kerr2(a, st);
            if (m == 3) return ode_kerr3(a, st);
            if (m == 4) return ode_kerr4(a, st);
            return fail(QRL_E_UNSUPPORTED, "%s", "qrl_ode_step: kerr_chain is built for num_modes 1..4");
        case QRL_SYS_CONST_AD:
            QRL_REQUIRE(m == 0, "qrl_ode_step: const_AD has no modes (num_modes = 0)");
            QRL_REQUIRE(a.A && a.D, "qrl_ode_step: const_AD needs A and D");
            if (q == 2) return ode_const2(a, st);
            if (q == 4) return ode_const4(a, st);
            if (q == 6) return ode_const6(a, st);
            if (q == 8) return ode_const8(a, st);
            return fail(QRL_E_UNSUPPORTED, "%s", "qrl_ode_step: const_AD is built for num_quads 2, 4, 6, 8");
        default:
            return fail(QRL_E_INVALID, "%s", "qrl_ode_step: unknown system");
    }
}
