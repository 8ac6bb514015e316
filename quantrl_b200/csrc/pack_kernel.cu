// Packed trajectory-cache rows (qrl_pack_rows) — sm_100a.
//
// Replaces the repeat/transpose/concatenate chain of BaseSB3Env.update_data (quantrl/envs/base.py:1314-1356) and the
// host-side column gather `data[:, t0:t1, :] = _data[:, :, data_idxs]` (base.py:1370-1372).
// The source blocks are time-major ((dim,E,n_obs)), the cache is env-major ((E,dim,n_data)): each CTA stages a
// [G envs x TT rows] tile of observations/properties in shared memory (coalesced reads along E*n_obs), then writes
// each env's TT*n_data doubles as one contiguous, coalesced run.
#include "common.cuh"

namespace qrl {

constexpr int PACK_G = 8;    // envs per tile
constexpr int PACK_TT = 32;  // time rows per tile

// dyn_bias: added to the device-resident time index — the pack launch that follows a generic Euler–Maruyama launch whose
// last block has ALREADY advanced dyn[0] by K (qrl_em_args.dyn_advance) passes -K and so sees the interval's first row
__global__ void __launch_bounds__(256) pack_rows_kernel(const qrl_pack_args a_in, const int64_t dyn_bias) {
    qrl_pack_args a = a_in;
    if (a_in.dyn) {   // CUDA-graph replay: episode bases + device-resident time index
        const int64_t t = a_in.dyn[0] + dyn_bias;
        a.T_step += t;
        if (a.packed) a.packed += t * (2 + a.n_actions + a.n_obs + a.n_props);
        if (a.selected) a.selected += t * (a.selected_row_pitch > 0 ? a.selected_row_pitch : (int64_t)a.n_sel);
    }
    extern __shared__ double tile[];  // [PACK_TT][PACK_G][n_src], n_src = n_obs + n_props
    const int E = a.n_envs, dim = a.dim, na = a.n_actions, no = a.n_obs, np = a.n_props;
    const int n_src = no + np;
    const int n_data = 1 + na + no + np + 1;
    const int e0 = blockIdx.x * PACK_G, t0 = blockIdx.y * PACK_TT;
    const int ge = min(PACK_G, E - e0), gt = min(PACK_TT, dim - t0);

    // stage observations: for each row t the ge*no doubles starting at Obs[t, e0, 0] are contiguous
    for (int idx = threadIdx.x; idx < gt * ge * no; idx += blockDim.x) {
        const int t = idx / (ge * no), r = idx % (ge * no);
        const int el = r / no, c = r % no;
        tile[(t * PACK_G + el) * n_src + c] = a.observations[((int64_t)(t0 + t) * E + e0) * no + r];
    }
    if (np > 0 && a.properties) {
        for (int idx = threadIdx.x; idx < gt * ge * np; idx += blockDim.x) {
            const int t = idx / (ge * np), r = idx % (ge * np);
            const int el = r / np, c = r % np;
            tile[(t * PACK_G + el) * n_src + no + c] = a.properties[((int64_t)(t0 + t) * E + e0) * np + r];
        }
    }
    __syncthreads();

    auto column = [&](int el, int t, int col) -> double {
        if (col == 0) return a.T_step[t0 + t];
        if (col <= na) return a.actions[(int64_t)(e0 + el) * na + (col - 1)] * (a.action_scale ? a.action_scale[col - 1] : 1.0);
        if (col <= na + n_src) return tile[(t * PACK_G + el) * n_src + (col - 1 - na)];
        return a.rewards_cum[e0 + el];
    };

    if (a.packed) {
        const int per_env = gt * n_data;
        for (int idx = threadIdx.x; idx < ge * per_env; idx += blockDim.x) {
            const int el = idx / per_env, r = idx % per_env;
            const int t = r / n_data, col = r % n_data;
            a.packed[(int64_t)(e0 + el) * a.packed_stride_e + (int64_t)t0 * n_data + r] = column(el, t, col);
        }
    }
    if (a.selected && a.n_sel > 0) {
        const int ns = a.n_sel;
        const int64_t pitch = a.selected_row_pitch > 0 ? a.selected_row_pitch : ns;
        const int per_env = gt * ns;
        for (int idx = threadIdx.x; idx < ge * per_env; idx += blockDim.x) {
            const int el = idx / per_env, r = idx % per_env;
            const int t = r / ns, j = r % ns;
            int col = a.idxs[j];
            if (col < 0) col += n_data;
            a.selected[(int64_t)(e0 + el) * a.selected_stride_e + (int64_t)(t0 + t) * pitch + j] = column(el, t, col);
        }
    }
}

}  // namespace qrl

namespace qrl {
int pack_rows_launch(const qrl_pack_args& a, int64_t dyn_bias, cudaStream_t st) {
    QRL_REQUIRE(a.n_envs >= 0 && a.dim >= 0 && a.n_actions >= 0 && a.n_obs >= 1 && a.n_props >= 0 && a.n_sel >= 0,
                "qrl_pack_rows: negative size");
    if (a.n_envs == 0 || a.dim == 0) return 0;
    QRL_REQUIRE(a.T_step && a.observations && a.rewards_cum, "qrl_pack_rows: NULL input");
    QRL_REQUIRE(a.n_actions == 0 || a.actions, "qrl_pack_rows: actions is NULL");
    QRL_REQUIRE(a.n_props == 0 || a.properties, "qrl_pack_rows: properties is NULL");
    QRL_REQUIRE(a.n_sel == 0 || !a.selected || a.idxs, "qrl_pack_rows: idxs is NULL");
    QRL_REQUIRE(!a.packed || a.packed_stride_e >= (int64_t)a.dim * (2 + a.n_actions + a.n_obs + a.n_props),
                "qrl_pack_rows: packed_stride_e too small");
    QRL_REQUIRE(a.selected_row_pitch == 0 || a.selected_row_pitch >= a.n_sel, "qrl_pack_rows: selected_row_pitch < n_sel");
    QRL_REQUIRE(!a.selected || a.selected_stride_e >= (int64_t)a.dim * (a.selected_row_pitch > 0 ? a.selected_row_pitch : a.n_sel),
                "qrl_pack_rows: selected_stride_e too small");
    const size_t smem = (size_t)PACK_TT * PACK_G * (a.n_obs + a.n_props) * sizeof(double);
    QRL_REQUIRE(smem <= 200 * 1024, "qrl_pack_rows: n_obs + n_props too large for the staging tile");
    static OncePerDevice attr;
    if (smem > 48 * 1024 && attr.first())
        QRL_CUDA(cudaFuncSetAttribute(pack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((a.n_envs + PACK_G - 1) / PACK_G, (a.dim + PACK_TT - 1) / PACK_TT);
    pack_rows_kernel<<<grid, 256, smem, st>>>(a, dyn_bias);
    return after_launch("pack_rows_kernel");
}
}  // namespace qrl

extern "C" int qrl_pack_rows(const qrl_pack_args* args, void* stream) {
    using namespace qrl;
    QRL_REQUIRE(args != nullptr, "qrl_pack_rows: args is NULL");
    return pack_rows_launch(*args, 0, (cudaStream_t)stream);
}
