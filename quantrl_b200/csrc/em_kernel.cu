// Fused Euler–Maruyama action interval (qrl_em_step) — sm_100a: dispatcher + the generic lane-per-component kernel
// (any n <= 16, time-dependent coefficients, K = 0 reset rows).  The tiled fast path lives in em_tile_kernel.cu.
//
// One launch replaces K iterations of the reference's Python loop
//   Y[i+1] = (I + A dt) @ Y[i] + n * Ws[t_idx+i]            quantrl/envs/stochastic.py:218-242
// plus the noise draws (stochastic.py:162-170, envs/base.py:408-424), Observations = States + noise (base.py:462),
// the reward hook on all K+1 rows (base.py:471-473), rewards += Reward[-1] (base.py:1311) and the truncation
// reductions (base.py:485-499).
//
// Mapping: LPE (= next power of two >= n) lanes cooperate on one environment; lane c owns state component c, row c of
// M = I + A dt (registers) and generates its own pair of normals (Wiener increment + measurement noise).  The n
// components are exchanged with warp shuffles each sub-step, rows of States/Observations are written with fully
// coalesced 8-byte-per-lane stores (one warp = 32 consecutive doubles of the (K+1,E,n) block).
#include <stdlib.h>
#include "common.cuh"
#include <chrono>
#include "philox.cuh"

namespace qrl {

template <int LPE>
__global__ void __launch_bounds__(128) em_step_kernel(const qrl_em_args a_in, const PhiloxKeys keys) {
    const qrl_em_args a = resolve_dyn(a_in);
    // the generator's lookup tables in shared memory (philox.cuh): random 16-byte reads cost ~10 shared-memory wavefronts per
    // warp, ~16+ L1 tag look-ups when they go through the read-only path (config-3 size: 72.5 vs 65.5 us per interval)
    __shared__ __align__(16) unsigned char rng_tab_smem[8192];
    RngTabSmem rng_tab;
    rng_tab.lg_base = rng_tab.sc_base = 0u;
    if (a.rng_mode == QRL_RNG_PHILOX) rng_tab = rng_tables_to_smem(rng_tab_smem, threadIdx.x, blockDim.x);
    if (a.peer_pulled && threadIdx.x == 0) em_wait_pulled(a);     // publish-only exchange: flow control before anything is written
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = tid / LPE;
    const int c = tid % LPE;
    const int n = a.n;
    const int E = a.n_envs;
    const int K = a.K;
    const bool live = (e < E) && (c < n);
    const int el = e < E ? e : E - 1;      // clamped indices keep dead lanes' loads in bounds
    const int cl = c < n ? c : n - 1;
    const int lane_base = (threadIdx.x & 31) & ~(LPE - 1);

    const double dt = a.t_ssz;
    const double sqrt_dt = sqrt(dt);
    const bool philox = a.rng_mode == QRL_RNG_PHILOX;
    const uint32_t env_g = (uint32_t)(a.env_offset + el);
    const double std_c = (philox && a.obs_std) ? a.obs_std[(int64_t)el * a.obs_std_stride_e + cl] : 0.0;
    const double w_c = (a.reward_kind != QRL_REWARD_NONE && live) ? a.reward_w[cl] : 0.0;
    const int64_t row = (int64_t)E * n;           // doubles per time row
    const int64_t off = (int64_t)el * n + cl;     // this lane's element within a row

    double x = live ? a.x0[off] : 0.0;
    double M[LPE];
    double nzc = 0.0;
    auto load_coeffs = [&](int i) {
        const double* Ar = a.A + (int64_t)i * a.A_stride_k + (int64_t)el * a.A_stride_e + (int64_t)cl * n;
#pragma unroll
        for (int j = 0; j < LPE; ++j) {
            const double aij = (live && j < n) ? em_drift(a, Ar - (int64_t)cl * n, el, cl * n + j) : 0.0;
            // M = I + A * t_ssz, unfused like the reference (stochastic.py:219-222)
            M[j] = __dadd_rn((j == c) ? 1.0 : 0.0, __dmul_rn(aij, dt));
        }
        nzc = live ? a.nz[(int64_t)i * a.nz_stride_k + (int64_t)el * a.nz_stride_e + cl] : 0.0;
    };
    const bool coeffs_vary = (a.A_stride_k != 0) || (a.nz_stride_k != 0);
    if (K > 0) load_coeffs(0);   // K == 0 (reset row only): A / nz may be NULL

    double lo = INFINITY, hi = -INFINITY;
    double r_last = 0.0, obs = 0.0;
    // measurement noise of row 0 = time index t_idx: z1 of the previous sub-step's call (or of the reset call)
    double v = 0.0;
    if (philox && a.obs_std) {
        double z0, z1;
        const uint32_t tau = a.t_idx > 0 ? (uint32_t)(a.t_idx - 1) : 0xFFFFFFFFu;
        normal_pair(philox4x32_10(make_uint4(tau, a.episode, env_g, (uint32_t)c), keys), rng_tab, z0, z1);
        v = __dmul_rn(std_c, z1);
    }

#pragma unroll 2
    for (int i = 0; i <= K; ++i) {
        if (!philox) v = a.obs_noise ? a.obs_noise[(int64_t)i * row + off] : 0.0;
        // ---- outputs of row i ----
        obs = __dadd_rn(x, v);
        if (live) {
            if (a.states) a.states[(int64_t)i * row + off] = x;
            if (a.observations) a.observations[(int64_t)i * row + off] = obs;
            lo = fmin(lo, obs);
            hi = fmax(hi, obs);
        }
        if (a.reward_kind != QRL_REWARD_NONE) {
            const double s = (a.reward_kind == QRL_REWARD_NEG_WSQ_OBS) ? obs : x;
            double r = w_c * s * s;
#pragma unroll
            for (int o = LPE >> 1; o > 0; o >>= 1) r += shfl_xor_d(r, o);
            r_last = -r;
            if (live && c == 0 && a.reward) a.reward[(int64_t)i * E + e] = r_last;
        }
        if (i == K) break;
        // ---- noise of sub-step t_idx + i: Wiener increment now, measurement noise of the next row ----
        double w;
        if (philox) {
            double z0, z1;
            normal_pair(philox4x32_10(make_uint4((uint32_t)(a.t_idx + i), a.episode, env_g, (uint32_t)c), keys), rng_tab, z0, z1);
            w = __dmul_rn(sqrt_dt, z0);
            v = __dmul_rn(std_c, z1);
        } else {
            w = a.W[(int64_t)i * row + off];
        }
        // ---- recurrence ----
        if (coeffs_vary && i > 0) load_coeffs(i);
        // two half-sums, see em_row_update() in common.cuh (columns >= n of a padded lane group hold M = 0, x = 0)
        double lo = __dmul_rn(nzc, w), hi = 0.0;
        if (n == LPE) {
#pragma unroll
            for (int j = 0; j < LPE; ++j) {
                const double xj = shfl_d(x, lane_base + j);
                if (j < (LPE + 1) / 2) lo = fma(M[j], xj, lo); else hi = fma(M[j], xj, hi);
            }
        } else {
            const int H = (n + 1) >> 1;
#pragma unroll
            for (int j = 0; j < LPE; ++j) {
                const double xj = shfl_d(x, lane_base + j);
                const double r = fma(M[j], xj, j < H ? lo : hi);
                if (j < H) lo = r; else hi = r;
            }
        }
        x = __dadd_rn(lo, hi);
    }

    if (live) {
        if (a.x_last) a.x_last[off] = x;
        if (a.obs_last) a.obs_last[off] = obs;
        if (a.obs_last_b) a.obs_last_b[off] = obs;
        if (c == 0 && a.reward_kind != QRL_REWARD_NONE) {
            if (a.reward_last) a.reward_last[e] = r_last;
            if (a.reward_last_b) a.reward_last_b[e] = r_last;
            if (a.rewards_cum) a.rewards_cum[e] += r_last;
        }
        if (a.nan_flag && !isfinite(x)) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
    em_finish_launch(a);
}

template <int LPE>
static int launch_em(const qrl_em_args& a, cudaStream_t st) {
    const int threads = 128;
    const int64_t total = (int64_t)a.n_envs * LPE;
    const int blocks = (int)((total + threads - 1) / threads);
    em_step_kernel<LPE><<<blocks, threads, 0, st>>>(a, make_philox_keys(a.seed));
    return after_launch("em_step_kernel");
}

int em_tile_launch(const qrl_em_args& a, cudaStream_t st);  // em_tile_kernel.cu
bool em_tile3_supports(const qrl_em_args& a);                // em_tile3_kernel.cu
int em_tile3_launch(const qrl_em_args& a, cudaStream_t st);
int pack_rows_launch(const qrl_pack_args& a, int64_t dyn_bias, cudaStream_t st);  // pack_kernel.cu

}  // namespace qrl

extern "C" int qrl_em_step(const qrl_em_args* args, void* stream) {
    using namespace qrl;
    QRL_REQUIRE(args != nullptr, "qrl_em_step: args is NULL");
    const qrl_em_args& a = *args;
    QRL_REQUIRE(a.n >= 1 && a.n <= 16, "qrl_em_step: n must be in 1..16");
    QRL_REQUIRE(a.n_envs >= 0 && a.K >= 0, "qrl_em_step: n_envs and K must be non-negative");
    QRL_REQUIRE(a.rng_mode == QRL_RNG_FED || a.rng_mode == QRL_RNG_PHILOX, "qrl_em_step: bad rng_mode");
    QRL_REQUIRE(a.reward_kind >= QRL_REWARD_NONE && a.reward_kind <= QRL_REWARD_NEG_WSQ_STATE,
                "qrl_em_step: bad reward_kind");
    if (a.n_envs == 0) return 0;
    QRL_REQUIRE(a.x0 != nullptr, "qrl_em_step: x0 is NULL");
    QRL_REQUIRE(a.K == 0 || (a.A != nullptr && a.nz != nullptr), "qrl_em_step: A / nz is NULL");
    QRL_REQUIRE(a.rng_mode != QRL_RNG_FED || a.K == 0 || a.W != nullptr, "qrl_em_step: fed mode needs W");
    QRL_REQUIRE(a.reward_kind == QRL_REWARD_NONE || a.reward_w != nullptr, "qrl_em_step: reward_w is NULL");
    QRL_REQUIRE(a.t_idx >= 0 && a.env_offset >= 0, "qrl_em_step: negative t_idx / env_offset");
    QRL_REQUIRE(a.A_act == nullptr || (a.actions != nullptr && a.n_actions >= 1 && a.A_stride_k == 0),
                "qrl_em_step: the affine drift functor needs actions, n_actions >= 1 and a time-invariant A");
    QRL_REQUIRE(a.finish_counter != nullptr || (a.minmax_out == nullptr && a.dyn_advance == 0),
                "qrl_em_step: minmax_out / dyn_advance need finish_counter");
    QRL_REQUIRE(a.minmax_out == nullptr || a.obs_minmax != nullptr, "qrl_em_step: minmax_out needs obs_minmax");
    QRL_REQUIRE(a.done_flag == nullptr || a.finish_counter != nullptr, "qrl_em_step: done_flag needs finish_counter");
    QRL_REQUIRE(a.peer_flags == nullptr || (a.finish_counter != nullptr && ((a.flags & QRL_EM_PEER_PUBLISH) || a.peer_epoch != nullptr) &&
                                            a.peer_world >= 1 && a.peer_rank >= 0 && a.peer_rank < a.peer_world),
                "qrl_em_step: the peer exchange needs finish_counter, peer_epoch (barrier mode) and 0 <= peer_rank < peer_world");
    QRL_REQUIRE(!(a.flags & QRL_EM_PEER_PUBLISH) || (a.peer_flags != nullptr && a.peer_dst >= 0 && a.peer_dst < a.peer_world),
                "qrl_em_step: the publish-only exchange needs peer_flags and 0 <= peer_dst < peer_world");
    QRL_REQUIRE(a.reward_last_b == nullptr || a.reward_kind != QRL_REWARD_NONE, "qrl_em_step: reward_last_b needs a fused reward");
    if (a.packed) {
        QRL_REQUIRE(a.reward_kind != QRL_REWARD_NONE && a.rewards_cum != nullptr,
                    "qrl_em_step: packed rows need a fused reward functor and rewards_cum");
        QRL_REQUIRE(a.T_step && a.sel_idxs && a.n_sel >= 1 && a.n_actions >= 0 && (a.n_actions == 0 || a.actions),
                    "qrl_em_step: packed rows need T_step, actions and sel_idxs");
        QRL_REQUIRE(a.packed_row_pitch == 0 || a.packed_row_pitch >= a.n_sel, "qrl_em_step: packed_row_pitch < n_sel");
        QRL_REQUIRE(a.packed_stride_e >= (int64_t)(a.K + 1) * (a.packed_row_pitch > 0 ? a.packed_row_pitch : a.n_sel),
                    "qrl_em_step: packed_stride_e too small");
    }
    cudaStream_t st = (cudaStream_t)stream;
    // Tiled kernel: one CTA per SM, noise generation spread over 15 warps -> best while the batch cannot fill the chip
    // with one lane per component (measured crossover on B200: E*n ~ 32k, tools/size_sweep.py).  Beyond that the
    // generic kernel has enough warps per scheduler and no shared-memory round trip.
    // n > 4: the recurrence lane cannot hold M = I + A dt (n*n doubles) in registers next to the producers' working set
    // (spills under the 128-register cap of a 512-thread CTA) and the generic kernel wins from E = 4096 on
    // (tools/variants_em.py: n = 8, E = 4096: 72 vs 108 us), so the tiled kernel is the default only up to n = 4.
    // The third-generation tiled kernel (em_tile3_kernel.cu: n in {2, 4, 8}, several recurrence lanes per environment for
    // n = 8) has no such limit.  Measured on B200 with the cumulative-reward cache column fused (tools/t3_sweep.py, us per
    // 100 sub-steps, tiled vs generic + pack launch): n = 8: E = 8192 89 vs 122, 16384 168 vs 227, 65536 622 vs 851;
    // n = 4: E = 8192 47 vs 85, 65536 317 vs 410 (287 without cache rows) — it takes every case it supports up to 2^19 lanes.
    const bool t3 = em_tile3_supports(a);
    static const char* lanes_env = getenv("QRL_EM_TILED_MAX_LANES");   // tuning knob: crossover to the generic kernel
    const int64_t max_lanes = lanes_env ? atoll(lanes_env) : (t3 ? 524288 : 32768);
    const bool small_batch = ((int64_t)a.n_envs * a.n <= max_lanes && (a.n <= 4 || t3)) || (a.flags & QRL_EM_FORCE_TILED);
    const bool tiled = a.K >= 1 && a.n <= 8 && a.A_stride_k == 0 && a.nz_stride_k == 0 && small_batch &&
                       !(a.flags & QRL_EM_FORCE_GENERIC);
    if (tiled) return t3 ? em_tile3_launch(a, st) : em_tile_launch(a, st);
    QRL_REQUIRE(!a.packed || a.observations != nullptr, "qrl_em_step: the generic kernel packs from the observations block");
    int rc;
    if (a.n <= 1) rc = launch_em<1>(a, st);
    else if (a.n <= 2) rc = launch_em<2>(a, st);
    else if (a.n <= 4) rc = launch_em<4>(a, st);
    else if (a.n <= 8) rc = launch_em<8>(a, st);
    else rc = launch_em<16>(a, st);
    if (rc != 0 || !a.packed) return rc;
    // generic path: the cache rows are written by the pack kernel right behind the integration (same stream)
    qrl_pack_args p;
    memset(&p, 0, sizeof(p));
    p.n_envs = a.n_envs; p.dim = a.K + 1; p.n_actions = a.n_actions; p.n_obs = a.n; p.n_props = 0; p.n_sel = a.n_sel;
    p.T_step = a.T_step; p.actions = a.actions; p.observations = a.observations; p.rewards_cum = a.rewards_cum;
    p.idxs = a.sel_idxs; p.selected = a.packed; p.selected_stride_e = a.packed_stride_e;
    p.selected_row_pitch = a.packed_row_pitch;
    p.action_scale = a.A_act ? a.action_scale : nullptr;
    // the time index the pack launch must see is the one the integration launch STARTED from: with QRL_EM_HOST_TIDX the
    // host has already folded it into T_step / packed (no device-side index at all); otherwise the integration kernel's
    // last block has advanced dyn[0] by dyn_advance before this launch runs (stream order), which the bias undoes
    int64_t bias = 0;
    if (a.dyn && !(a.flags & QRL_EM_HOST_TIDX)) {
        p.dyn = a.dyn;
        if (a.finish_counter) bias = -a.dyn_advance;
    }
    return pack_rows_launch(p, bias, st);
}

// qrl_em_step + the host's wait for this step's results (header: ABI 5).  The spin reads a pinned host word the launch's
// last block stores with a system-scope release: on x86 the PCIe writes that precede it are visible once it is.
extern "C" int qrl_em_step_host(const qrl_em_args* args, void* stream, const volatile uint64_t* done_host, int32_t timeout_ms) {
    using namespace qrl;
    QRL_REQUIRE(args != nullptr, "qrl_em_step_host: args is NULL");
    QRL_REQUIRE(done_host == nullptr || args->done_flag != nullptr, "qrl_em_step_host: done_host needs args->done_flag");
    const int rc = qrl_em_step(args, stream);
    if (rc != 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (done_host == nullptr || args->n_envs == 0) {
        QRL_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    const uint64_t want = args->done_value;
    const auto t0 = std::chrono::steady_clock::now();
    const auto limit = std::chrono::milliseconds(timeout_ms > 0 ? timeout_ms : 2000);
    for (unsigned spins = 0;; ++spins) {
        if (*done_host == want) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return 0;
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if ((spins & 1023u) == 1023u && std::chrono::steady_clock::now() - t0 > limit) break;
    }
    QRL_CUDA(cudaStreamSynchronize(st));    // a failed launch surfaces here
    QRL_REQUIRE(*done_host == want, "qrl_em_step_host: the launch completed without storing done_value (done_flag and "
                                    "done_host do not name the same pinned word?)");
    return 0;
}
