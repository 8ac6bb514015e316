// Tiled, warp-specialised Euler–Maruyama action interval — the fast path of qrl_em_step on sm_100a.
//
// Why: with one lane per state component (em_kernel.cu) config 3 (E = 4096) puts a single warp on each SM
// sub-partition and the kernel is latency bound (ncu r1a: IPC 0.25, FP64 pipe 12 %).  The noise generation
// (Philox4x32-10 + FP64 Box–Muller, > 80 % of the necessary instructions) does not depend on the state, so it is
// spread over 15 "worker" warps per CTA, while ONE warp runs the short sequential recurrence for all environments
// of the CTA.
//
// One CTA owns G <= 32 environments for the whole interval; time is cut into chunks of C sub-steps that flow through
// shared-memory tiles in a 3-stage pipeline with ONE __syncthreads per chunk and no other barrier:
//   iteration it:  workers   output(chunk it-2)           Xt[it & 1] + Vt[(it-2) % 3] -> global
//                  workers   produce noise(chunk it)      -> Wt[it & 1], Vt[it % 3]     (Philox, or fed loads)
//                  warp 0    recurrence(chunk it-1)       Wt[(it-1) & 1] -> Xt[(it-1) & 1]
// The measurement-noise tiles are triple buffered so that the output stage (which reads chunk it-2) and the producers
// (which write chunk it) never touch the same buffer inside an iteration.
//
// The kernel is issue bound (one instruction per clock per scheduler; ncu r1d: the r1 version executed 411
// instructions per pair of normals, only 165 of them in the generator), so the work distribution is chosen for
// instruction economy:
//   * production: a worker owns ONE (environment, component) column for the whole launch and walks the sub-steps
//     with a fixed stride — counter words, noise scale and shared-memory column are loop invariants, the only
//     per-pair integer work is `tau += stride`.  Two pairs per trip (ILP) and the Philox rounds of the NEXT trip are
//     issued in the same basic block as the Box–Muller polynomials of the current one, so the integer, IMAD and FP64
//     pipes overlap inside a warp (4 independent dependency chains per warp, ~15 per scheduler);
//   * output (n = 2, 4, 8): a worker owns a 16-byte piece (two components) of a tile row, which is also a 16-byte
//     piece of the global row: 128-bit shared loads of states and noise, two adds, 128-bit stores into the States /
//     Observations blocks (a warp writes 512 contiguous bytes), min/max, and the reward as ONE chain of FMAs that
//     hops over the n/2 adjacent lanes of an environment with a shuffle (same order of operations as a sequential
//     sum over the components).  No staging tile, no barrier; ~90 instructions per 32 (row, env) items (ncu r1e:
//     the one-item-per-lane version needed 200, a third of them for NaN-aware fmin/fmax);
//   * output (other n, unaligned blocks, and the cache rows): one (row, environment) item per worker.
// Replaces the same reference code as em_kernel.cu (quantrl/envs/stochastic.py:178-242, envs/base.py:408-426,462,
// 471-473,485-499) plus the packed rows of BaseSB3Env.update_data (envs/base.py:1314-1372).
#include <stdlib.h>
#include "common.cuh"
#include "philox.cuh"

namespace qrl {

constexpr int EM_TILE_THREADS = 512;
constexpr int EM_TILE_WORKERS = EM_TILE_THREADS - 32;
constexpr int EM_TILE_CMAX = 32;

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" :: "n"(EM_TILE_WORKERS) : "memory"); }

// NaN-ignoring running extrema (lo / hi are never NaN): 3 instructions each instead of the ~8 of fmin / fmax
__device__ __forceinline__ void track_minmax(double v, double& lo, double& hi) {
    lo = (v < lo) ? v : lo;
    hi = (v > hi) ? v : hi;
}

template <int N, int ILP>
__global__ void __launch_bounds__(EM_TILE_THREADS, 1)
em_tile_kernel(const qrl_em_args a_in, const PhiloxKeys keys, const int G, const int C, const int nphase) {
    // Programmatic dependent launch (QRL_EM_PDL): this grid may be scheduled while the previous kernel of the stream is
    // still draining; everything it reads (dyn, x0, cumulative rewards) is only valid after the wait.  Its own
    // dependents are released at once — they block on the same wait and, at 200 KB of shared memory per CTA, cannot
    // become resident before this grid's CTAs have exited anyway; what overlaps is the launch latency.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // `a` holds the launch arguments; the device-resident {t_idx, episode} are folded in after the wait below
    qrl_em_args a = a_in;
    extern __shared__ __align__(128) double sm[];
    const int E = a.n_envs, K = a.K;
    const int e0 = blockIdx.x * G;
    const int Gl = min(G, E - e0);             // live environments of this CTA
    const int RP = G * N;                      // doubles per tile row (same for every CTA: the column map is static)
    const int tile = C * RP;                   // doubles per tile buffer
    double* Wt = sm;                           // [2][tile] Wiener increments, already scaled by sqrt(dt)
    double* Vt = Wt + 2 * tile;                // [3][tile] measurement noise, already scaled by obs_std
    double* Xt = Vt + 3 * tile;                // [2][tile] states, written by the recurrence warp
    double* x0S = Xt + 2 * tile;               // [RP] initial state (row 0)
    double* v0S = x0S + RP;                    // [RP] measurement noise of row 0
    double* stdS = v0S + RP;                   // [RP] measurement-noise std per (env, comp)
    double* cumS = stdS + RP;                  // [G]  cumulative reward after this interval
    double* wS = cumS + G;                     // [N]  reward weights
    int* selS = reinterpret_cast<int*>(wS + N);  // [n_sel]

    const int tid = threadIdx.x;
    const bool is_rec = tid < 32;
    const int wt = tid - 32;                   // worker id (0..EM_TILE_WORKERS-1)
    const bool philox = a.rng_mode == QRL_RNG_PHILOX;
    const bool has_reward = a.reward_kind != QRL_REWARD_NONE;
    const bool reward_on_obs = a.reward_kind == QRL_REWARD_NEG_WSQ_OBS;
    const double sqrt_dt = sqrt(a.t_ssz);
    const int NC = (K + C - 1) / C;
    const int64_t row = (int64_t)E * N;
    const int n_data = 1 + a.n_actions + N + 1;
    const int n_sel = a.n_sel;
    const int pitch = a.packed ? (int)(a.packed_row_pitch > 0 ? a.packed_row_pitch : a.n_sel) : 0;

    // constants of a step sequence first (measurement-noise stds, reward weights, cache column selection: no step kernel
    // ever writes them), so that with programmatic dependent launch a CTA that became resident while the previous grid
    // is still draining has its first global round trip behind it when the wait returns
    for (int i = tid; i < Gl * N; i += EM_TILE_THREADS) {
        const int el = i / N, c = i - el * N;
        stdS[i] = (philox && a.obs_std) ? a.obs_std[(int64_t)(e0 + el) * a.obs_std_stride_e + c] : 0.0;
    }
    if (tid < N) wS[tid] = has_reward ? a.reward_w[tid] : 0.0;
    if (a.packed)
        for (int i = tid; i < n_sel; i += EM_TILE_THREADS) {
            const int s = a.sel_idxs[i];
            selS[i] = s < 0 ? s + n_data : s;
        }
    // ... and with QRL_EM_PDL (a device-resident loop of single-kernel steps: the drift base, the action-affine part and
    // the noise prefixes are constants of the sequence) the recurrence lanes pull their coefficient lines into L1
    if ((a.flags & QRL_EM_PDL) && is_rec && tid < Gl) {
        const char* Ab = reinterpret_cast<const char*>(a.A + (int64_t)(e0 + tid) * a.A_stride_e);
        for (int b = 0; b < N * N * 8; b += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(Ab + b));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nz + (int64_t)(e0 + tid) * a.nz_stride_e));
        if (a.A_act && tid < a.n_actions) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.A_act + (int64_t)tid * N * N));
    }
    // everything below reads what the previous kernel of the stream wrote (dyn, x0, cumulative rewards, actions)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the initial states do not depend on the device-resident time index: their loads go out first, so that the two
    // global round trips (x0, dyn) overlap
    {
        double x0r[(32 * N + EM_TILE_THREADS - 1) / EM_TILE_THREADS];
#pragma unroll
        for (int u = 0; u < (32 * N + EM_TILE_THREADS - 1) / EM_TILE_THREADS; ++u) {
            const int i = tid + u * EM_TILE_THREADS;
            x0r[u] = i < Gl * N ? a.x0[(int64_t)e0 * N + i] : 0.0;
        }
        a = resolve_dyn(a_in);
#pragma unroll
        for (int u = 0; u < (32 * N + EM_TILE_THREADS - 1) / EM_TILE_THREADS; ++u) {
            const int i = tid + u * EM_TILE_THREADS;
            if (i < Gl * N) x0S[i] = x0r[u];
        }
    }
    if (a.peer_pulled && tid == 0) em_wait_pulled(a);     // publish-only exchange: flow control before anything is written
    __syncthreads();
    // does any selected cache column have to be written row by row?  (the cumulative-reward column is only known
    // after the last sub-step and is written for all K+1 rows at the end)
    bool pack_rows = false;
    if (a.packed)
        for (int j = 0; j < n_sel; ++j) pack_rows |= (selS[j] != n_data - 1);
    // 128-bit accesses need even N (tile rows and every env's N-run are then 16-byte aligned) and aligned bases
    const bool vec_g = (N % 2 == 0) && ((((uintptr_t)a.states) | ((uintptr_t)a.observations)) % 16 == 0);
    const bool vec_l = (N % 2 == 0) && ((((uintptr_t)a.x_last) | ((uintptr_t)a.obs_last)) % 16 == 0);
    // piece pass: n/2 adjacent lanes per environment
    constexpr bool PIECE_N = (N == 2 || N == 4 || N == 8);
    constexpr int L = PIECE_N ? N / 2 : 1;
    // (the piece pass walks the blocks with 32-bit element offsets)
    const bool piece_out = PIECE_N && vec_g && !(a.flags & QRL_EM_ITEM_OUTPUT) && (int64_t)(K + 1) * row < (1ll << 31);

    double lo = INFINITY, hi = -INFINITY;

    // ---- output stage, item pass: one (row, env) item per worker and trip ------------------------------------------
    // rows r_first .. r_first+ns-1 whose states / noises sit in Xb / Vb (row-major [slot][env][comp], row pitch RP).
    // the item does everything but the cache rows: adds, States / Observations rows, min/max, reward, last-row outputs.
    const int o_dsl = EM_TILE_WORKERS / Gl, o_del = EM_TILE_WORKERS - o_dsl * Gl;
    const int o_sl0 = wt / Gl, o_el0 = wt - o_sl0 * Gl;
    auto item_pass = [&](const double* Xb, const double* Vb, int r_first, int ns) {
        int sl = o_sl0, el = o_el0;
        while (sl < ns) {
            const int r = r_first + sl, e = e0 + el;
            const int off = sl * RP + el * N;
            double xv[N], ov[N];
            if constexpr (N % 2 == 0) {
                const double2* xp = reinterpret_cast<const double2*>(Xb + off);
                const double2* vp = reinterpret_cast<const double2*>(Vb + off);
#pragma unroll
                for (int c = 0; c < N / 2; ++c) {
                    const double2 x2 = xp[c], v2 = vp[c];
                    xv[2 * c] = x2.x; xv[2 * c + 1] = x2.y;
                    ov[2 * c] = __dadd_rn(x2.x, v2.x); ov[2 * c + 1] = __dadd_rn(x2.y, v2.y);
                }
            } else {
#pragma unroll
                for (int c = 0; c < N; ++c) { xv[c] = Xb[off + c]; ov[c] = __dadd_rn(xv[c], Vb[off + c]); }
            }
            {
                const int64_t g = (int64_t)r * row + (int64_t)e * N;
                if (vec_g) {
                    if constexpr (N % 2 == 0) {
                        if (a.states) {
                            double2* sp = reinterpret_cast<double2*>(a.states + g);
#pragma unroll
                            for (int c = 0; c < N / 2; ++c) sp[c] = make_double2(xv[2 * c], xv[2 * c + 1]);
                        }
                        if (a.observations) {
                            double2* op = reinterpret_cast<double2*>(a.observations + g);
#pragma unroll
                            for (int c = 0; c < N / 2; ++c) op[c] = make_double2(ov[2 * c], ov[2 * c + 1]);
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        if (a.states) a.states[g + c] = xv[c];
                        if (a.observations) a.observations[g + c] = ov[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < N; ++c) track_minmax(ov[c], lo, hi);
                double rv = 0.0;
                if (has_reward) {
                    double acc = 0.0;
                    if (reward_on_obs) {
#pragma unroll
                        for (int c = 0; c < N; ++c) acc = fma(wS[c] * ov[c], ov[c], acc);
                    } else {
#pragma unroll
                        for (int c = 0; c < N; ++c) acc = fma(wS[c] * xv[c], xv[c], acc);
                    }
                    rv = -acc;
                    if (a.reward) a.reward[(int64_t)r * E + e] = rv;
                }
                if (r == K) {
                    if (has_reward) {
                        if (a.reward_last) a.reward_last[e] = rv;
                        if (a.reward_last_b) a.reward_last_b[e] = rv;
                        double cum = rv;
                        if (a.rewards_cum) { cum += a.rewards_cum[e]; a.rewards_cum[e] = cum; }
                        cumS[el] = cum;
                    }
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        if (a.x_last) a.x_last[(int64_t)e * N + c] = xv[c];
                        if (a.obs_last) a.obs_last[(int64_t)e * N + c] = ov[c];
                        if (a.obs_last_b) a.obs_last_b[(int64_t)e * N + c] = ov[c];
                    }
                }
            }
            sl += o_dsl;
            el += o_del;
            if (el >= Gl) { el -= Gl; ++sl; }
        }
    };

    // ---- output stage, piece pass: one 16-byte piece (two components) of a row per worker and trip ----------------
    const int PR = Gl * L;                     // pieces per live row
    const int q_dsl = EM_TILE_WORKERS / PR, q_dpc = EM_TILE_WORKERS - q_dsl * PR;
    const int q_sl0 = wt / PR, q_pc0 = wt - q_sl0 * PR;
    const int q_h = wt & (L - 1);              // which piece of its environment this lane holds (pieces of an env sit
                                               // in adjacent lanes: EM_TILE_WORKERS and PR are multiples of L)
    double q_w0 = 0.0, q_w1 = 0.0;             // reward weights of this lane's two components
    if constexpr (PIECE_N) { q_w0 = wS[2 * q_h]; q_w1 = wS[2 * q_h + 1]; }
    auto piece_pass = [&](const double* Xb, const double* Vb, int r_first, int ns) {
        if constexpr (PIECE_N) {
            int sl = q_sl0, pc = q_pc0;
            // running element offset of this lane's piece in the (K+1, E, N) blocks: one 32-bit add per trip instead
            // of a 64-bit multiply-add chain per tensor (the pass is issue bound: ncu r1k, 136 instructions per trip)
            const uint32_t row32 = (uint32_t)row;
            uint32_t g = (uint32_t)(r_first + sl) * row32 + (uint32_t)(e0 * N + 2 * pc);
            const uint32_t g_step = (uint32_t)q_dsl * row32 + (uint32_t)(2 * q_dpc), g_wrap = row32 - (uint32_t)(2 * PR);
            // warp-uniform trip count (the shuffles below need all 32 lanes): lane 0 holds the smallest piece index
            int sl_lead = __shfl_sync(0xffffffffu, sl, 0);
            while (sl_lead < ns) {
                const bool live = sl < ns;
                const int off = live ? sl * RP + 2 * pc : 0;
                const double2 x2 = *reinterpret_cast<const double2*>(Xb + off);
                const double2 v2 = *reinterpret_cast<const double2*>(Vb + off);
                const double2 o2 = make_double2(__dadd_rn(x2.x, v2.x), __dadd_rn(x2.y, v2.y));
                const int r = r_first + sl;
                if (live) {
                    if (a.states) *reinterpret_cast<double2*>(a.states + g) = x2;
                    if (a.observations) *reinterpret_cast<double2*>(a.observations + g) = o2;
                    track_minmax(o2.x, lo, hi);
                    track_minmax(o2.y, lo, hi);
                }
                double rv = 0.0;
                if (has_reward) {
                    // -sum_c w_c s_c^2 as one chain of FMAs in component order, hopping over the L lanes of the env
                    const double s0 = reward_on_obs ? o2.x : x2.x, s1 = reward_on_obs ? o2.y : x2.y;
                    const double ws0 = q_w0 * s0, ws1 = q_w1 * s1;
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        const double t = fma(ws1, s1, fma(ws0, s0, acc));
                        if (j == L - 1) { rv = -t; break; }
                        const double up = __shfl_up_sync(0xffffffffu, t, 1);
                        if (q_h == j + 1) acc = up;
                    }
                }
                if (live) {
                    const int el = pc / L, e = e0 + el;
                    if (has_reward && q_h == L - 1 && a.reward) a.reward[(uint32_t)r * (uint32_t)E + (uint32_t)e] = rv;
                    if (r == K) {
                        if (has_reward && q_h == L - 1) {
                            if (a.reward_last) a.reward_last[e] = rv;
                            if (a.reward_last_b) a.reward_last_b[e] = rv;
                            double cum = rv;
                            if (a.rewards_cum) { cum += a.rewards_cum[e]; a.rewards_cum[e] = cum; }
                            cumS[el] = cum;
                        }
                        if (a.obs_last_b) { a.obs_last_b[(int64_t)e0 * N + 2 * pc] = o2.x; a.obs_last_b[(int64_t)e0 * N + 2 * pc + 1] = o2.y; }
                        // last rows: one 16-byte store per piece (obs_last may be the learner's peer-mapped buffer:
                        // 8-byte stores would double the NVLink write packets)
                        const int64_t gl = (int64_t)e0 * N + 2 * pc;
                        if (vec_l) {
                            if (a.x_last) *reinterpret_cast<double2*>(a.x_last + gl) = x2;
                            if (a.obs_last) *reinterpret_cast<double2*>(a.obs_last + gl) = o2;
                        } else {
                            if (a.x_last) { a.x_last[gl] = x2.x; a.x_last[gl + 1] = x2.y; }
                            if (a.obs_last) { a.obs_last[gl] = o2.x; a.obs_last[gl + 1] = o2.y; }
                        }
                    }
                }
                sl += q_dsl;
                pc += q_dpc;
                g += g_step;
                if (pc >= PR) { pc -= PR; ++sl; g += g_wrap; }
                sl_lead = __shfl_sync(0xffffffffu, sl, 0);
            }
        }
    };
    // ---- output stage, cache rows: packed[e, r, j] = column sel[j] of [T_step[r], actions[e,:], Obs[r,e,:], cum. reward]
    // (envs/base.py:1314-1372).  The rows r_first .. r_first+ns-1 of ONE environment are a contiguous run of ns * pitch
    // doubles of the env-major cache, so a warp walks such a run 32 elements at a time (lane-contiguous 8-byte stores =
    // 256-byte segments) and gathers each element from the tiles; (env, trip) pairs are dealt round-robin to the 15
    // worker warps.  The cumulative-reward column is skipped here and written for all rows after the last sub-step.
    const uint32_t pk_magic = pitch > 0 ? 0xFFFFFFFFu / (uint32_t)pitch + 1u : 0u;   // idx / pitch for idx < 2^16
    auto pack_pass = [&](const double* Xb, const double* Vb, int r_first, int ns) {
        const int run = ns * pitch;                       // doubles per environment
        const int trips = (run + 31) >> 5;
        const int lane = wt & 31;
        int t = wt >> 5;                                  // worker warp 0..14
        int el = t / trips, trip = t - el * trips;
        const int d_el = (EM_TILE_WORKERS / 32) / trips, d_trip = (EM_TILE_WORKERS / 32) - d_el * trips;
        while (el < Gl) {
            const int idx = (trip << 5) + lane;
            if (idx < run) {
                const int sl = pitch == 1 ? idx : (int)__umulhi((uint32_t)idx, pk_magic), j = idx - sl * pitch;
                const int col = j < n_sel ? selS[j] : n_data - 1;
                if (col != n_data - 1) {
                    const int r = r_first + sl, e = e0 + el;
                    double val;
                    if (col == 0) val = a.T_step[r];
                    else if (col <= a.n_actions) val = em_action(a, e, col - 1);
                    else {
                        const int off = sl * RP + el * N + (col - 1 - a.n_actions);
                        val = __dadd_rn(Xb[off], Vb[off]);
                    }
                    a.packed[(int64_t)e * a.packed_stride_e + (int64_t)r_first * pitch + idx] = val;
                }
            }
            el += d_el;
            trip += d_trip;
            if (trip >= trips) { trip -= trips; ++el; }
        }
    };
    auto output_rows = [&](const double* Xb, const double* Vb, int r_first, int ns) {
        if (piece_out) piece_pass(Xb, Vb, r_first, ns);
        else item_pass(Xb, Vb, r_first, ns);
        if (pack_rows) pack_pass(Xb, Vb, r_first, ns);
    };

    // ---- recurrence warp: coefficients and state in registers ----------------------------------------------------
    double M[N][N], nz[N], x[N];
    const int el_r = tid;                      // env handled by this recurrence lane
    const bool rec_live = is_rec && el_r < Gl;
    // ---- producer: the (env, comp) column of this worker and its position in the sub-step walk ------------------
    const int p_phase = wt / RP, p_col = wt - p_phase * RP;
    const int p_el = p_col / N, p_c = p_col - p_el * N;
    const bool p_live = !is_rec && p_phase < nphase && p_el < Gl;
    const uint32_t p_env = (uint32_t)(a.env_offset + e0 + p_el);
    const uint32_t tau_base = (uint32_t)a.t_idx;
    double p_sd = 0.0;
    int pj = p_phase;                          // next sub-step (within the interval) this worker generates
    uint4 nxt[ILP];

    if (rec_live) {
        const int e = e0 + el_r;
        const double* Ar = a.A + (int64_t)e * a.A_stride_e;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) M[r][c] = Ar[r * N + c];
        if (a.A_act) {
            // action-affine drift functor, same accumulation order as em_drift(): A += u_k * A_k for k = 0, 1, ...
#pragma unroll 1
            for (int k = 0; k < a.n_actions; ++k) {
                const double u = a.actions[(int64_t)e * a.n_actions + k] * (a.action_scale ? a.action_scale[k] : 1.0);
                const double* Ak = a.A_act + (int64_t)k * N * N;
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) M[r][c] = fma(u, Ak[r * N + c], M[r][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < N; ++r) {
#pragma unroll
            for (int c = 0; c < N; ++c) M[r][c] = __dadd_rn((r == c) ? 1.0 : 0.0, __dmul_rn(M[r][c], a.t_ssz));
            nz[r] = a.nz[(int64_t)e * a.nz_stride_e + r];
            x[r] = x0S[el_r * N + r];
        }
    } else if (!is_rec) {
        // ---- workers: row 0, Observations[0] = States[0] + noise[t_idx]  (envs/base.py:462) ----
        for (int i = wt; i < Gl * N; i += EM_TILE_WORKERS) {
            const int el = i / N, c = i - el * N;
            double v = 0.0;
            if (philox) {
                if (a.obs_std) {
                    double z0, z1;
                    const uint32_t tau = a.t_idx > 0 ? (uint32_t)(a.t_idx - 1) : 0xFFFFFFFFu;
                    philox_normal2(keys, tau, a.episode, (uint32_t)(a.env_offset + e0 + el), (uint32_t)c, z0, z1);
                    v = __dmul_rn(stdS[i], z1);
                }
            } else if (a.obs_noise) {
                v = a.obs_noise[(int64_t)e0 * N + i];
            }
            v0S[i] = v;
        }
        if (philox && p_live) {
            p_sd = stdS[p_col];
#pragma unroll
            for (int u = 0; u < ILP; ++u)
                nxt[u] = philox4x32_10(make_uint4(tau_base + (uint32_t)(pj + u * nphase), a.episode, p_env, (uint32_t)p_c), keys);
        }
        worker_barrier();
        output_rows(x0S, v0S, 0, 1);
    }

    const bool dbg_no_prod = a.flags & 2, dbg_no_rec = a.flags & 4, dbg_no_out = a.flags & 8;  // diagnostics only

    for (int it = 0; it < NC + 2; ++it) {
        if (!is_rec) {
            // ---------------- output chunk `it - 2` ----------------
            if (it >= 2 && !dbg_no_out) {
                const int co = it - 2;
                output_rows(Xt + (co & 1) * tile, Vt + (co % 3) * tile, co * C + 1, min(C, K - co * C));
            }
            // ---------------- produce noise of chunk `it` ----------------
            if (it < NC && !dbg_no_prod) {
                const int j0 = it * C;
                const int j_end = min(j0 + C, K);
                double* Wb = Wt + (it & 1) * tile;
                double* Vb = Vt + (it % 3) * tile;
                if (philox) {
                    if (p_live) {
                        // C is a multiple of ILP * nphase (plan_tiles), so a trip never straddles a chunk boundary and
                        // the prefetched counters of the next trip stay valid across chunks.
                        double* wp = Wb + (pj - j0) * RP + p_col;
                        double* vp = Vb + (pj - j0) * RP + p_col;
                        const int step = ILP * nphase;
                        for (; pj < j_end; pj += step, wp += step * RP, vp += step * RP) {
                            uint4 cur[ILP];
#pragma unroll
                            for (int u = 0; u < ILP; ++u) {
                                cur[u] = nxt[u];
                                // next trip of this chain: unconditional (a pair past the end of the interval is
                                // computed and dropped), so the whole trip is one basic block
                                nxt[u] = philox4x32_10(make_uint4(tau_base + (uint32_t)(pj + step + u * nphase), a.episode, p_env, (uint32_t)p_c), keys);
                            }
                            double z0[ILP], z1[ILP];
#pragma unroll
                            for (int u = 0; u < ILP; ++u) normal_pair(cur[u], z0[u], z1[u]);
#pragma unroll
                            for (int u = 0; u < ILP; ++u) {
                                if (u == 0 || pj + u * nphase < j_end) {
                                    wp[u * nphase * RP] = __dmul_rn(sqrt_dt, z0[u]);
                                    vp[u * nphase * RP] = __dmul_rn(p_sd, z1[u]);
                                }
                            }
                        }
                    }
                } else {
                    // fed noise: 4 independent loads in flight per lane (the loop is global-latency bound otherwise)
                    const int ns = j_end - j0;
                    const int RL = Gl * N;
                    const int total = ns * RL;
                    const int64_t g0 = (int64_t)j0 * row + (int64_t)e0 * N;
                    for (int base = wt; base < total; base += 4 * EM_TILE_WORKERS) {
                        double wv[4], vv[4];
                        int sidx[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int idx = base + u * EM_TILE_WORKERS;
                            wv[u] = vv[u] = 0.0;
                            sidx[u] = -1;
                            if (idx < total) {
                                const int s2 = idx / RL, rem = idx - s2 * RL;
                                const int64_t g = g0 + (int64_t)s2 * row + rem;
                                sidx[u] = s2 * RP + rem;
                                wv[u] = a.W[g];
                                if (a.obs_noise) vv[u] = a.obs_noise[g + row];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (sidx[u] >= 0) { Wb[sidx[u]] = wv[u]; Vb[sidx[u]] = vv[u]; }
                    }
                }
            }
        } else if (it >= 1 && it <= NC && rec_live && !dbg_no_rec) {
            // ---------------- recurrence of chunk `it - 1` (the only sequential part) ----------------
            const int cr = it - 1;
            const int ns = min(C, K - cr * C);
            const double* Wb = Wt + (cr & 1) * tile + el_r * N;
            double* Xb = Xt + (cr & 1) * tile + el_r * N;
            double w[N];
            auto load_w = [&](int slot) {
                if constexpr (N % 2 == 0) {   // RP even and 16-byte aligned tiles: 128-bit shared-memory accesses
                    const double2* wp = reinterpret_cast<const double2*>(Wb + slot * RP);
#pragma unroll
                    for (int c = 0; c < N / 2; ++c) { const double2 t = wp[c]; w[2 * c] = t.x; w[2 * c + 1] = t.y; }
                } else {
#pragma unroll
                    for (int c = 0; c < N; ++c) w[c] = Wb[slot * RP + c];
                }
            };
            load_w(0);
#pragma unroll 2
            for (int slot = 0; slot < ns; ++slot) {
                double xn[N];
#pragma unroll
                for (int r = 0; r < N; ++r) xn[r] = em_row_update<N>(M[r], x, nz[r], w[r]);
                if (slot + 1 < ns) load_w(slot + 1);   // the next increments arrive while this sub-step's chain drains
#pragma unroll
                for (int c = 0; c < N; ++c) x[c] = xn[c];
                if constexpr (N % 2 == 0) {
                    double2* xp = reinterpret_cast<double2*>(Xb + slot * RP);
#pragma unroll
                    for (int c = 0; c < N / 2; ++c) xp[c] = make_double2(xn[2 * c], xn[2 * c + 1]);
                } else {
#pragma unroll
                    for (int c = 0; c < N; ++c) Xb[slot * RP + c] = xn[c];
                }
            }
        }
        __syncthreads();
    }

    if (rec_live && a.nan_flag) {
        bool bad = false;
#pragma unroll
        for (int c = 0; c < N; ++c) bad |= !isfinite(x[c]);
        if (bad) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
    }
    // cumulative-reward column of every cache row of this interval (envs/base.py:1311, 1342-1349): one warp per env,
    // lanes along the rows of that env's contiguous run
    if (a.packed) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int j = 0; j < n_sel; ++j) {
            if (selS[j] != n_data - 1) continue;
            for (int el = warp; el < Gl; el += EM_TILE_THREADS / 32) {
                double* dst = a.packed + (int64_t)(e0 + el) * a.packed_stride_e + j;
                const double cum = cumS[el];
                for (int r = lane; r <= K; r += 32) dst[(int64_t)r * pitch] = cum;
            }
        }
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
    em_finish_launch(a);
}

struct TilePlan { int G, C, nphase, ilp, grid; size_t smem; };

static TilePlan plan_tiles(int E, int N, int K, int n_sel, int sms) {
    TilePlan p;
    p.G = E <= 32 * sms ? (E + sms - 1) / sms : 32;
    if (p.G & 1) ++p.G;                        // even G keeps every tile row 16-byte aligned whatever N
    if (p.G > 32) p.G = 32;
    if (p.G < 2) p.G = 2;
    p.grid = (E + p.G - 1) / p.G;
    const int RP = p.G * N;
    const size_t budget = 210 * 1024;
    const size_t fixed = (size_t)(3 * RP + p.G + N) * sizeof(double) + (size_t)n_sel * sizeof(int) + 256;
    const size_t per_c = (size_t)(7 * RP) * sizeof(double);
    int cmax = (int)((budget - fixed) / per_c);
    if (cmax > K) cmax = K;
    if (cmax > EM_TILE_CMAX) cmax = EM_TILE_CMAX;
    if (cmax < 1) cmax = 1;
    // producers: nphase workers share one (env, comp) column and interleave its sub-steps; the chunk length is a
    // multiple of ILP * nphase so that every worker generates the same number of pairs per chunk
    p.nphase = EM_TILE_WORKERS / RP;
    if (p.nphase < 1) p.nphase = 1;
    if (p.nphase > cmax) p.nphase = cmax;
    // Pairs per trip: measured on B200 at config 3, two pairs per trip (ILP 2) are no faster than one (25.6 vs 28.8 us:
    // the loop is bound by issue slots and pipe contention between warps, not by dependency latency, and the second
    // pair costs registers), so 1 is the default; QRL_EM_TILE_ILP=2 keeps the variant reachable for tuning.
    p.ilp = 1;
    static const char* ilp_env = getenv("QRL_EM_TILE_ILP");
    if (ilp_env && atoi(ilp_env) >= 2 && cmax >= 2 * p.nphase) p.ilp = 2;
    p.C = cmax / (p.ilp * p.nphase) * (p.ilp * p.nphase);
    static const char* c_env = getenv("QRL_EM_TILE_C");
    if (c_env) {   // tuning knob: a smaller multiple of ilp * nphase
        const int want = atoi(c_env) / (p.ilp * p.nphase) * (p.ilp * p.nphase);
        if (want >= p.ilp * p.nphase && want <= p.C) p.C = want;
    }
    p.smem = per_c * p.C + fixed;
    return p;
}

template <int N>
static int launch_tile(const qrl_em_args& a, cudaStream_t st, int sms) {
    const TilePlan p = plan_tiles(a.n_envs, N, a.K, a.packed ? a.n_sel : 0, sms);
    static OncePerDevice smem_attr;
    if (smem_attr.first()) {
        QRL_CUDA(cudaFuncSetAttribute(em_tile_kernel<N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        QRL_CUDA(cudaFuncSetAttribute(em_tile_kernel<N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    const PhiloxKeys keys = make_philox_keys(a.seed);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(EM_TILE_THREADS);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (a.flags & QRL_EM_PDL) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.ilp == 2) QRL_CUDA(cudaLaunchKernelEx(&cfg, em_tile_kernel<N, 2>, a, keys, p.G, p.C, p.nphase));
    else QRL_CUDA(cudaLaunchKernelEx(&cfg, em_tile_kernel<N, 1>, a, keys, p.G, p.C, p.nphase));
    return after_launch("em_tile_kernel");
}

// entry used by the qrl_em_step dispatcher (em_kernel.cu); returns QRL_E_UNSUPPORTED when the case is not tiled
int em_tile_launch(const qrl_em_args& a, cudaStream_t st) {
    const int sms = device_sm_count();
    switch (a.n) {
        case 1: return launch_tile<1>(a, st, sms);
        case 2: return launch_tile<2>(a, st, sms);
        case 3: return launch_tile<3>(a, st, sms);
        case 4: return launch_tile<4>(a, st, sms);
        case 5: return launch_tile<5>(a, st, sms);
        case 6: return launch_tile<6>(a, st, sms);
        case 7: return launch_tile<7>(a, st, sms);
        case 8: return launch_tile<8>(a, st, sms);
        default: return QRL_E_UNSUPPORTED;
    }
}

}  // namespace qrl
