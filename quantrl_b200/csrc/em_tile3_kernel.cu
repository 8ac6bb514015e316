// Tiled Euler–Maruyama action interval, third generation — the fast path of qrl_em_step for n in {2, 4, 8} on sm_100a.
//
// Same decomposition as em_tile_kernel.cu (noise generation spread over many "worker" warps, the short sequential
// recurrence on a few lanes, rows leaving as 16-byte pieces), rebuilt around what ncu showed on the second generation
// (profiles/r01q_ncu_summary.md: issue slots 48 % busy, 1.8 warps per issue waiting at the per-chunk __syncthreads, one
// 128-register allocation shared by all roles, n = 8 spilling 568 bytes in the recurrence lane):
//   * roles synchronise through shared-memory mbarriers per ring stage instead of one CTA barrier per chunk:
//       full_w[s]  workers -> recurrence   "the noise of the chunk in stage s is complete"
//       full_x[s]  recurrence -> workers   "the states of that chunk are complete"
//       empty[s]   workers -> workers      "the rows of that chunk have left, the stage may be refilled"
//     so a warp only ever waits for the data it needs, and the recurrence drifts freely behind the producers;
//   * the recurrence overwrites the Wiener increments IN PLACE with the new states (each element is read and written by
//     the same lane), so a ring stage is two tiles (W/X, V) instead of 7/3 — room for deeper rings of smaller chunks;
//   * n = 8: four recurrence lanes per environment, each owning one 16-byte piece (two rows of M = I + A dt, 2n doubles
//     in registers); the pieces of a sub-step meet in the tile the row is written to anyway (STS.128, __syncwarp,
//     LDS.128), so n = 8 needs 32 + 16 registers of state instead of 128 + 16 and nothing spills.  n <= 4 keeps one lane
//     per environment and the state in registers.  The row update itself is two independent half-chains
//     (em_row_update, common.cuh): the recurrence is the slowest pipeline stage when its FP64 chain is long;
//   * no role needs more than 80 registers: 768 threads per CTA (6 warps per scheduler instead of 4).
// Arithmetic, operation order and the Philox stream are those of em_kernel.cu / em_tile_kernel.cu: States and
// Observations agree bit for bit (tests/test_em_gpu.py).  Replaces quantrl/envs/stochastic.py:178-242,
// envs/base.py:408-426,462,471-473,485-499 and the packed rows of BaseSB3Env.update_data (envs/base.py:1314-1372).
#include <stdlib.h>
#include "common.cuh"
#include "philox.cuh"

namespace qrl {

constexpr int T3_THREADS_MAX = 768;
constexpr int T3_STAGES_MAX = 8;
constexpr int T3_BAR_BYTES = 256;   // 3 * T3_STAGES_MAX mbarriers, rounded up: the tiles behind them stay 128-byte aligned
constexpr int T3_TAB_BYTES = 8192;  // shared-memory copy of the generator's two lookup tables (philox.cuh)

struct T3Plan { int G, C, D, nphase, nrt, threads, grid; size_t smem; };

__device__ __forceinline__ uint32_t t3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t3_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t3_smem(bar)) : "memory");
}
// Wait for a phase of an mbarrier.  A failed probe backs off with nanosleep: the recurrence warps wait most of the time
// (their stage has 60 % slack) and a tight probe loop was 14 % of all executed warp instructions, issued in competition
// with the generator (ncu r02k: 2.5 M of 17.7 M).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = t3_smem(bar);
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void t3_minmax(double v, double& lo, double& hi) {   // NaN never wins (lo / hi are never NaN)
    lo = (v < lo) ? v : lo;
    hi = (v > hi) ? v : hi;
}

// Development instrumentation (tools/microbench/t3_timeline.cu builds this file with -DQRL_T3_TIMELINE): per warp, the
// cycles spent producing / writing rows / waiting on each kind of mbarrier, and absolute times of the kernel's phases.
#ifdef QRL_T3_TIMELINE
__device__ unsigned long long t3_tl[256][32][8];
#define T3_TL_DECL long long tl_t0 = 0; unsigned long long tl_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define T3_TL_BEGIN tl_t0 = clock64();
#define T3_TL_END(slot) tl_acc[slot] += (unsigned long long)(clock64() - tl_t0);
#define T3_TL_MARK(slot) tl_acc[slot] = global_timer_ns();
__device__ unsigned long long t3_trace[2][256];
#define T3_TRACE(who, idx) if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (who == 0 ? threadIdx.x == 0 : threadIdx.x == NRT) && (idx) < 256) t3_trace[who][idx] = global_timer_ns();
#define T3_TL_FLUSH if ((threadIdx.x & 31) == 0 && blockIdx.x < 256) { for (int i_ = 0; i_ < 8; ++i_) t3_tl[blockIdx.x][threadIdx.x >> 5][i_] = tl_acc[i_]; }
#else
#define T3_TL_DECL
#define T3_TL_BEGIN
#define T3_TL_END(slot)
#define T3_TL_MARK(slot)
#define T3_TL_FLUSH
#define T3_TRACE(who, idx)
#endif

template <int N, int LR>
__global__ void __launch_bounds__(T3_THREADS_MAX, 1)
em_tile3_kernel(const qrl_em_args a_in, const PhiloxKeys keys, const T3Plan pl) {
    static_assert(N == 2 || N == 4 || N == 8, "piece layout: n/2 lanes per environment");
    static_assert(N % LR == 0 && LR <= 8, "LR recurrence lanes per environment, N / LR rows each");
    constexpr int L = N / 2;                   // 16-byte pieces per environment row
    constexpr int RO = N / LR;                 // rows of M per recurrence lane
    // programmatic dependent launch (QRL_EM_PDL): see em_tile_kernel.cu
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    T3_TL_DECL
    T3_TL_MARK(0)
    qrl_em_args a = a_in;
    extern __shared__ __align__(128) unsigned char t3_raw[];
    const int G = pl.G, C = pl.C, D = pl.D, nphase = pl.nphase, NRT = pl.nrt;
    const int NT = blockDim.x, NW = NT - NRT;
    const int E = a.n_envs, K = a.K;
    const int e0 = blockIdx.x * G;
    const int Gl = min(G, E - e0);             // live environments of this CTA
    const int RP = G * N;                      // doubles per tile row
    const int tile = C * RP;                   // doubles per tile
    uint64_t* full_w = reinterpret_cast<uint64_t*>(t3_raw);
    uint64_t* full_x = full_w + T3_STAGES_MAX;
    uint64_t* empty = full_x + T3_STAGES_MAX;
    double* Wt = reinterpret_cast<double*>(t3_raw + T3_BAR_BYTES + T3_TAB_BYTES);   // [D][tile] increments, overwritten by the states
    double* Vt = Wt + D * tile;                // [D][tile] measurement noise, already scaled by obs_std
    double* cumS = Vt + D * tile;              // [G]  cumulative reward before (prologue) / after (last row) this interval
    double* mmS = cumS + G;                    // [64] per-warp min / max partials
    int* selS = reinterpret_cast<int*>(mmS + 64);  // [n_sel]

    const int tid = threadIdx.x;
    const bool is_rec = tid < NRT;
    const int wt = tid - NRT;                  // worker id
    const bool philox = a.rng_mode == QRL_RNG_PHILOX;
    const bool has_reward = a.reward_kind != QRL_REWARD_NONE;
    const bool reward_on_obs = a.reward_kind == QRL_REWARD_NEG_WSQ_OBS;
    const int NC = (K + C) / C;                // chunks of C rows over the K+1 rows of the interval
    const int n_data = 1 + a.n_actions + N + 1;
    const int n_sel = a.n_sel;
    const int pitch = a.packed ? (int)(a.packed_row_pitch > 0 ? a.packed_row_pitch : a.n_sel) : 0;

    if (tid == 0) {
        for (int s = 0; s < D; ++s) {
            mbar_init(full_w + s, (uint32_t)NW);
            mbar_init(full_x + s, (uint32_t)NRT);
            mbar_init(empty + s, (uint32_t)NW);
        }
    }
    // ---- prologue.  Every global load of the launch's set-up is ISSUED before the first of them is consumed, in both
    // roles: one round trip to L2 instead of four serial ones (tables, noise scales, x0, drift rows; the timeline of round
    // 2's first version showed 2.1 us from kernel entry to the barrier and 3.4 us to the first pair of normals).
    // Constants of a step sequence first (no step kernel writes them): behind the PDL wait they cost nothing.
    double2 tab_v = make_double2(0.0, 0.0);
    if (tid < 512) tab_v = tid < 256 ? kRngLogTab[tid] : kRngSinCosTab[tid - 256];
    int sel_v = 0;
    if (a.packed && tid < n_sel) sel_v = a.sel_idxs[tid];
    if ((a.flags & QRL_EM_PDL) && is_rec && (tid % LR) == 0 && tid / LR < Gl) {
        const char* Ab = reinterpret_cast<const char*>(a.A + (int64_t)(e0 + tid / LR) * a.A_stride_e);
        for (int b = 0; b < N * N * 8; b += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(Ab + b));
    }
    // everything below reads what the previous kernel of the stream wrote (dyn, x0, cumulative rewards, actions)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    a = resolve_dyn(a_in);
    double cum_v = 0.0;
    if (tid < Gl && a.rewards_cum) cum_v = a.rewards_cum[e0 + tid];
    RngTabSmem rng_tab;
    rng_tab.lg_base = t3_smem(t3_raw + T3_BAR_BYTES);
    rng_tab.sc_base = rng_tab.lg_base + 4096u;
    // second half of the prologue, called by either role once ITS loads are in flight: shared-memory copies, then the
    // CTA barrier (bar.sync 0 from two call sites: roles are whole warps)
    auto publish_prologue = [&]() {
        double2* tabS = reinterpret_cast<double2*>(t3_raw + T3_BAR_BYTES);
        if (tid < 512) tabS[tid] = tab_v;
        for (int i = tid + NT; i < 512; i += NT) tabS[i] = i < 256 ? kRngLogTab[i] : kRngSinCosTab[i - 256];   // (CTAs below 512 threads)
        if (a.packed) {
            if (tid < n_sel) selS[tid] = sel_v < 0 ? sel_v + n_data : sel_v;
            for (int i = tid + NT; i < n_sel; i += NT) {
                const int s2 = a.sel_idxs[i];
                selS[i] = s2 < 0 ? s2 + n_data : s2;
            }
        }
        if (tid < Gl) cumS[tid] = cum_v;       // cumulative reward BEFORE this interval (the last row adds to it)
        if (a.peer_pulled && tid == 0) em_wait_pulled(a);     // publish-only exchange: flow control before anything is written
        asm volatile("bar.sync 0;" ::: "memory");
        T3_TL_MARK(1)
        T3_TRACE(0, 252)
    };

    const double sqrt_dt = sqrt(a.t_ssz);
    const uint32_t row32 = (uint32_t)(E * N);  // doubles per time row (the dispatcher guarantees (K+1) * E * N < 2^31)
    double lo = INFINITY, hi = -INFINITY;

    // Chunks are cut over ROWS: chunk c holds rows c*C .. c*C+C-1 of the K+1 rows of the interval, row r in slot r - c*C.
    // Row 0 (States[0] = x0, Observations[0] = x0 + noise[t_idx], envs/base.py:462) is simply the first row of chunk 0;
    // row r >= 1 is the result of sub-step r-1.  Every role therefore has ONE loop body for all rows (one inlined copy
    // of each pass: the kernel is sensitive to instruction-cache pressure).
    if (!is_rec) {
        // ================================= workers: produce noise, write rows ==================================

        // ---- output, piece pass: one 16-byte piece (two components) of a row per worker and trip: LDS.128 of states and
        // noise, two adds, STG.128 into the States / Observations blocks (a warp writes 512 contiguous bytes), NaN-ignoring
        // min/max, and the reward as ONE chain of FMAs in component order that hops over the L adjacent lanes of an
        // environment.  (A pass in which every worker owns one piece column and walks the rows with loop-invariant
        // addressing was measured too: fewer instructions per row, but workers / (pieces per row) is rarely an integer and
        // the idle remainder costs more than the invariants save — n = 8, G = 56: 94 vs 85 us.)
        const int PR = Gl * L;                 // pieces per live row
        const int q_dsl = NW / PR, q_dpc = NW - q_dsl * PR;
        const int q_sl0 = wt / PR, q_pc0 = wt - q_sl0 * PR;
        const int q_h = wt & (L - 1);          // which piece of its environment this lane holds (adjacent lanes)
        const double q_w0 = has_reward ? a.reward_w[2 * q_h] : 0.0, q_w1 = has_reward ? a.reward_w[2 * q_h + 1] : 0.0;
        auto piece_pass = [&](const double* Xb, const double* Vb, int r_first, int ns) {
            int sl = q_sl0, pc = q_pc0;
            uint32_t g = (uint32_t)(r_first + sl) * row32 + (uint32_t)(e0 * N + 2 * pc);
            const uint32_t g_step = (uint32_t)q_dsl * row32 + (uint32_t)(2 * q_dpc), g_wrap = row32 - (uint32_t)(2 * PR);
            int sl_lead = __shfl_sync(0xffffffffu, sl, 0);   // warp-uniform trip count (the shuffles need all lanes)
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
                    t3_minmax(o2.x, lo, hi);
                    t3_minmax(o2.y, lo, hi);
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
                            const double cum = a.rewards_cum ? rv + cumS[el] : rv;   // (the old value was fetched in the prologue)
                            if (a.rewards_cum) a.rewards_cum[e] = cum;
                            cumS[el] = cum;
                        }
                        const int64_t gl = (int64_t)e0 * N + 2 * pc;
                        if (a.x_last) *reinterpret_cast<double2*>(a.x_last + gl) = x2;
                        if (a.obs_last) *reinterpret_cast<double2*>(a.obs_last + gl) = o2;
                        if (a.obs_last_b) *reinterpret_cast<double2*>(a.obs_last_b + gl) = o2;
                    }
                }
                sl += q_dsl;
                pc += q_dpc;
                g += g_step;
                if (pc >= PR) { pc -= PR; ++sl; g += g_wrap; }
                sl_lead = __shfl_sync(0xffffffffu, sl, 0);
            }
        };
        // ---- output, cache rows: packed[e, r, j] = column sel[j] of [T_step[r], actions[e,:], Obs[r,e,:], cum. reward]
        // (envs/base.py:1314-1372); a warp walks one environment's contiguous run of ns * pitch doubles 32 at a time.
        // The cumulative-reward column is only known after the last sub-step and is written for all rows at the end.
        const uint32_t pk_magic = pitch > 0 ? 0xFFFFFFFFu / (uint32_t)pitch + 1u : 0u;   // idx / pitch for idx < 2^16
        auto pack_pass = [&](const double* Xb, const double* Vb, int r_first, int ns) {
            const int run = ns * pitch;
            const int trips = (run + 31) >> 5;
            const int lane = wt & 31, n_ww = NW >> 5;
            int t = wt >> 5;
            int el = t / trips, trip = t - el * trips;
            const int d_el = n_ww / trips, d_trip = n_ww - d_el * trips;
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

        // ---- producer: the (env, comp) column of this worker and its position in the row walk ----
        // the call with tau = t_idx + r - 1 gives z0 -> Wiener increment of sub-step r-1 (row r's input) and z1 -> measurement
        // noise of row r; row 0's measurement noise is z1 of the call before (tau = t_idx - 1, or the reset call)
        const int p_phase = wt / RP, p_col = wt - p_phase * RP;
        const int p_el = p_col / N, p_c = p_col - p_el * N;
        const bool p_live = p_phase < nphase && p_el < Gl;
        const uint32_t p_env = (uint32_t)(a.env_offset + e0 + p_el);
        const uint32_t tau_base = (uint32_t)a.t_idx - 1u;   // tau of row r is tau_base + r (row 0 of an episode: 0xFFFFFFFF, the reset call)
        double p_sd = 0.0;
        int pr = p_phase;                      // next row this worker generates (phase 0 starts with row 0)
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (philox && p_live) {
            if (a.obs_std) p_sd = a.obs_std[(int64_t)(e0 + p_el) * a.obs_std_stride_e + p_c];
            nxt = philox4x32_10(make_uint4(tau_base + (uint32_t)pr, a.episode, p_env, (uint32_t)p_c), keys);
        }
        publish_prologue();
        bool pack_rows = false;                // any cache column besides the cumulative reward (written at the end)?
        if (a.packed)
            for (int j = 0; j < n_sel; ++j) pack_rows |= (selS[j] != n_data - 1);

        int s_out = 0, s_in = 0;               // ring stages of the chunk being written out / produced
        uint32_t par_out = 0u, par_in = 1u;    // parities: full_x of the chunk written out; empty of the stage refilled
        for (int it = 0; it < NC + 2; ++it) {
            // ---------------- rows of chunk `it - 2` ----------------
            if (it >= 2) {
                const int co = it - 2;
                T3_TL_BEGIN
                T3_TRACE(1, 4 * it + 0)
                mbar_wait(full_x + s_out, par_out);
                T3_TRACE(1, 4 * it + 1)
                T3_TL_END(4)
                T3_TL_BEGIN
                const double* Xb = Wt + s_out * tile;
                const double* Vb = Vt + s_out * tile;
                const int ns = min(C, K + 1 - co * C);
                piece_pass(Xb, Vb, co * C, ns);
                if (pack_rows) pack_pass(Xb, Vb, co * C, ns);
                T3_TL_END(3)
                mbar_arrive(empty + s_out);
                if (++s_out == D) { s_out = 0; par_out ^= 1u; }
            }
            // ---------------- noise of chunk `it` ----------------
            if (it < NC) {
                T3_TL_BEGIN
                if (it >= D) mbar_wait(empty + s_in, par_in);     // the rows of chunk it - D have left this stage
                T3_TRACE(1, 4 * it + 2)
                T3_TL_END(5)
                T3_TL_BEGIN
                const int r0 = it * C;
                const int r_end = min(r0 + C, K + 1);
                double* Wb = Wt + s_in * tile;
                double* Vb = Vt + s_in * tile;
                if (philox) {
                    if (p_live) {
                        // C is a multiple of nphase (plan), so a worker's walk never straddles a chunk boundary and the
                        // prefetched counter of its next pair stays valid across chunks
                        double* wp = Wb + (pr - r0) * RP + p_col;
                        double* vp = Vb + (pr - r0) * RP + p_col;
                        for (; pr < r_end; pr += nphase, wp += nphase * RP, vp += nphase * RP) {
                            const uint4 cur = nxt;
                            // the rounds of the next pair are issued in the same basic block as this pair's Box–Muller
                            nxt = philox4x32_10(make_uint4(tau_base + (uint32_t)(pr + nphase), a.episode, p_env, (uint32_t)p_c), keys);
                            double z0, z1;
                            normal_pair(cur, rng_tab, z0, z1);
                            *wp = __dmul_rn(sqrt_dt, z0);          // (row 0: never read — the recurrence puts x0 there)
                            *vp = __dmul_rn(p_sd, z1);
                        }
                    }
                } else {
                    // fed noise: 4 independent loads in flight per lane.  Row r takes W[r-1] (r >= 1) and obs_noise[r]
                    const int ns = r_end - r0;
                    const int RL = Gl * N;
                    const int total = ns * RL;
                    const int64_t g0 = (int64_t)r0 * row32 + (int64_t)e0 * N;
                    for (int base = wt; base < total; base += 4 * NW) {
                        double wv[4], vv[4];
                        int sidx[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int idx = base + u * NW;
                            wv[u] = vv[u] = 0.0;
                            sidx[u] = -1;
                            if (idx < total) {
                                const int s2 = idx / RL, rem = idx - s2 * RL;
                                const int64_t gi = g0 + (int64_t)s2 * row32 + rem;
                                sidx[u] = s2 * RP + rem;
                                if (r0 + s2 > 0) wv[u] = a.W[gi - row32];
                                if (a.obs_noise) vv[u] = a.obs_noise[gi];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (sidx[u] >= 0) { Wb[sidx[u]] = wv[u]; Vb[sidx[u]] = vv[u]; }
                    }
                }
                T3_TL_END(2)
                T3_TRACE(1, 4 * it + 3)
                mbar_arrive(full_w + s_in);
                if (++s_in == D) { s_in = 0; par_in ^= 1u; }
            }
        }
    } else {
        // ================================= recurrence: LR lanes per environment ==================================
        // lane q of an environment owns RO = N / LR consecutive rows of M = I + A dt and the matching pieces of every tile
        // row.  One lane per environment keeps the state in registers; with several lanes the pieces of a sub-step meet in
        // the tile the row is written to anyway (STS, __syncwarp, LDS.128 broadcast)
        const int el_r = tid / LR, q = tid - el_r * LR;
        const bool rec_live = el_r < Gl;
        const int el_c = rec_live ? el_r : Gl - 1;     // dead lanes shadow the last environment (loads only)
        double M[RO][N], nzr[RO], x[N];
        {
            const int e = e0 + el_c;
            const double* Ar = a.A + (int64_t)e * a.A_stride_e + (int64_t)(RO * q) * N;
#pragma unroll
            for (int i = 0; i < RO; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c) M[i][c] = Ar[i * N + c];
#pragma unroll
            for (int i = 0; i < RO; ++i) nzr[i] = a.nz[(int64_t)e * a.nz_stride_e + RO * q + i];
#pragma unroll
            for (int c = 0; c < N; ++c) x[c] = a.x0[(int64_t)e * N + c];
        }
        double x0r[RO];                        // this lane's piece of row 0
#pragma unroll
        for (int i = 0; i < RO; ++i) x0r[i] = a.x0[(int64_t)(e0 + el_c) * N + RO * q + i];
        // The raw actions are ISSUED here and consumed behind the CTA barrier: the workers (whose first chunk of noise does
        // not depend on the actions) must not wait for them — with step(host actions) they come from pinned host memory
        // over PCIe, a ~2 us round trip that now hides behind the production of chunk 0.
        double u0 = 0.0;
        if (a.A_act) u0 = a.actions[(int64_t)(e0 + el_c) * a.n_actions];
        publish_prologue();
        {
            const int e = e0 + el_c;
            if (a.A_act) {
                // action-affine drift functor, same accumulation order as em_drift(): A += u_k * A_k for k = 0, 1, ...
#pragma unroll 1
                for (int k = 0; k < a.n_actions; ++k) {
                    const double u = (k == 0 ? u0 : a.actions[(int64_t)e * a.n_actions + k]) * (a.action_scale ? a.action_scale[k] : 1.0);
                    const double* Ak = a.A_act + (int64_t)k * N * N + (int64_t)(RO * q) * N;
#pragma unroll
                    for (int i = 0; i < RO; ++i)
#pragma unroll
                        for (int c = 0; c < N; ++c) M[i][c] = fma(u, Ak[i * N + c], M[i][c]);
                }
            }
#pragma unroll
            for (int i = 0; i < RO; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c)
                    M[i][c] = __dadd_rn((RO * q + i == c) ? 1.0 : 0.0, __dmul_rn(M[i][c], a.t_ssz));   // unfused like the reference
        }
        int s = 0;
        uint32_t par = 0u;
        for (int c = 0; c < NC; ++c) {
            T3_TL_BEGIN
            T3_TRACE(0, 2 * c)
            mbar_wait(full_w + s, par);
            T3_TRACE(0, 2 * c + 1)
            T3_TL_END(4)
            T3_TL_BEGIN
            const int ns = min(C, K + 1 - c * C);
            double* Wb = Wt + s * tile + el_c * N + RO * q;     // this lane's pieces in every row of the stage
            double w[RO];
            auto load_w = [&](const double* src) {
                if constexpr (RO == 1) {
                    w[0] = src[0];
                } else {
#pragma unroll
                    for (int i = 0; i < RO / 2; ++i) {
                        const double2 t = *reinterpret_cast<const double2*>(src + 2 * i);
                        w[2 * i] = t.x; w[2 * i + 1] = t.y;
                    }
                }
            };
            auto store_x = [&](double* dst, const double (&v)[RO]) {
                if constexpr (RO == 1) {
                    dst[0] = v[0];
                } else {
#pragma unroll
                    for (int i = 0; i < RO / 2; ++i) *reinterpret_cast<double2*>(dst + 2 * i) = make_double2(v[2 * i], v[2 * i + 1]);
                }
            };
            int slot = 0;
            if (c == 0) {                      // row 0 = the initial state
                if (rec_live) store_x(Wb, x0r);
                slot = 1;
            }
            if (slot < ns) load_w(Wb + slot * RP);
#pragma unroll 2
            for (; slot < ns; ++slot) {
                double xn[RO];
#pragma unroll
                for (int i = 0; i < RO; ++i) xn[i] = em_row_update<N>(M[i], x, nzr[i], w[i]);
                double* xrow = Wb + slot * RP;
                if (slot + 1 < ns) load_w(xrow + RP);      // next increments: off the chain
                if (rec_live) store_x(xrow, xn);
                if constexpr (LR == 1) {
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] = xn[i];
                } else {
                    __syncwarp();              // the LR lanes of an environment sit in one warp
                    const double* xall = xrow - RO * q;
#pragma unroll
                    for (int p2 = 0; p2 < N / 2; ++p2) {
                        const double2 t = *reinterpret_cast<const double2*>(xall + 2 * p2);
                        x[2 * p2] = t.x; x[2 * p2 + 1] = t.y;
                    }
                }
            }
            T3_TL_END(2)
            mbar_arrive(full_x + s);
            if (++s == D) { s = 0; par ^= 1u; }
        }
        if (rec_live && a.nan_flag) {
            bool bad = false;
#pragma unroll
            for (int c = 0; c < N; ++c) bad |= !isfinite(x[c]);
            if (bad) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
        }
    }
    T3_TL_MARK(6)
    T3_TRACE(0, 254)
    T3_TRACE(1, 254)
    // block min / max: warp partials meet in shared memory across the barrier that ends the pipeline anyway
    if (a.obs_minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, shfl_xor_d(lo, o));
            hi = fmax(hi, shfl_xor_d(hi, o));
        }
        if ((tid & 31) == 0) { mmS[tid >> 5] = lo; mmS[32 + (tid >> 5)] = hi; }
    }
    __syncthreads();
    T3_TRACE(0, 255)
    // cumulative-reward column of every cache row of this interval (envs/base.py:1311, 1342-1349): one warp per env,
    // lanes along the rows of that env's contiguous run
    if (a.packed) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int j = 0; j < n_sel; ++j) {
            if (selS[j] != n_data - 1) continue;
            for (int el = warp; el < Gl; el += NT / 32) {
                double* dst = a.packed + (int64_t)(e0 + el) * a.packed_stride_e + j;
                const double cum = cumS[el];
                for (int r = lane; r <= K; r += 32) dst[(int64_t)r * pitch] = cum;
            }
        }
    }
    if (a.obs_minmax && tid < 32) {            // one atomic pair per CTA (NaN never wins, like the reference's comparisons)
        lo = tid < NT / 32 ? mmS[tid] : INFINITY;
        hi = tid < NT / 32 ? mmS[32 + tid] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, shfl_xor_d(lo, o));
            hi = fmax(hi, shfl_xor_d(hi, o));
        }
        if (tid == 0) {
            atomic_min_d(a.obs_minmax + 0, lo);
            atomic_max_d(a.obs_minmax + 1, hi);
        }
    }
    T3_TL_MARK(7)
    T3_TL_FLUSH
    em_finish_launch(a);
}

// Tile plan.  G environments per CTA (one wave of CTAs whenever the batch allows), n/2 recurrence lanes per environment,
// every remaining warp a worker; nphase workers share one (env, comp) column and interleave its sub-steps; the ring holds
// D stages of C sub-steps (C a multiple of nphase).  Tuning knobs: QRL_EM_T3_C, QRL_EM_T3_D, QRL_EM_T3_THREADS.
static int t3_rec_lanes(int N) {
    static const char* lr_env = getenv("QRL_EM_T3_LR");
    // measured on B200 (tools/microbench/t3_timeline.cu): under FP64-pipe contention a recurrence warp advances at ~12
    // cycles per FP64 instruction, so the stage time is set by the FP64 instructions PER LANE and sub-step: n = 4 with one
    // lane per environment (24) takes 16-18 us per 100 sub-steps, two lanes (12) 8 us, four lanes (6 + exchange) 6.6 us
    const int dflt = N == 8 ? 4 : (N == 4 ? 2 : 1);
    int lr = lr_env ? atoi(lr_env) : dflt;
    if (lr < 1 || lr > N || (N % lr) != 0 || (N == 8 && lr < 4)) lr = dflt;   // n = 8 needs <= 2 rows per lane
    return lr;
}

static T3Plan t3_plan(int E, int N, int K, int n_sel, int sms) {
    T3Plan p;
    const int L = t3_rec_lanes(N);             // recurrence lanes per environment (LR of the kernel)
    static const char* th_env = getenv("QRL_EM_T3_THREADS");
    int threads = th_env ? atoi(th_env) : T3_THREADS_MAX;
    threads = threads / 32 * 32;
    if (threads < 128 || threads > T3_THREADS_MAX) threads = T3_THREADS_MAX;
    // workers >= columns: G * N <= threads - (recurrence lanes rounded up to warps)
    int g_cap = 64;
    while (g_cap > 2 && g_cap * N + (g_cap * L + 31) / 32 * 32 > threads) g_cap -= 2;
    int G = (E + sms - 1) / sms;
    if (G & 1) ++G;                            // even G keeps every tile row a whole number of 16-byte pieces for n = 2 too
    if (G > g_cap) {                           // several waves: balance them
        const int waves = (E + g_cap * sms - 1) / (g_cap * sms);
        G = (E + waves * sms - 1) / (waves * sms);
        if (G & 1) ++G;
        if (G > g_cap) G = g_cap;
    }
    if (G < 2) G = 2;
    p.G = G;
    p.grid = (E + G - 1) / G;
    p.nrt = (G * L + 31) / 32 * 32;
    p.threads = threads;
    const int NW = threads - p.nrt, RP = G * N;
    p.nphase = NW / RP;
    if (p.nphase < 1) p.nphase = 1;
    if (p.nphase > K) p.nphase = K;
    static const char* d_env = getenv("QRL_EM_T3_D");
    static const char* c_env = getenv("QRL_EM_T3_C");
    p.D = d_env ? atoi(d_env) : 3;
    if (p.D < 3) p.D = 3;
    if (p.D > T3_STAGES_MAX) p.D = T3_STAGES_MAX;
    const size_t budget = 208 * 1024;
    const size_t fixed = T3_BAR_BYTES + T3_TAB_BYTES + (size_t)(G + 64) * sizeof(double) + (size_t)n_sel * sizeof(int) + 64;
    const size_t per_row = (size_t)2 * RP * sizeof(double);
    int c_want = c_env ? atoi(c_env) : 36;     // few, long chunks: every chunk costs ~0.2 us of hand-over per role
    if (c_want < 1) c_want = 36;
    int c_max = (int)((budget - fixed) / (per_row * p.D));
    while (c_max < p.nphase && p.D > 3) { --p.D; c_max = (int)((budget - fixed) / (per_row * p.D)); }
    if (c_max < p.nphase) p.nphase = c_max > 0 ? c_max : 1;
    int C = c_want < c_max ? c_want : c_max;
    if (C > K + 1) C = (K + p.nphase) / p.nphase * p.nphase;     // one chunk of K+1 rows: still a multiple of nphase
    C = C / p.nphase * p.nphase;
    if (C < p.nphase) C = p.nphase;
    p.C = C;
    p.smem = fixed + per_row * p.D * p.C;
    return p;
}

template <int N, int LR>
static int launch_tile3(const qrl_em_args& a, cudaStream_t st, int sms) {
    const T3Plan p = t3_plan(a.n_envs, N, a.K, a.packed ? a.n_sel : 0, sms);
    static OncePerDevice smem_attr;
    if (smem_attr.first())
        QRL_CUDA(cudaFuncSetAttribute(em_tile3_kernel<N, LR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const PhiloxKeys keys = make_philox_keys(a.seed);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (a.flags & QRL_EM_PDL) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QRL_CUDA(cudaLaunchKernelEx(&cfg, em_tile3_kernel<N, LR>, a, keys, p));
    return after_launch("em_tile3_kernel");
}

// Does the third-generation kernel take this launch?  n in {2, 4, 8}, 16-byte aligned row blocks and last-row buffers,
// 32-bit row offsets; everything else stays with em_tile_kernel.cu / em_kernel.cu.
bool em_tile3_supports(const qrl_em_args& a) {
    if (!(a.n == 2 || a.n == 4 || a.n == 8) || a.K < 1 || a.A_stride_k != 0 || a.nz_stride_k != 0) return false;
    const uintptr_t al = (uintptr_t)a.states | (uintptr_t)a.observations | (uintptr_t)a.x_last | (uintptr_t)a.obs_last |
                         (uintptr_t)a.obs_last_b;
    if (al % 16 != 0) return false;
    if ((int64_t)(a.K + 1) * a.n_envs * a.n >= (1ll << 31)) return false;
    if (a.flags & (QRL_EM_ITEM_OUTPUT | 2 | 4 | 8)) return false;     // diagnostics of the second-generation kernel
    static const char* off = getenv("QRL_EM_NO_T3");
    return !(off && atoi(off) != 0);
}

int em_tile3_launch(const qrl_em_args& a, cudaStream_t st) {
    const int sms = device_sm_count();
    const int lr = t3_rec_lanes(a.n);
    switch (a.n * 16 + lr) {
        case 2 * 16 + 1: return launch_tile3<2, 1>(a, st, sms);
        case 2 * 16 + 2: return launch_tile3<2, 2>(a, st, sms);
        case 4 * 16 + 1: return launch_tile3<4, 1>(a, st, sms);
        case 4 * 16 + 2: return launch_tile3<4, 2>(a, st, sms);
        case 4 * 16 + 4: return launch_tile3<4, 4>(a, st, sms);
        case 8 * 16 + 4: return launch_tile3<8, 4>(a, st, sms);
        case 8 * 16 + 8: return launch_tile3<8, 8>(a, st, sms);
        default: return QRL_E_UNSUPPORTED;
    }
}

}  // namespace qrl
