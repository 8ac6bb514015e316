// Tiled, warp-specialised Euler–Maruyama action interval — the fast path of qrl_em_step on sm_100a.
//
// Why: with one lane per state component (em_kernel.cu) config 3 (E = 4096) puts a single warp on each SM
// sub-partition and the kernel is latency bound (ncu r1a: IPC 0.25, FP64 pipe 12 %).  The noise generation
// (Philox4x32-10 + FP64 Box–Muller, ~95 % of the instructions) does not depend on the state, so it is spread over
// 15 "worker" warps per CTA, while ONE warp runs the short sequential recurrence for all environments of the CTA.
//
// One CTA owns G <= 32 environments for the whole interval; time is cut into chunks of C sub-steps that flow through
// double-buffered shared-memory tiles in a 3-stage pipeline, one __syncthreads per chunk:
//   iteration it:  workers   produce noise(chunk it)      -> Wt/Vt[it & 1]          (Philox, or fed loads)
//                  warp 0    recurrence(chunk it-1)        Wt[(it-1) & 1] -> Xt[(it-1) & 1]   (x-update only: the warp
//                                                          is the sequential critical path, ncu r1b)
//                  workers   output(chunk it-2)            Ot = Xt[it & 1] + Vt[it & 1], then -> global:
//                              * Observations rows: lane-contiguous (fully coalesced) 8-byte stores;
//                              * States rows: each tile row is one contiguous run of the (K+1,E,n) block and
//                                leaves through the TMA engine (cp.async.bulk shared -> global), no LSU work;
//                              * Reward rows: one coalesced 8-byte store per (row, env);
//                              * packed cache rows [t, actions, obs, cum. reward] (env-major): element-wise,
//                                lane-contiguous 8-byte stores;
//                              * min/max of the observations (truncation check).
// Replaces the same reference code as em_kernel.cu (quantrl/envs/stochastic.py:178-242, envs/base.py:408-426,462,
// 471-473,485-499) plus the packed rows of BaseSB3Env.update_data (envs/base.py:1314-1372).
#include <stdlib.h>
#include "common.cuh"
#include "philox.cuh"

namespace qrl {

#ifndef QRL_EM_TILE_THREADS
#define QRL_EM_TILE_THREADS 512
#endif
#ifndef QRL_EM_TILE_ILP
#define QRL_EM_TILE_ILP 1
#endif
constexpr int EM_TILE_THREADS = QRL_EM_TILE_THREADS;
constexpr int EM_TILE_WORKERS = EM_TILE_THREADS - 32;
constexpr int EM_TILE_ILP = QRL_EM_TILE_ILP;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before the barrier that precedes the copy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" :: "n"(EM_TILE_WORKERS) : "memory"); }

template <int N>
__global__ void __launch_bounds__(EM_TILE_THREADS, 1)
em_tile_kernel(const qrl_em_args a_in, const PhiloxKeys keys, const int G, const int C) {
    const qrl_em_args a = resolve_dyn(a_in);
    extern __shared__ __align__(128) double sm[];
    const int E = a.n_envs, K = a.K;
    const int e0 = blockIdx.x * G;
    const int Gl = min(G, E - e0);             // live environments of this CTA
    const int RP = Gl * N;                     // doubles per tile row
    const int tile = C * G * N;                // doubles per tile buffer (even: G is even, see plan_tiles)
    double* NB = sm;                           // [2][2][tile] noise buffers {W, V}[b]: Wiener increments and measurement
                                               // noise (scaled) of one chunk sit next to each other, so that the pair
                                               // doubles as the staging area of the packed cache rows once consumed
    double* Xt = NB + 4 * tile;                // [2][tile] states, written by the recurrence warp
    double* Ot = Xt + 2 * tile;                // [tile]    observations = states + noise, written by the workers
    double* Pt = Ot + tile;                    // [G][C][pitch] packed cache rows of one chunk, in each env's global layout
    const int pitch = a.packed ? (int)(a.packed_row_pitch > 0 ? a.packed_row_pitch : a.n_sel) : 0;
    double* x0S = Pt + (size_t)C * G * pitch;  // [G*N] initial state (row 0)
    double* stdS = x0S + G * N;                // [G*N] measurement-noise std per (env, comp)
    double* cumS = stdS + G * N;               // [G]   cumulative reward after this interval
    double* wS = cumS + G;                     // [N]   reward weights
    int* selS = reinterpret_cast<int*>(wS + N);  // [n_sel]

    const int tid = threadIdx.x;
    const bool is_rec = tid < 32;
    const int wt = tid - 32;                   // worker lane id (0..EM_TILE_WORKERS-1)
    const bool philox = a.rng_mode == QRL_RNG_PHILOX;
    const bool has_reward = a.reward_kind != QRL_REWARD_NONE;
    const double sqrt_dt = sqrt(a.t_ssz);
    const int NC = (K + C - 1) / C;
    const int64_t row = (int64_t)E * N;
    const int n_data = 1 + a.n_actions + N + 1;
    const int n_sel = a.n_sel;

    for (int i = tid; i < RP; i += EM_TILE_THREADS) {
        const int el = i / N, c = i - el * N;
        stdS[i] = (philox && a.obs_std) ? a.obs_std[(int64_t)(e0 + el) * a.obs_std_stride_e + c] : 0.0;
        x0S[i] = a.x0[(int64_t)e0 * N + i];
    }
    if (tid < N) wS[tid] = has_reward ? a.reward_w[tid] : 0.0;
    if (a.packed)
        for (int i = tid; i < n_sel; i += EM_TILE_THREADS) {
            const int s = a.sel_idxs[i];
            selS[i] = s < 0 ? s + n_data : s;
        }
    __syncthreads();
    bool sel_identity = a.packed != nullptr && n_sel == n_data;
    if (sel_identity)
        for (int j = 0; j < n_sel; ++j) sel_identity &= (selS[j] == j);

    // TMA eligibility of the packed rows: every env's run of this CTA starts 16-byte aligned and has a 16-byte size
    const bool tma_p = a.packed != nullptr && (pitch % 2 == 0) && (a.packed_stride_e % 2 == 0) &&
                       ((uintptr_t)a.packed % 16 == 0);

    double lo = INFINITY, hi = -INFINITY;

    // ---- worker helpers: everything derived from a finished observation row block in Ot ------------------------
    // reward rows, last-row outputs and packed cache rows for rows r_first .. r_first+ns-1 whose states/observations
    // sit in Xb/Ob (row-major [slot][env][comp])
    auto emit_derived = [&](const double* Xb, const double* Ob, int r_first, int ns, double* stage) {
        if (a.packed) {
            // Rows are built in shared memory as [env][slot][pitch] = the global layout of each env's run, then leave
            // as ONE TMA bulk store per env (ns * pitch * 8 contiguous bytes) when the 16-byte rules hold, else through
            // lane-contiguous 8-byte stores.  (Direct per-row stores from the lanes cost +12 us per launch at config 3:
            // ~300 instructions per row of address arithmetic and 8-byte pieces of 7 different sectors.)
            const int per_env = ns * pitch;
            const bool staged = stage != nullptr;
            int el = wt / ns, sl = wt - el * ns;
            const int d_el = EM_TILE_WORKERS / ns, d_sl = EM_TILE_WORKERS - d_el * ns;
            while (el < Gl) {
                const int r = r_first + sl, e = e0 + el;
                const double* orow = Ob + sl * RP + el * N;
                double* dst = staged ? stage + (el * ns + sl) * pitch
                                     : a.packed + (int64_t)e * a.packed_stride_e + (int64_t)r * pitch;
                if (sel_identity) {
                    // all n_data columns in their natural order (the reference's cache_all_data layout): straight copy
                    dst[0] = a.T_step[r];
                    for (int k = 0; k < a.n_actions; ++k) dst[1 + k] = em_action(a, e, k);
#pragma unroll
                    for (int c = 0; c < N; ++c) dst[1 + a.n_actions + c] = orow[c];
                    dst[n_data - 1] = 0.0;   // cumulative reward: patched after the last sub-step
                } else {
                    for (int j = 0; j < n_sel; ++j) {
                        const int col = selS[j];
                        double val = 0.0;  // cumulative-reward column: only known after the last sub-step, patched below
                        if (col == 0) val = a.T_step[r];
                        else if (col <= a.n_actions) val = em_action(a, e, col - 1);
                        else if (col <= a.n_actions + N) val = orow[col - 1 - a.n_actions];
                        dst[j] = val;
                    }
                }
                if (staged)
                    for (int j = n_sel; j < pitch; ++j) dst[j] = 0.0;   // padding columns travel with the bulk store
                el += d_el;
                sl += d_sl;
                if (sl >= ns) { sl -= ns; ++el; }
            }
            if (staged) {
                const bool tma_now = tma_p && (((int64_t)r_first * pitch) % 2 == 0) && (per_env % 2 == 0);
                if (tma_now) fence_async_smem();
                worker_barrier();
                if (tma_now) {
                    if (wt < 32) {
                        for (int e2 = wt; e2 < Gl; e2 += 32)
                            bulk_store(a.packed + (int64_t)(e0 + e2) * a.packed_stride_e + (int64_t)r_first * pitch,
                                       stage + e2 * per_env, per_env * 8);
                        bulk_commit();
                    }
                } else {
                    int e2 = wt / per_env, off = wt - e2 * per_env;
                    const int d_e2 = EM_TILE_WORKERS / per_env, d_off = EM_TILE_WORKERS - d_e2 * per_env;
                    while (e2 < Gl) {
                        a.packed[(int64_t)(e0 + e2) * a.packed_stride_e + (int64_t)r_first * pitch + off] = stage[e2 * per_env + off];
                        e2 += d_e2;
                        off += d_off;
                        if (off >= per_env) { off -= per_env; ++e2; }
                    }
                }
            }
        }
        if (has_reward || a.x_last || a.obs_last) {
            for (int item = wt; item < ns * Gl; item += EM_TILE_WORKERS) {
                const int sl = item / Gl, el = item - sl * Gl;
                const int r = r_first + sl, e = e0 + el;
                const double* xr = Xb + sl * RP + el * N;
                const double* orow = Ob + sl * RP + el * N;
                double rv = 0.0;
                if (has_reward) {
                    const double* s = (a.reward_kind == QRL_REWARD_NEG_WSQ_OBS) ? orow : xr;
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < N; ++c) acc = fma(wS[c] * s[c], s[c], acc);
                    rv = -acc;
                    if (a.reward) a.reward[(int64_t)r * E + e] = rv;
                }
                if (r == K) {
                    if (has_reward) {
                        if (a.reward_last) a.reward_last[e] = rv;
                        double cum = rv;
                        if (a.rewards_cum) { cum += a.rewards_cum[e]; a.rewards_cum[e] = cum; }
                        cumS[el] = cum;
                    }
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        if (a.x_last) a.x_last[(int64_t)e * N + c] = xr[c];
                        if (a.obs_last) a.obs_last[(int64_t)e * N + c] = orow[c];
                    }
                }
            }
        }
    };

    // ---- recurrence warp: coefficients and state in registers ----------------------------------------------------
    double M[N][N], nz[N], x[N];
    const int el_r = tid;                      // env handled by this recurrence lane
    const bool rec_live = is_rec && el_r < Gl;
    if (rec_live) {
        const int e = e0 + el_r;
        const double* Ar = a.A + (int64_t)e * a.A_stride_e;
#pragma unroll
        for (int r = 0; r < N; ++r) {
#pragma unroll
            for (int c = 0; c < N; ++c) M[r][c] = __dadd_rn((r == c) ? 1.0 : 0.0, __dmul_rn(em_drift(a, Ar, e, r * N + c), a.t_ssz));
            nz[r] = a.nz[(int64_t)e * a.nz_stride_e + r];
            x[r] = x0S[el_r * N + r];
        }
    } else if (!is_rec) {
        // ---- workers: row 0, Observations[0] = States[0] + noise[t_idx]  (envs/base.py:462) ----
        for (int i = wt; i < RP; i += EM_TILE_WORKERS) {
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
            const double xv = x0S[i], o = __dadd_rn(xv, v);
            Ot[i] = o;
            if (a.states) a.states[(int64_t)e0 * N + i] = xv;
            if (a.observations) a.observations[(int64_t)e0 * N + i] = o;
            lo = fmin(lo, o);
            hi = fmax(hi, o);
        }
        worker_barrier();
        emit_derived(x0S, Ot, 0, 1, nullptr);
    }

    // incremental (slot, rem) walkers avoid runtime integer divisions in the hot loops
    const int p_dslot = EM_TILE_WORKERS / RP, p_drem = EM_TILE_WORKERS - p_dslot * RP;
    const bool dbg_no_prod = a.flags & 2, dbg_no_rec = a.flags & 4, dbg_no_out = a.flags & 8;  // diagnostics only

    for (int it = 0; it < NC + 2; ++it) {
        if (!is_rec) {
            const bool out_now = it >= 2;
            const int co = it - 2;
            // ---------------- output chunk `it - 2`, part A: observations = states + noise ----------------
            if (out_now) {
                const int ns = min(C, K - co * C);
                const double* Xb = Xt + (co & 1) * tile;
                const double* Vb = NB + (co & 1) * 2 * tile + tile;
                const int r0 = co * C;
                if (wt < 32) bulk_wait_read();   // the bulk stores of the previous chunk have read Pt by now (long ago)
                if (!dbg_no_out) {
                    // states and observations leave through the LSU right here: lane-contiguous 8-byte stores = fully
                    // coalesced runs of the (K+1,E,n) blocks.  (TMA bulk stores of the state rows were tried: the
                    // recurrence warp may not overwrite its tile before the engine has read it, and waiting for that
                    // at the end of every iteration cost more barrier time than the stores cost issue slots.)
                    int s2 = wt / RP, rem = wt - s2 * RP;
                    const int64_t gbase = (int64_t)(r0 + 1) * row + (int64_t)e0 * N;
                    while (s2 < ns) {
                        const int idx = s2 * RP + rem;
                        const double o = __dadd_rn(Xb[idx], Vb[idx]);
                        Ot[idx] = o;
                        if (a.observations) a.observations[gbase + (int64_t)s2 * row + rem] = o;
                        if (a.states) a.states[gbase + (int64_t)s2 * row + rem] = Xb[idx];
                        lo = fmin(lo, o);
                        hi = fmax(hi, o);
                        s2 += p_dslot;
                        rem += p_drem;
                        if (rem >= RP) { rem -= RP; ++s2; }
                    }
                }
                worker_barrier();   // Ot complete; every read of Vt[it & 1] is done before it is refilled below
                // ---------------- part B: reward rows, last-row outputs, packed rows ----------------
                if (!dbg_no_out) {
                    // Packed rows: direct per-row stores by default.  The shared-memory-staged variant (one TMA bulk
                    // store per env run, flag QRL_EM_PACK_TMA) moves the bytes for free but needs a proxy fence and a
                    // second worker barrier per chunk and measured slower on B200 (49.7 vs 43.3 us at config 3).
                    emit_derived(Xb, Ot, r0 + 1, ns, (a.flags & QRL_EM_PACK_TMA) ? Pt : nullptr);
                }
            }
            // ---------------- produce noise of chunk `it` ----------------
            if (it < NC && !dbg_no_prod) {
                const int j0 = it * C;
                const int ns = min(C, K - j0);
                double* Wb = NB + (it & 1) * 2 * tile;
                double* Vb = NB + (it & 1) * 2 * tile + tile;
                int slot = wt / RP, rem = wt - slot * RP;
                if (philox) {
                    // Software pipeline: the Philox rounds (integer ALU / IMAD pipes) of the NEXT pair are issued inside
                    // the same straight-line block as the Box-Muller polynomials (FP64 pipe) of the CURRENT pair, so the
                    // half-rate pipes overlap instead of all warps sitting in the same phase (ncu r1c: issue slots 50 %
                    // busy with ALU 31 %, FMA-heavy 29 %, FP64 20 % when the phases ran back to back).
                    // EM_TILE_ILP pairs per trip give that many independent DFMA chains (DFMA latency 8.4 cycles, measured).
                    const uint32_t envg0 = (uint32_t)(a.env_offset + e0);
                    const uint32_t tau0 = (uint32_t)(a.t_idx + j0);
                    int sl[EM_TILE_ILP], rm[EM_TILE_ILP];
                    uint4 bits[EM_TILE_ILP];
#pragma unroll
                    for (int u = 0; u < EM_TILE_ILP; ++u) {
                        sl[u] = slot;
                        rm[u] = rem;
                        bits[u] = philox4x32_10(make_uint4(tau0 + slot, a.episode, envg0 + rem / N, (uint32_t)(rem % N)), keys);
                        slot += p_dslot;
                        rem += p_drem;
                        if (rem >= RP) { rem -= RP; ++slot; }
                    }
                    while (sl[0] < ns) {
                        uint4 cur[EM_TILE_ILP];
                        int idx[EM_TILE_ILP];
                        double sd[EM_TILE_ILP];
                        bool ok[EM_TILE_ILP];
#pragma unroll
                        for (int u = 0; u < EM_TILE_ILP; ++u) {
                            cur[u] = bits[u];
                            ok[u] = sl[u] < ns;
                            idx[u] = sl[u] * RP + rm[u];
                            sd[u] = stdS[rm[u]];
                            // next pair of this chain: unconditional (a pair past the end of the chunk is computed and
                            // dropped), so the whole trip is one basic block the scheduler can interleave freely
                            sl[u] = slot;
                            rm[u] = rem;
                            bits[u] = philox4x32_10(make_uint4(tau0 + slot, a.episode, envg0 + rem / N, (uint32_t)(rem % N)), keys);
                            slot += p_dslot;
                            rem += p_drem;
                            if (rem >= RP) { rem -= RP; ++slot; }
                        }
                        double z0[EM_TILE_ILP], z1[EM_TILE_ILP];
#pragma unroll
                        for (int u = 0; u < EM_TILE_ILP; ++u) normal_pair(cur[u], z0[u], z1[u]);
#pragma unroll
                        for (int u = 0; u < EM_TILE_ILP; ++u) {
                            if (ok[u]) {
                                Wb[idx[u]] = __dmul_rn(sqrt_dt, z0[u]);
                                Vb[idx[u]] = __dmul_rn(sd[u], z1[u]);
                            }
                        }
                    }
                } else {
                    // fed noise: 4 independent loads in flight per lane (the loop is global-latency bound otherwise)
                    const int total = ns * RP;
                    const int64_t g0 = (int64_t)j0 * row + (int64_t)e0 * N;
                    for (int base = wt; base < total; base += 4 * EM_TILE_WORKERS) {
                        double wv[4], vv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int idx = base + u * EM_TILE_WORKERS;
                            wv[u] = vv[u] = 0.0;
                            if (idx < total) {
                                const int s2 = idx / RP;
                                const int64_t g = g0 + (int64_t)s2 * row + (idx - s2 * RP);
                                wv[u] = a.W[g];
                                if (a.obs_noise) vv[u] = a.obs_noise[g + row];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int idx = base + u * EM_TILE_WORKERS;
                            if (idx < total) { Wb[idx] = wv[u]; Vb[idx] = vv[u]; }
                        }
                    }
                }
            }
        } else if (it >= 1 && it <= NC && rec_live && !dbg_no_rec) {
            // ---------------- recurrence of chunk `it - 1` (the only sequential part) ----------------
            const int cr = it - 1;
            const int ns = min(C, K - cr * C);
            const double* Wb = NB + (cr & 1) * 2 * tile + el_r * N;
            double* Xb = Xt + (cr & 1) * tile + el_r * N;
            for (int slot = 0; slot < ns; ++slot) {
                double w[N], xn[N];
                if constexpr (N % 2 == 0) {   // RP even and 16-byte aligned tiles: 128-bit shared-memory accesses
                    const double2* wp = reinterpret_cast<const double2*>(Wb + slot * RP);
#pragma unroll
                    for (int c = 0; c < N / 2; ++c) { const double2 t = wp[c]; w[2 * c] = t.x; w[2 * c + 1] = t.y; }
                } else {
#pragma unroll
                    for (int c = 0; c < N; ++c) w[c] = Wb[slot * RP + c];
                }
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double acc = __dmul_rn(M[r][0], x[0]);
#pragma unroll
                    for (int c = 1; c < N; ++c) acc = fma(M[r][c], x[c], acc);
                    xn[r] = fma(nz[r], w[r], acc);
                }
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
        if (bad) *a.nan_flag = 1;
    }
    // cumulative-reward column of every packed row of this interval (envs/base.py:1311, 1342-1349)
    if (a.packed) {
        for (int j = 0; j < n_sel; ++j) {
            if (selS[j] != n_data - 1) continue;
            for (int item = tid; item < Gl * (K + 1); item += EM_TILE_THREADS) {
                const int el = item / (K + 1), r = item - el * (K + 1);
                a.packed[(int64_t)(e0 + el) * a.packed_stride_e + (int64_t)r * pitch + j] = cumS[el];
            }
        }
    }
    if (!is_rec && wt < 32) bulk_wait_all();   // the bulk stores have reached global memory
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

struct TilePlan { int G, C, grid; size_t smem; };

static TilePlan plan_tiles(int E, int N, int K, int n_sel, int pitch, int sms) {
    TilePlan p;
    p.G = E <= 32 * sms ? (E + sms - 1) / sms : 32;
    if (p.G & 1) ++p.G;                        // even G keeps every tile row 16-byte aligned for the bulk stores
    if (p.G > 32) p.G = 32;
    if (p.G < 2) p.G = 2;
    p.grid = (E + p.G - 1) / p.G;
    const size_t budget = 200 * 1024;
    const size_t fixed = (size_t)(2 * p.G * N + p.G + N) * sizeof(double) + (size_t)n_sel * sizeof(int) + 256;
    const size_t per_c = (size_t)(7 * p.G * N + p.G * pitch) * sizeof(double);
    int C = (int)((budget - fixed) / per_c);
    if (C > K) C = K;
    if (C > 32) C = 32;
    if (C < 1) C = 1;
    // balance the chunks (K = 100, C = 27 -> 4 chunks of 25)
    const int nc = (K + C - 1) / C;
    C = (K + nc - 1) / nc;
    p.C = C;
    p.smem = per_c * C + fixed;
    if (const char* pad = getenv("QRL_DEBUG_SMEM_PAD")) p.smem += (size_t)atoi(pad);
    return p;
}

template <int N>
static int launch_tile(const qrl_em_args& a, cudaStream_t st, int sms) {
    const int pitch = a.packed ? (int)(a.packed_row_pitch > 0 ? a.packed_row_pitch : a.n_sel) : 0;
    const TilePlan p = plan_tiles(a.n_envs, N, a.K, a.packed ? a.n_sel : 0, pitch, sms);
    static bool attr_set = false;
    if (!attr_set) {
        QRL_CUDA(cudaFuncSetAttribute(em_tile_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    em_tile_kernel<N><<<p.grid, EM_TILE_THREADS, p.smem, st>>>(a, make_philox_keys(a.seed), p.G, p.C);
    return after_launch("em_tile_kernel");
}

// entry used by the qrl_em_step dispatcher (em_kernel.cu); returns QRL_E_UNSUPPORTED when the case is not tiled
int em_tile_launch(const qrl_em_args& a, cudaStream_t st) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        QRL_CUDA(cudaGetDevice(&dev));
        QRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
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
