// Tiled, warp-specialised Euler–Maruyama action interval — the fast path of qrl_em_step on sm_100a.
//
// Why: with one lane per state component (em_kernel.cu) config 3 (E = 4096) puts a single warp on each SM
// sub-partition and the kernel is latency bound (ncu r1a: IPC 0.25, FP64 pipe 12 %).  The noise generation
// (Philox4x32-10 + FP64 Box–Muller, ~95 % of the instructions) does not depend on the state, so it is spread over
// 15 "worker" warps per CTA, while ONE warp runs the short sequential recurrence for all environments of the CTA.
//
// One CTA owns G <= 32 environments for the whole interval; time is cut into chunks of C sub-steps that flow through
// double-buffered shared-memory tiles in a 3-stage pipeline, one __syncthreads per chunk:
//   iteration it:  workers   produce noise(chunk it)        -> Wt/Vt[it & 1]       (Philox or fed loads)
//                  warp 0    recurrence(chunk it-1)          Wt/Vt[(it-1) & 1] -> Xt/Ot[(it-1) & 1]
//                  workers   output(chunk it-2)              Xt/Ot[it & 1] -> States, Observations, Reward,
//                                                            packed cache rows, min/max  (coalesced 16-byte stores)
// Replaces the same reference code as em_kernel.cu (quantrl/envs/stochastic.py:178-242, envs/base.py:408-426,462,
// 471-473,485-499) plus the packed rows of BaseSB3Env.update_data (envs/base.py:1314-1372).
#include "common.cuh"
#include "philox.cuh"

namespace qrl {

constexpr int EM_TILE_THREADS = 512;
constexpr int EM_TILE_WORKERS = EM_TILE_THREADS - 32;

template <int N>
__global__ void __launch_bounds__(EM_TILE_THREADS, 1) em_tile_kernel(const qrl_em_args a, const PhiloxKeys keys, const int G, const int C) {
    extern __shared__ double sm[];
    const int E = a.n_envs, K = a.K;
    const int e0 = blockIdx.x * G;
    const int Gl = min(G, E - e0);             // live environments of this CTA
    const int RP = Gl * N;                     // doubles per tile row
    const int tile = C * G * N;                // doubles per tile buffer
    double* Wt = sm;                           // [2][tile]
    double* Vt = Wt + 2 * tile;
    double* Xt = Vt + 2 * tile;
    double* Ot = Xt + 2 * tile;
    double* stdS = Ot + 2 * tile;              // [G*N] measurement-noise std per (env, comp)
    double* cumS = stdS + G * N;               // [G]   cumulative reward after this interval
    int* selS = reinterpret_cast<int*>(cumS + G);  // [n_sel]

    const int tid = threadIdx.x;
    const bool is_rec = tid < 32;
    const int wt = tid - 32;                   // worker lane id (0..EM_TILE_WORKERS-1)
    const bool philox = a.rng_mode == QRL_RNG_PHILOX;
    const double sqrt_dt = sqrt(a.t_ssz);
    const int NC = (K + C - 1) / C;
    const int64_t row = (int64_t)E * N;
    const int n_data = 1 + a.n_actions + N + 1;

    for (int i = tid; i < RP; i += EM_TILE_THREADS) {
        const int el = i / N, c = i - el * N;
        stdS[i] = (philox && a.obs_std) ? a.obs_std[(int64_t)(e0 + el) * a.obs_std_stride_e + c] : 0.0;
    }
    if (a.packed)
        for (int i = tid; i < a.n_sel; i += EM_TILE_THREADS) {
            const int s = a.sel_idxs[i];
            selS[i] = s < 0 ? s + n_data : s;
        }
    __syncthreads();

    double lo = INFINITY, hi = -INFINITY;

    // ---- per-row outputs shared by the prologue (row 0) and the output stage -------------------------------------
    auto emit_row = [&](int r, int el, const double* x, const double* o) {
        const int e = e0 + el;
        const int64_t off = (int64_t)r * row + (int64_t)e * N;
        if (a.states) {
#pragma unroll
            for (int c = 0; c < N; ++c) a.states[off + c] = x[c];
        }
        if (a.observations) {
#pragma unroll
            for (int c = 0; c < N; ++c) a.observations[off + c] = o[c];
        }
#pragma unroll
        for (int c = 0; c < N; ++c) { lo = fmin(lo, o[c]); hi = fmax(hi, o[c]); }
        if (a.reward_kind != QRL_REWARD_NONE) {
            const double* s = (a.reward_kind == QRL_REWARD_NEG_WSQ_OBS) ? o : x;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) acc = fma(a.reward_w[c] * s[c], s[c], acc);
            const double rv = -acc;
            if (a.reward) a.reward[(int64_t)r * E + e] = rv;
            if (r == K) {
                if (a.reward_last) a.reward_last[e] = rv;
                double cum = rv;
                if (a.rewards_cum) { cum += a.rewards_cum[e]; a.rewards_cum[e] = cum; }
                cumS[el] = cum;
            }
        }
        if (r == K) {
#pragma unroll
            for (int c = 0; c < N; ++c) {
                if (a.x_last) a.x_last[(int64_t)e * N + c] = x[c];
                if (a.obs_last) a.obs_last[(int64_t)e * N + c] = o[c];
            }
        }
    };
    auto pack_row = [&](int r, int el, const double* o) {
        const int e = e0 + el;
        double* dst = a.packed + (int64_t)e * a.packed_stride_e + (int64_t)r * a.n_sel;
        for (int j = 0; j < a.n_sel; ++j) {
            const int col = selS[j];
            if (col == 0) dst[j] = a.T_step[r];
            else if (col <= a.n_actions) dst[j] = a.actions[(int64_t)e * a.n_actions + (col - 1)];
            else if (col <= a.n_actions + N) dst[j] = o[col - 1 - a.n_actions];
            // the cumulative-reward column is only known after the last sub-step: patched below
        }
    };

    // ---- recurrence warp state ---------------------------------------------------------------------------------
    double M[N][N], nz[N], x[N];
    const int el_r = tid;                      // env handled by this recurrence lane
    const bool rec_live = is_rec && el_r < Gl;
    if (rec_live) {
        const int e = e0 + el_r;
        const double* Ar = a.A + (int64_t)e * a.A_stride_e;
#pragma unroll
        for (int r = 0; r < N; ++r) {
#pragma unroll
            for (int c = 0; c < N; ++c) M[r][c] = __dadd_rn((r == c) ? 1.0 : 0.0, __dmul_rn(Ar[r * N + c], a.t_ssz));
            nz[r] = a.nz[(int64_t)e * a.nz_stride_e + r];
            x[r] = a.x0[(int64_t)e * N + r];
        }
        // row 0: Observations[0] = States[0] + noise[t_idx]  (envs/base.py:462)
        double o[N];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double v = 0.0;
            if (philox) {
                if (a.obs_std) {
                    double z0, z1;
                    const uint32_t tau = a.t_idx > 0 ? (uint32_t)(a.t_idx - 1) : 0xFFFFFFFFu;
                    philox_normal2(keys, tau, a.episode, (uint32_t)(a.env_offset + e), (uint32_t)c, z0, z1);
                    v = __dmul_rn(stdS[el_r * N + c], z1);
                }
            } else if (a.obs_noise) {
                v = a.obs_noise[(int64_t)e * N + c];
            }
            o[c] = __dadd_rn(x[c], v);
        }
        emit_row(0, el_r, x, o);
        if (a.packed) pack_row(0, el_r, o);
    }

    // incremental (slot, rem) walkers avoid runtime integer divisions in the hot loops
    const int p_dslot = EM_TILE_WORKERS / RP, p_drem = EM_TILE_WORKERS - p_dslot * RP;
    const int o_dslot = EM_TILE_WORKERS / Gl, o_drem = EM_TILE_WORKERS - o_dslot * Gl;

    for (int it = 0; it < NC + 2; ++it) {
        if (!is_rec) {
            // ---------------- produce noise of chunk `it` ----------------
            if (it < NC) {
                const int j0 = it * C;
                const int ns = min(C, K - j0);
                double* Wb = Wt + (it & 1) * tile;
                double* Vb = Vt + (it & 1) * tile;
                int slot = wt / RP, rem = wt - slot * RP;
                while (slot < ns) {
                    const int el = rem / N, c = rem - el * N;
                    const int idx = slot * RP + rem;
                    if (philox) {
                        double z0, z1;
                        philox_normal2(keys, (uint32_t)(a.t_idx + j0 + slot), a.episode, (uint32_t)(a.env_offset + e0 + el),
                                       (uint32_t)c, z0, z1);
                        Wb[idx] = __dmul_rn(sqrt_dt, z0);
                        Vb[idx] = __dmul_rn(stdS[rem], z1);
                    } else {
                        const int64_t g = (int64_t)(j0 + slot) * row + (int64_t)e0 * N + rem;
                        Wb[idx] = a.W[g];
                        Vb[idx] = a.obs_noise ? a.obs_noise[g + row] : 0.0;
                    }
                    slot += p_dslot;
                    rem += p_drem;
                    if (rem >= RP) { rem -= RP; ++slot; }
                }
            }
            // ---------------- output chunk `it - 2` ----------------
            if (it >= 2) {
                const int co = it - 2;
                const int r0 = co * C;
                const int ns = min(C, K - r0);
                const double* Xb = Xt + (co & 1) * tile;
                const double* Ob = Ot + (co & 1) * tile;
                // (a) time-major blocks: item = (slot, env), env fastest -> consecutive lanes write consecutive
                //     N*8-byte pieces of one row of the (K+1,E,N) blocks
                int slot = wt / Gl, el = wt - slot * Gl;
                while (slot < ns) {
                    double xs[N], os[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) { xs[c] = Xb[slot * RP + el * N + c]; os[c] = Ob[slot * RP + el * N + c]; }
                    emit_row(r0 + slot + 1, el, xs, os);
                    slot += o_dslot;
                    el += o_drem;
                    if (el >= Gl) { el -= Gl; ++slot; }
                }
                // (b) env-major packed cache rows: item = (env, slot), slot fastest -> consecutive lanes write
                //     consecutive rows of one env's contiguous run
                if (a.packed) {
                    for (int item = wt; item < Gl * ns; item += EM_TILE_WORKERS) {
                        const int el2 = item / ns, sl = item - el2 * ns;
                        double os[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) os[c] = Ob[sl * RP + el2 * N + c];
                        pack_row(r0 + sl + 1, el2, os);
                    }
                }
            }
        } else if (it >= 1 && it <= NC && rec_live) {
            // ---------------- recurrence of chunk `it - 1` ----------------
            const int cr = it - 1;
            const int ns = min(C, K - cr * C);
            const double* Wb = Wt + (cr & 1) * tile;
            const double* Vb = Vt + (cr & 1) * tile;
            double* Xb = Xt + (cr & 1) * tile;
            double* Ob = Ot + (cr & 1) * tile;
            for (int slot = 0; slot < ns; ++slot) {
                const int base = slot * RP + el_r * N;
                double w[N], v[N], xn[N];
#pragma unroll
                for (int c = 0; c < N; ++c) { w[c] = Wb[base + c]; v[c] = Vb[base + c]; }
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double acc = __dmul_rn(M[r][0], x[0]);
#pragma unroll
                    for (int c = 1; c < N; ++c) acc = fma(M[r][c], x[c], acc);
                    xn[r] = fma(nz[r], w[r], acc);
                }
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    x[c] = xn[c];
                    Xb[base + c] = xn[c];
                    Ob[base + c] = __dadd_rn(xn[c], v[c]);
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
        for (int j = 0; j < a.n_sel; ++j) {
            if (selS[j] != n_data - 1) continue;
            for (int item = tid; item < Gl * (K + 1); item += EM_TILE_THREADS) {
                const int el = item / (K + 1), r = item - el * (K + 1);
                a.packed[(int64_t)(e0 + el) * a.packed_stride_e + (int64_t)r * a.n_sel + j] = cumS[el];
            }
        }
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

struct TilePlan { int G, C, grid; size_t smem; };

static TilePlan plan_tiles(int E, int N, int K, int n_sel, int sms) {
    TilePlan p;
    p.G = E <= 32 * sms ? (E + sms - 1) / sms : 32;
    if (p.G < 1) p.G = 1;
    p.grid = (E + p.G - 1) / p.G;
    const size_t budget = 200 * 1024;
    const size_t fixed = (size_t)(p.G * N + p.G) * sizeof(double) + (size_t)n_sel * sizeof(int) + 64;
    int C = (int)((budget - fixed) / ((size_t)8 * p.G * N * sizeof(double)));
    if (C > K) C = K;
    if (C > 32) C = 32;
    if (C < 1) C = 1;
    // balance the chunks (K = 100, C = 32 -> 4 chunks of 25)
    const int nc = (K + C - 1) / C;
    C = (K + nc - 1) / nc;
    p.C = C;
    p.smem = (size_t)8 * C * p.G * N * sizeof(double) + fixed;
    return p;
}

template <int N>
static int launch_tile(const qrl_em_args& a, cudaStream_t st, int sms) {
    const TilePlan p = plan_tiles(a.n_envs, N, a.K, a.packed ? a.n_sel : 0, sms);
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
