/*
 * quantrl_b200 — C ABI of the B200-native hot path of QuantRL (Sampreet/quantrl v0.0.8).
 *
 * The reference is pure Python and has no FFI; its plugin seams are the Python ABCs
 * `BaseBackend` / `BaseIVPSolver` / `BaseEnv` and two registries.  This header is the boundary *beneath* those seams:
 * each entry point names the reference code it replaces (paths relative to the reference root).  The Python host
 * (the quantrl_b200 Python package) binds it with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer to contiguous FP64 unless the name
 *     starts with `h_` (host pointer).  Buffers are caller-owned; the library allocates nothing that outlives a call
 *     except a small per-device scratch (qrl_release() frees it).
 *   - every launch is an asynchronous enqueue on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) on the CURRENT device; no host synchronisation unless stated.
 *   - return value: 0 = ok, < 0 = invalid argument (QRL_E_*), > 0 = cudaError_t.  qrl_last_error() returns a
 *     thread-local message.  Nothing throws, nothing exits.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns a cudaError_t.
 *   - "strides" are in units of doubles; a stride of 0 broadcasts that axis.
 *
 * RNG stream (in-kernel noise, `rng_mode == QRL_RNG_PHILOX`), restated by oracle/em_oracle.py:philox_normals
 *   (x0,x1,x2,x3) = Philox4x32-10(counter = (tau, episode, env_global, comp), key = (seed lo32, seed hi32))
 *   u1 = 2 - bits(0x3FF<<52 | x1<<20 | x0>>12 | 1) in (0,1);  r = sqrt(-2 ln u1)
 *   j = x3>>24 (8 bits);  frac = ((x3 & 0xFFFFFF)<<28 | x2>>4) / 2^52;  theta = (j + frac) * 2 pi / 256;
 *   z0 = r sin theta;  z1 = r cos theta     (Box-Muller; ln and sin/cos are evaluated from 256-entry tables + short
 *   polynomials, sqrt by one third-order correction of the hardware seed: |dz| <= 5 * 2^-53 * r against libm)
 *   Stochastic path (qrl_em_step): the call with tau = global sub-step index k gives z0 -> Wiener increment of sub-step
 *   k -> k+1 (scaled by sqrt(t_ssz)) and z1 -> measurement noise of time index k+1 (scaled by obs_std), so every
 *   trajectory row after the first depends on exactly one call; the measurement noise of time index 0 (the reset
 *   observation) is z1 of the call with tau = 0xFFFFFFFF.
 *   Deterministic path (qrl_ode_step): measurement noise of time index t, component d = z[d & 1] of the call with
 *   counter (t, episode, env_global, 0x10000 + d/2).
 *   `env_global = env_offset + local env index` makes results independent of how environments are sharded over GPUs.
 */
#ifndef QUANTRL_B200_H
#define QUANTRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QRL_ABI_VERSION 5

#if defined(__GNUC__)
#define QRL_API __attribute__((visibility("default")))
#else
#define QRL_API
#endif

#define QRL_E_INVALID   (-1)  /* bad argument (message in qrl_last_error) */
#define QRL_E_UNSUPPORTED (-2) /* size/system not compiled in */

/* bits OR-ed (atomically) into the `nan_flag` status word of qrl_em_args / qrl_ode_args by the kernels; the host reads
 * the word wherever it synchronises anyway and must treat QRL_STATUS_PEER_TIMEOUT as fatal */
#define QRL_STATUS_NONFINITE    1  /* a state became NaN / Inf (the reference carries on silently: check_truncation is
                                    * NaN-blind, quantrl/envs/base.py:495-499) */
#define QRL_STATUS_PEER_TIMEOUT 2  /* the fused cross-GPU barrier gave up after ~2 s: a peer never arrived, the gathered
                                    * step results are stale or partial */

#define QRL_RNG_FED    0
#define QRL_RNG_PHILOX 1

/* reward functors evaluated in the kernel epilogue (replace a user `get_reward` hook, quantrl/envs/base.py:307-317) */
#define QRL_REWARD_NONE        0  /* leave `reward` untouched: the Python `get_reward()` hook fills it */
#define QRL_REWARD_NEG_WSQ_OBS 1  /* r = -sum_c w_c * Observations_c^2 */
#define QRL_REWARD_NEG_WSQ_STATE 2 /* r = -sum_c w_c * States_c^2 */
/* deterministic path (qrl_ode_step, num_quads == 4): analytic logarithmic negativity of modes (0,1), restating
 * QCMSolver.get_entanglement_logarithmic_negativity_2 (quantrl/solvers/measure.py:401-440) */
#define QRL_REWARD_ENTAN_LN_OBS   3 /* evaluated on the covariance block of the Observations */
#define QRL_REWARD_ENTAN_LN_STATE 4 /* evaluated on the covariance block of the States */
/* complete quantum synchronization of modes (0,1), QCMSolver.get_synchronization_complete (measure.py:442-468) */
#define QRL_REWARD_SYNC_C_OBS     5
#define QRL_REWARD_SYNC_C_STATE   6
/* quantum phase synchronization of modes (0,1), QCMSolver.get_synchronization_phase (measure.py:470-514); uses the
 * mode amplitudes too (num_modes >= 2) */
#define QRL_REWARD_SYNC_P_OBS     7
#define QRL_REWARD_SYNC_P_STATE   8

/* deterministic system functors (replace user hooks get_mode_rates/get_A/get_D, quantrl/envs/deterministic.py:748-818);
 * formulas in oracle/systems.py */
#define QRL_SYS_OPTOMECH    1  /* m=2, q=4, 7 params/env, 1 action */
#define QRL_SYS_KERR_CHAIN  2  /* m=N, q=2N (N = 1..4), 5+N params/env, 1 action */
#define QRL_SYS_CONST_AD    3  /* m=0; A (E,q,q), D (E,q,q) constant over the interval (is_A_constant/is_D_constant) */

#define QRL_ODE_RK4    0     /* classic RK4, fixed step t_ssz / substeps */
#define QRL_ODE_DOPRI5 1     /* Dormand–Prince 5(4), free-running steps + dense output on the grid */
/* generic explicit Runge–Kutta engine (csrc/ode_erk.cuh): steps are clipped to the grid points; one step-size
 * controller per environment for the adaptive pairs; EULER / MIDPOINT use the fixed step t_ssz / substeps.  These are the
 * explicit-RK method names of the reference's solver plugins (torchdiffeq: quantrl/solvers/torch.py:28, diffrax:
 * quantrl/solvers/jax.py:29, SciPy: quantrl/solvers/numpy.py:31) */
#define QRL_ODE_BOSH3     2  /* Bogacki–Shampine 3(2)   ('bosh3', SciPy 'RK23') */
#define QRL_ODE_TSIT5     3  /* Tsitouras 5(4)          ('tsit5') */
#define QRL_ODE_HEUN21    4  /* Heun 2(1)               ('adaptive_heun', 'heun') */
#define QRL_ODE_FEHLBERG2 5  /* Fehlberg 2(1)           ('fehlberg2') */
#define QRL_ODE_DOP853    6  /* Dormand–Prince 8(5,3)   ('dop853' / 'DOP853'; stands in for 'dopri8'); d <= 64 */
#define QRL_ODE_EULER     7  /* explicit Euler          ('euler') */
#define QRL_ODE_MIDPOINT  8  /* explicit midpoint rule  ('midpoint') */

QRL_API int         qrl_abi_version(void);
QRL_API const char* qrl_last_error(void);
/* number of CUDA devices visible (0 if none / driver missing); never fails */
QRL_API int         qrl_device_count(void);
/* free the per-device scratch of the current device */
QRL_API int         qrl_release(void);
/* number of kernel launches this library has enqueued in this process (all entry points) */
QRL_API uint64_t    qrl_launch_count(void);
/* Device-side address of page-locked HOST memory (cudaHostAlloc / cudaHostRegister, e.g. a torch pinned tensor) so that
 * it can be passed where this header asks for a device pointer: qrl_em_step then reads the actions from, and writes
 * obs_last / reward_last / minmax_out to, host memory directly over PCIe (no staging copies in the step, replaces
 * the H2D of BaseSB3Env.step_async and the D2H of step_wait, quantrl/envs/base.py:1245,1357).
 * Returns QRL_E_INVALID if the memory is pageable or not mapped for the current device. */
QRL_API int         qrl_host_device_pointer(const void* h_ptr, void** device_ptr);

/* ------------------------------------------------------------------------------------------------------------
 * Euler–Maruyama action interval for E independent linear SDE environments.
 * Replaces, fused into ONE launch: LinearEnv._update_states + func (quantrl/envs/stochastic.py:178-242; the Python
 * `for` of backends/numpy.py:178-186), the Wiener draw of LinearEnv.reset (stochastic.py:157-170, philox mode), the
 * measurement-noise draw and add of BaseEnv.reset/update (envs/base.py:408-426,462), the reward hook on all K+1 rows
 * (base.py:471-473), `rewards += Reward[dim-1]` (base.py:1311) and the min/max reductions of check_truncation
 * (base.py:485-499).
 *
 *   Y[0]   = x0
 *   Y[i+1] = (I + A_i * t_ssz) @ Y[i] + nz_i * W_i            i = 0..K-1
 *   Obs[i] = Y[i] + noise_i                                   i = 0..K
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct qrl_em_args {
    int32_t  n;            /* state dimension (n_observations); 1..16 */
    int32_t  n_envs;       /* E (local to this device) */
    int32_t  K;            /* sub-steps in this interval = len(T_step) - 1 */
    int32_t  rng_mode;     /* QRL_RNG_FED | QRL_RNG_PHILOX */
    double   t_ssz;        /* step size */

    const double* A;       /* drift A(t_idx+i, args): element [i,e,r,c] at A[i*A_stride_k + e*A_stride_e + r*n + c] */
    int64_t  A_stride_k;   /* 0: constant over the interval */
    int64_t  A_stride_e;   /* 0: shared by all envs, else n*n */
    const double* nz;      /* noise prefixes: element [i,e,c] at nz[i*nz_stride_k + e*nz_stride_e + c] */
    int64_t  nz_stride_k;
    int64_t  nz_stride_e;

    /* FED: the reference's own draws */
    const double* W;         /* (K,E,n)   Ws[t_idx : t_idx+K], already scaled by sqrt(t_ssz) */
    const double* obs_noise; /* (K+1,E,n) Observation_noises[t_idx : t_idx+K+1], already scaled; NULL = no noise */
    /* PHILOX */
    uint64_t seed;
    uint32_t episode;
    uint32_t _pad0;
    int64_t  t_idx;          /* global time index of row 0 */
    int64_t  env_offset;     /* global index of local env 0 */
    const double* obs_std;   /* element [e,c] at obs_std[e*obs_std_stride_e + c]; NULL = no measurement noise */
    int64_t  obs_std_stride_e;

    const double* x0;        /* (E,n)  States[-1] of the previous interval */
    double*  states;         /* (K+1,E,n) or NULL (trajectory off) */
    double*  observations;   /* (K+1,E,n) or NULL */
    double*  x_last;         /* (E,n) = Y[K]; may alias x0; or NULL */
    double*  obs_last;       /* (E,n) = Obs[K]; or NULL */

    int32_t  reward_kind;    /* QRL_REWARD_* */
    int32_t  flags;          /* QRL_EM_* bits */
    const double* reward_w;  /* (n,) weights */
    double*  reward;         /* (K+1,E) or NULL */
    double*  reward_last;    /* (E,) = Reward[K]; or NULL */
    double*  rewards_cum;    /* (E,) in/out: += Reward[K]; or NULL */

    double*  obs_minmax;     /* [min, max] running reduction (caller initialises to [+inf, -inf]); or NULL */
    int32_t* nan_flag;       /* status word: QRL_STATUS_* bits are OR-ed in (never cleared by the library); or NULL */

    /* fused trajectory-cache rows (BaseSB3Env.update_data, quantrl/envs/base.py:1314-1372), n_properties == 0 only:
     *   packed[e, t, j] = column sel_idxs[j] of [T_step[t], actions[e,:], Obs[t,e,:], rewards_cum[e] (after this step)]
     * requires reward_kind != QRL_REWARD_NONE and rewards_cum != NULL.  NULL = no packing. */
    double*  packed;         /* element [e,t,j] at packed[e*packed_stride_e + t*packed_row_pitch + j], t = 0..K */
    int64_t  packed_stride_e;
    int64_t  packed_row_pitch; /* doubles per cache row, >= n_sel (0 = n_sel).  An even pitch and stride keep every env's
                                * run 16-byte aligned, which lets the tiled kernel write it with TMA bulk stores */
    const double*  T_step;   /* (K+1,) */
    const double*  actions;  /* (E,n_actions) */
    const int32_t* sel_idxs; /* (n_sel,) indices into the n_data = 1 + n_actions + n + 1 columns (negative = from end) */
    int32_t  n_actions;
    int32_t  n_sel;

    /* CUDA-graph replay: optional DEVICE pointer to {int64 t_idx, int64 episode}.  When non-NULL the kernel reads the
     * two scalars from there instead of `t_idx` / `episode`, and treats W, obs_noise, T_step and packed as the bases
     * of the WHOLE episode (time index 0): it adds t_idx rows itself.  The launch is then identical from step to step
     * and can be captured once and replayed (the caller advances dyn[0] by K on the same stream). */
    const int64_t* dyn;

    /* action-affine drift functor (the shipped "affine" system family): when A_act != NULL
     *   A_e = A[e] + sum_a (actions[e,a] * action_scale[a]) * A_act[a]        (A keeps its strides; A_stride_k must be 0)
     * is evaluated inside the launch (no coefficient tensor, no separate scaling kernel) and the actions column of the
     * packed rows holds the scaled action.  action_scale NULL = 1. */
    const double* A_act;        /* (n_actions, n, n) */
    const double* action_scale; /* (n_actions,) = action_maximums, or NULL */

    /* per-launch bookkeeping done by the LAST thread block to finish (ABI 2) — lets a captured step consist of this one
     * kernel, without a reset copy for the min/max pair and without a second kernel that advances dyn[0]:
     *   finish_counter  DEVICE scratch, one uint32 that is 0 before the launch; every block increments it when it is
     *                   done, the last one runs the two actions below and leaves it 0 again.  NULL = off.
     *   minmax_out      when non-NULL: [min, max] accumulated in obs_minmax by THIS launch is copied to minmax_out[0..1]
     *                   and obs_minmax is re-initialised to [+inf, -inf] (so obs_minmax is private scratch that the
     *                   caller initialises once, and minmax_out needs no initialisation at all).
     *   dyn_advance     when != 0 and dyn != NULL: added to dyn[0] after every block has read it. */
    uint32_t* finish_counter;
    double*   minmax_out;
    int64_t   dyn_advance;

    /* cross-GPU barrier fused into the end of the launch (ABI 3; needs finish_counter) — the only exchange step of the
     * sharded path: after every block of this rank has stored its slice of obs_last / reward_last into the learner
     * rank's peer-mapped buffer (NVLink), the last block publishes a new epoch number in every rank's flag array and
     * waits until all ranks have published theirs, so that kernel completion == "every rank's step result is visible
     * on every rank".  Replaces a separate barrier launch per step.
     *   peer_flags   DEVICE array of peer_world pointers; peer_flags[p] = rank p's flag array (peer_world uint64,
     *                zero-initialised, peer-mapped symmetric memory); NULL = off
     *   peer_epoch   DEVICE uint64 of this rank, zero-initialised; incremented once per launch
     * A rank that waits longer than ~2 s sets QRL_STATUS_PEER_TIMEOUT in nan_flag (if given) and returns instead of
     * hanging the device. */
    uint64_t* const* peer_flags;
    uint64_t* peer_epoch;
    int32_t   peer_rank;
    int32_t   peer_world;

    /* publish-only exchange (ABI 4, flags & QRL_EM_PEER_PUBLISH) — what the sharded path uses: the step results stay in
     * THIS rank's memory (obs_last / reward_last point into a ring of slices that the host rotates launch by launch),
     * every block fences at device scope only, and the last block releases at system scope and stores
     * peer_publish_value into peer_flags[peer_dst][peer_rank] — one 8-byte NVLink store per launch; nobody waits.  The
     * learner rank gathers the slices with qrl_peer_pull, which polls those flags.  Flow control: when peer_pulled is
     * non-NULL the launch spins at its start until *peer_pulled >= peer_pulled_min (the learner has gathered the epoch
     * that last used the ring slice this launch overwrites); ~2 s time-out -> QRL_STATUS_PEER_TIMEOUT. */
    uint64_t  peer_publish_value;
    const uint64_t* peer_pulled;
    uint64_t  peer_pulled_min;
    int32_t   peer_dst;
    int32_t   _pad1;
    /* optional second copy of the last rows (e.g. obs_last -> pinned host memory for step(), obs_last_b -> the ring) */
    double*   obs_last_b;
    double*   reward_last_b;
    /* completion word (ABI 5; needs finish_counter) — for a host that waits for THIS step's results and nothing else
     * (BaseSB3Env.step_wait, quantrl/envs/base.py:1250-1293): when done_flag is non-NULL every block fences at SYSTEM
     * scope before it counts itself done, and the last block, after the bookkeeping above, stores done_value to
     * *done_flag with a system-scope release.  done_flag is normally the device-side address of a pinned host word
     * (qrl_host_device_pointer) next to obs_last / reward_last / minmax_out in pinned memory: a host that reads
     * done_value there may read those results — without waiting for the grid to drain and the stream to report
     * completion (qrl_em_step_host does exactly this). */
    uint64_t* done_flag;
    uint64_t  done_value;
} qrl_em_args;

#define QRL_EM_FORCE_GENERIC 1   /* flags: always use the generic lane-per-component kernel (A/B testing) */
#define QRL_EM_FORCE_TILED   16  /* flags: use the tiled kernel whenever it supports the case, whatever the batch size */
#define QRL_EM_ITEM_OUTPUT  32  /* flags: tiled kernel uses its one-(row, env)-item-per-lane output pass even when the
                                 * 16-byte-piece pass applies (A/B testing; n not in {2,4,8} always takes it) */
#define QRL_EM_PDL          256 /* flags: launch the tiled kernel with programmatic stream serialization: it may be
                                 * scheduled while the previous kernel of the stream drains (it waits for that kernel's
                                 * completion before reading anything), which hides the launch latency between the
                                 * back-to-back step kernels of a device-resident loop */
#define QRL_EM_PACK_TMA     128 /* flags: reserved (accepted and ignored; the cache rows are written from registers) */
#define QRL_EM_PEER_PUBLISH 8192 /* flags: publish-only exchange (see peer_publish_value); needs peer_flags + finish_counter */
#define QRL_EM_HOST_TIDX    64  /* flags: `dyn` is only ADVANCED by this launch (dyn_advance), not read: t_idx, episode and the
                                 * pointers that depend on t_idx (W, obs_noise, T_step, packed) were set by the host for
                                 * this launch — a host that re-enqueues every step anyway saves the kernel one global round
                                 * trip before its first Philox counter */
/* flag bits 2, 4, 8 are diagnostics of the tiled kernel (skip produce / recurrence / output stage; results invalid) and
 * bits 512, 1024, 2048, 4096 of the peer barrier (gpu-scope fences only / signal without waiting: results not
 * synchronised; extra system fence after the flag stores / volatile polling + trailing system fence: A/B variants) */

QRL_API int qrl_em_step(const qrl_em_args* args, void* stream);

/* ABI 5: qrl_em_step + the host-side wait for the step's results in ONE call — the body of BaseSB3Env.step_wait
 * (quantrl/envs/base.py:1250-1293) between "actions are in pinned memory" and "observations / reward are in pinned
 * memory".  done_host NULL: enqueue, then cudaStreamSynchronize(stream).  done_host non-NULL (the HOST address of the
 * pinned word args->done_flag points to): enqueue, then spin on *done_host until it equals args->done_value; the launch
 * may still be draining when the call returns (later work on the stream is ordered behind it as usual).  After
 * timeout_ms (0 = 2000) without the word, falls back to cudaStreamSynchronize and reports what that returns; a word still
 * missing after a successful synchronisation is QRL_E_INVALID (done_flag / done_host do not name the same memory). */
QRL_API int qrl_em_step_host(const qrl_em_args* args, void* stream, const volatile uint64_t* done_host, int32_t timeout_ms);

/* ------------------------------------------------------------------------------------------------------------
 * Deterministic action interval: modes + Lyapunov covariance  dV/dt = A V + V A^T + D  for E environments.
 * Replaces LinearizedHOVecEnv._update_states -> BaseIVPSolver.step -> integrate (quantrl/envs/deterministic.py:
 * 646-651, quantrl/solvers/base.py:168-204, quantrl/solvers/numpy.py:87-178) together with the RHS `func` and the
 * user hooks it calls (deterministic.py:653-746, 820-867), plus the same epilogue as qrl_em_step.
 *   y = [Re a (m) | Im a (m) | vec(V) (q*q)],  d = 2m + q*q
 * RK4: `substeps` classic RK4 steps per grid interval.  DOPRI5: per-env adaptive Dormand–Prince 5(4) with dense
 * output on the grid, controller in oracle/lyap_oracle.py:dopri5_interval.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct qrl_ode_args {
    int32_t  system;         /* QRL_SYS_* */
    int32_t  num_modes;      /* m */
    int32_t  num_quads;      /* q */
    int32_t  n_envs;         /* E */
    int32_t  K;              /* grid intervals = len(T_step) - 1 */
    int32_t  method;         /* QRL_ODE_* */
    int32_t  substeps;       /* fixed-step methods (RK4, EULER, MIDPOINT): steps per grid interval (>= 1) */
    int32_t  n_actions;
    double   t0;             /* T_step[0] */
    double   t_ssz;          /* grid spacing */
    double   atol, rtol;     /* adaptive methods */

    const double* params;    /* (E, n_params) system parameters, row stride params_stride_e (0 = shared) */
    int64_t  params_stride_e;
    const double* actions;   /* (E, n_actions) already multiplied by action_maximums */
    const double* A;         /* QRL_SYS_CONST_AD: (E,q,q), stride A_stride_e */
    int64_t  A_stride_e;
    const double* D;
    int64_t  D_stride_e;

    const double* y0;        /* (E,d) */
    double*  states;         /* (K+1,E,d) or NULL */
    double*  observations;   /* (K+1,E,d) or NULL */
    double*  y_last;         /* (E,d); may alias y0 */
    double*  obs_last;       /* (E,d) or NULL */
    const double* obs_noise; /* FED measurement noise (K+1,E,d), already scaled; or NULL */
    /* PHILOX measurement noise (used when obs_noise is NULL and obs_std is not):
     * counter = (t_idx + row, episode, env_global, 0x10000 + d/2), component d takes z[d & 1] */
    uint64_t seed;
    uint32_t episode;
    uint32_t _pad0;
    int64_t  t_idx;
    int64_t  env_offset;
    const double* obs_std;   /* element [e,d] at obs_std[e*obs_std_stride_e + d]; or NULL */
    int64_t  obs_std_stride_e;
    double*  h_state;        /* adaptive methods: (E,) in/out carried step size (<= 0: start from t_ssz); or NULL */
    int32_t* n_steps;        /* adaptive methods: (E,2) accepted, rejected counters (accumulated); or NULL */

    double*  obs_minmax;     /* as in qrl_em_args */
    int32_t* nan_flag;
    int32_t  flags;          /* QRL_ODE_* bits */
    int32_t  reward_kind;    /* QRL_REWARD_NONE | QRL_REWARD_ENTAN_LN_* | QRL_REWARD_SYNC_C_* | QRL_REWARD_SYNC_P_* */
    double*  reward;         /* (K+1,E) or NULL */
    double*  reward_last;    /* (E,) = Reward[K]; or NULL */
    double*  rewards_cum;    /* (E,) in/out: += Reward[K]; or NULL */
    /* ABI 4: when non-NULL, `actions` holds the RAW actions (e.g. read in place from pinned host memory) and the functors
     * see actions[e,k] * action_scale[k]  (= float(action) * action_maximums, quantrl/envs/base.py:1244-1248) */
    const double* action_scale; /* (n_actions,) or NULL */
} qrl_ode_args;

/* Kernel selection of qrl_ode_step at num_quads = 4 (all variants perform the same operations; measured crossovers on
 * B200 in DESIGN.md 4.2):  RK4: n_envs <= 1536 split kernel, <= 3072 single-role lanes kernel, else one thread per env;
 * DOPRI5 and the other explicit-RK methods: n_envs <= 12288 lanes kernel, else one thread per env.  Other num_quads:
 * one thread per env. */
#define QRL_ODE_FORCE_THREAD 1  /* flags: RK4 / DOPRI5 with one thread per env whatever the batch size (A/B testing) */
#define QRL_ODE_FORCE_LANES  2  /* flags: RK4 / DOPRI5 with q*q lanes per env (q = 4 only) whatever the batch size; RK4: the
                                * warp-specialised split kernel (mode producer / covariance lanes / row writers) */
#define QRL_ODE_LANES_SINGLE 4 /* flags: with FORCE_LANES (or a small batch), the single-role lanes kernel (A/B testing) */
/* flags bits 8 / 16 / 32: diagnostics of the split kernel only (skip the producer / consumer / writer role; results are
 * then meaningless) — used by tools/prof_ode.py to time the roles alone */

QRL_API int qrl_ode_step(const qrl_ode_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Packed trajectory-cache rows, BaseSB3Env.update_data (quantrl/envs/base.py:1314-1356):
 *   packed[e, t, :] = [T_step[t], actions[e,:], Obs[t,e,:], Props[t,e,:], rewards_cum[e]]      (E, dim, n_data)
 * and the host cache columns `data[:, t0:t1, :] = _data[:, :, data_idxs]` (base.py:1370-1372):
 *   selected[e, t, j] = packed[e, t, idxs[j]]                                                  (E, dim, n_sel)
 * Either output may be NULL.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct qrl_pack_args {
    int32_t  n_envs, dim, n_actions, n_obs, n_props, n_sel;
    const double*  T_step;      /* (dim,) */
    const double*  actions;     /* (E,n_actions) */
    const double*  observations;/* (dim,E,n_obs) */
    const double*  properties;  /* (dim,E,n_props) or NULL */
    const double*  rewards_cum; /* (E,) */
    const int32_t* idxs;        /* (n_sel,) column indices into the n_data columns (negative = from the end) */
    double*  packed;            /* element [e,t,c] at packed[e*packed_stride_e + t*n_data + c]; or NULL */
    int64_t  packed_stride_e;   /* >= dim*n_data (lets the caller write straight into a (E, shape_T, n_data) cache) */
    double*  selected;          /* element [e,t,j] at selected[e*selected_stride_e + t*selected_row_pitch + j]; or NULL */
    int64_t  selected_stride_e; /* >= dim*selected_row_pitch */
    int64_t  selected_row_pitch;/* doubles per row, >= n_sel (0 = n_sel) */
    const int64_t* dyn;         /* optional DEVICE pointer {int64 t_idx, ...} (CUDA-graph replay): when non-NULL, T_step,
                                 * packed and selected are the bases of the whole episode and t_idx rows are added here */
    const double*  action_scale;/* (n_actions,) multiplied into the actions column; NULL = actions are already scaled */
} qrl_pack_args;

QRL_API int qrl_pack_rows(const qrl_pack_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Learner-side gather of one action step of the env-sharded path (the only exchange step: SURVEY.md §8e; replaces an
 * NCCL all_gather of [obs | reward] per step).  Rank p keeps the last rows of epoch e in slice e % out_slots of its own
 * ring (peer-mapped symmetric memory) and its step kernel stores e into flags[p] of the learner (QRL_EM_PEER_PUBLISH).
 * This launch — on the learner, typically on a side stream so that it overlaps the next step — gathers epoch
 * e = *pull_epoch + 1: the blocks that copy rank p's slice first poll flags[p] >= e, so slices of early ranks move
 * while late ranks still integrate; the last block then advances *pull_epoch and stores e into every rank's `pulled`
 * word (the credit that lets rank p reuse the ring slice).  Ring slice of rank p (count_p environments):
 *   [ obs (count_p, n_obs) | reward (count_p,) ]   at peer_ring[p] + (e % out_slots) * slot_stride[p]   (doubles)
 * A rank that does not publish within ~2 s sets QRL_STATUS_PEER_TIMEOUT in *status and the launch returns.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct qrl_peer_pull_args {
    int32_t  world;               /* number of ranks */
    int32_t  n_obs;
    int32_t  out_slots;           /* slices per ring */
    int32_t  blocks_per_rank;     /* 0 = default */
    const uint64_t* flags;        /* (world,) this rank's flag array, written by the step kernels of all ranks */
    const double* const* peer_ring; /* DEVICE table (world,): rank p's ring, mapped into this process */
    const int64_t* slot_stride;   /* DEVICE (world,) doubles between the slices of rank p's ring */
    const int64_t* env_offset;    /* DEVICE (world,) global index of rank p's first environment */
    const int64_t* env_count;     /* DEVICE (world,) environments of rank p */
    double*  obs_all;             /* (E_total, n_obs) gathered observations */
    double*  reward_all;          /* (E_total,) gathered rewards */
    uint64_t* pull_epoch;         /* DEVICE counter of this rank, zero-initialised */
    uint64_t* const* peer_pulled; /* DEVICE table (world,): rank p's `pulled` word, mapped into this process */
    uint32_t* finish_counter;     /* DEVICE scratch, 0 before the launch and after it */
    int32_t*  status;             /* QRL_STATUS_* bits; or NULL */
} qrl_peer_pull_args;

QRL_API int qrl_peer_pull(const qrl_peer_pull_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Copy-engine variant of the same exchange (the default of the sharded path): nothing changes in the step kernel, no SM
 * is involved.  After the step launch of epoch e on `main_stream` a rank calls qrl_peer_push, which enqueues
 *   main_stream:  record(step_done)
 *   side_stream:  wait(step_done); copy ring slice -> the learner's gathered block (peer-mapped, NVLink, copy engine);
 *                 copy the 8-byte epoch word -> learner.flags[rank]; record(copy_done)
 * so the transfer of step e overlaps step e+1 and the step stream never waits.  The epoch word follows the data on the
 * same stream, so a flag value >= e means the slice of epoch e has landed.  The caller owns streams and events
 * (cudaStream_t / cudaEvent_t passed as void*); before it reuses a ring slice it passes that slice's copy_done event to
 * qrl_peer_event_wait (returns at once when the copy has finished — it always has, the ring being several steps deep).
 * The learner waits for an epoch with qrl_peer_wait: one warp that polls flags[0..world) >= epoch on `stream`
 * (~2 s time-out -> QRL_STATUS_PEER_TIMEOUT in *status).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct qrl_peer_push_args {
    const void* src;          /* this rank's ring slice of the epoch (device): observations */
    void*       dst;          /* this rank's slice of the learner's gathered block (peer-mapped device pointer) */
    int64_t     bytes;
    const void* src2;         /* second piece (rewards), or NULL */
    void*       dst2;
    int64_t     bytes2;
    const void* flag_src;     /* 8-byte epoch word: pinned HOST memory written by the caller, or device memory */
    void*       flag_dst;     /* learner.flags[rank] (peer-mapped device pointer) */
    void*       step_done;    /* cudaEvent_t */
    void*       copy_done;    /* cudaEvent_t */
} qrl_peer_push_args;

QRL_API int qrl_peer_push(const qrl_peer_push_args* args, void* main_stream, void* side_stream);
/* The same with the driver calls off the caller's thread: a context owns a worker thread (bound to the current device) and
 * `side_stream`.  qrl_peer_ctx_push records step_done on `main_stream` itself (that must happen in stream order, right
 * behind the step launch: one driver call) and queues the rest — stream wait, copies, copy_done record — for the worker;
 * it returns a sequence number.  qrl_peer_ctx_wait(ctx, seq) blocks until job `seq` has been issued AND its copy_done event
 * has completed (seq 0: nothing to wait for).  A step loop that is device bound at ~27 us per step stays device bound:
 * the five driver calls of qrl_peer_push cost the calling thread ~12 us, this path ~3 us. */
QRL_API int      qrl_peer_ctx_create(void** ctx, void* side_stream);
QRL_API uint64_t qrl_peer_ctx_push(void* ctx, const qrl_peer_push_args* args, void* main_stream);   /* 0 = error */
QRL_API int      qrl_peer_ctx_wait(void* ctx, uint64_t seq);
QRL_API int      qrl_peer_ctx_destroy(void* ctx);
QRL_API int qrl_peer_event_wait(void* event);   /* cudaEventSynchronize */
QRL_API int qrl_peer_wait(const uint64_t* flags, int32_t world, uint64_t epoch, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Machine-peak micro-benchmarks used as roofline denominators (synchronous; return the measurement).
 * ------------------------------------------------------------------------------------------------------------ */
/* dependent-free DFMA stream on every SM: returns achieved FP64 TFLOP/s (FMA = 2 flop) in *tflops */
QRL_API int qrl_measure_fp64_peak(int iters, double* tflops);
/* one warp, one dependent DFMA chain: returns the FP64 FMA pipeline latency in SM cycles */
QRL_API int qrl_measure_fp64_latency(double* cycles_per_dfma);

#ifdef __cplusplus
}
#endif
#endif /* QUANTRL_B200_H */
