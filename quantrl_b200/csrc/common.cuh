// Shared helpers for the quantrl_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/quantrl_b200.h"

namespace qrl {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

inline int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

#define QRL_REQUIRE(cond, msg)                                              \
    do {                                                                    \
        if (!(cond)) return ::qrl::fail(QRL_E_INVALID, "%s", msg);          \
    } while (0)

#define QRL_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) return ::qrl::cuda_fail(_e, #call);          \
    } while (0)

inline int after_launch(const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    return 0;
}

// ---- device helpers ---------------------------------------------------------------------------------------
// CUDA-graph replay support: fold the device-resident {t_idx, episode} into a private copy of the arguments
__device__ __forceinline__ qrl_em_args resolve_dyn(const qrl_em_args& in) {
    qrl_em_args a = in;
    if (in.dyn) {
        const int64_t t = in.dyn[0];
        const int64_t rows = t * (int64_t)in.n_envs * in.n;
        a.t_idx = t;
        a.episode = (uint32_t)in.dyn[1];
        if (a.W) a.W += rows;
        if (a.obs_noise) a.obs_noise += rows;
        if (a.T_step) a.T_step += t;
        if (a.packed) a.packed += t * (in.packed_row_pitch > 0 ? in.packed_row_pitch : (int64_t)in.n_sel);
    }
    return a;
}

// drift element (r,c) of environment e: coefficient tensor, or the action-affine functor A0 + sum_a u_a A_a
__device__ __forceinline__ double em_drift(const qrl_em_args& a, const double* Ae, int e, int rc) {
    double v = Ae[rc];
    if (a.A_act) {
        const int nn = a.n * a.n;
        for (int k = 0; k < a.n_actions; ++k) {
            const double u = a.actions[(int64_t)e * a.n_actions + k] * (a.action_scale ? a.action_scale[k] : 1.0);
            v = fma(u, a.A_act[k * nn + rc], v);
        }
    }
    return v;
}
__device__ __forceinline__ double em_action(const qrl_em_args& a, int e, int k) {
    const double u = a.actions[(int64_t)e * a.n_actions + k];
    return (a.A_act && a.action_scale) ? u * a.action_scale[k] : u;
}

__device__ __forceinline__ double shfl_d(double v, int src, int width = 32) {
    return __shfl_sync(0xffffffffu, v, src, width);
}
__device__ __forceinline__ double shfl_xor_d(double v, int mask, int width = 32) {
    return __shfl_xor_sync(0xffffffffu, v, mask, width);
}

__device__ __forceinline__ void atomic_min_d(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_d(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (v > __longlong_as_double((long long)old)) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// Block-wide min/max of per-thread running extrema followed by one atomic pair per block.
// NaN never wins a comparison, which matches the reference's `max(obs) > hi or min(obs) < lo` being False for NaN
// (quantrl/envs/base.py:495-499).
__device__ __forceinline__ void block_minmax_commit(double lo, double hi, double* obs_minmax) {
    __shared__ double s_lo[32], s_hi[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, shfl_xor_d(lo, o));
        hi = fmax(hi, shfl_xor_d(hi, o));
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { s_lo[w] = lo; s_hi[w] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < nw; ++i) { lo = fmin(lo, s_lo[i]); hi = fmax(hi, s_hi[i]); }
        atomic_min_d(obs_minmax + 0, lo);
        atomic_max_d(obs_minmax + 1, hi);
    }
}

}  // namespace qrl
