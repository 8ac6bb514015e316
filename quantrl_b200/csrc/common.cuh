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

// Per-DEVICE caches (the header promises "current device" semantics: a process may drive several GPUs).
constexpr int QRL_MAX_DEVICES = 64;
inline int current_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= QRL_MAX_DEVICES) d = 0;
    return d;
}
inline int device_sm_count() {
    static int sms[QRL_MAX_DEVICES] = {0};
    const int d = current_device_slot();
    if (sms[d] == 0) cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
    return sms[d] > 0 ? sms[d] : 148;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device property of the function: set it once per device
struct OncePerDevice {
    bool done[QRL_MAX_DEVICES] = {false};
    bool first() {
        const int d = current_device_slot();
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

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
    if (in.dyn && !(in.flags & QRL_EM_HOST_TIDX)) {
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
// One row of the Euler–Maruyama update  x'_r = sum_j M[r][j] x[j] + nz_r * w_r  (quantrl/envs/stochastic.py:229-242), as
// every kernel of this library evaluates it: the sum is split into two INDEPENDENT fused-multiply-add chains over the
// first H = ceil(n/2) and the remaining columns, the noise term seeds the first chain, and the halves meet in one add —
//     lo = nz_r * w_r;  lo = fma(M[r][j], x[j], lo)  j = 0 .. H-1
//     hi = 0;           hi = fma(M[r][j], x[j], hi)  j = H .. n-1
//     x'_r = lo + hi
// The recurrence is the one sequential dependency of the path: with FP64 results 8.4 cycles apart (and ~20 under pipe
// contention, profiles/r02c) the chain per sub-step is n/2 + 1 links instead of n + 1.  All kernels use exactly this order,
// so their trajectories agree bit for bit; against the reference (NumPy matmul) the difference is rounding (rtol 1e-12).
template <int N>
__device__ __forceinline__ double em_row_update(const double (&Mr)[N], const double (&x)[N], double nz, double w) {
#ifdef QRL_EM_CHAIN_UPDATE   // experiment only: the single chain of rounds 1
    double acc = __dmul_rn(Mr[0], x[0]);
#pragma unroll
    for (int j = 1; j < N; ++j) acc = fma(Mr[j], x[j], acc);
    return fma(nz, w, acc);
#endif
    constexpr int H = (N + 1) / 2;
    double lo = __dmul_rn(nz, w), hi = 0.0;
#pragma unroll
    for (int j = 0; j < H; ++j) lo = fma(Mr[j], x[j], lo);
#pragma unroll
    for (int j = H; j < N; ++j) hi = fma(Mr[j], x[j], hi);
    return __dadd_rn(lo, hi);
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

// Atomic min / max on doubles without a CAS loop: for IEEE-754 bit patterns, non-negative values order like signed
// integers and negative values order inversely like unsigned integers, so one integer atomic (a fire-and-forget RED)
// per call is enough whatever the signs of the stored and the new value.  NaN is never passed in (fmin/fmax drop it).
__device__ __forceinline__ void atomic_min_d(double* addr, double v) {
    const long long b = __double_as_longlong(v);
    if (b >= 0) atomicMin(reinterpret_cast<long long*>(addr), b);
    else atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)b);
}
__device__ __forceinline__ void atomic_max_d(double* addr, double v) {
    const long long b = __double_as_longlong(v);
    if (b >= 0) atomicMax(reinterpret_cast<long long*>(addr), b);
    else atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)b);
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

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Flow control of the publish-only exchange (qrl_em_args.peer_pulled): one thread per block, at the start of the launch,
// before anything is written — the learner must have gathered the epoch that last used the ring slice this launch
// overwrites.  Almost always true already (the learner trails by one step, the ring has three slices).
__device__ __forceinline__ void em_wait_pulled(const qrl_em_args& a) {
    if (a.peer_pulled == nullptr || a.peer_pulled_min == 0ull) return;
    const unsigned long long t0 = global_timer_ns();
    while (true) {
        // relaxed: the credit only gates this launch's own stores (a control dependency); nothing of the learner's is read.
        // (A system-scope ACQUIRE here costs every block of every launch a fence: measured +4 us per step.)
        unsigned long long seen;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.peer_pulled) : "memory");
        if (seen >= a.peer_pulled_min) return;
        if (global_timer_ns() - t0 > 2000000000ull) {
            if (a.nan_flag) atomicOr(a.nan_flag, QRL_STATUS_PEER_TIMEOUT);
            return;
        }
        __nanosleep(100);
    }
}

// Last-block-done bookkeeping of qrl_em_args (finish_counter / minmax_out / dyn_advance / peer barrier); call once per
// block, by all threads, after block_minmax_commit (whose atomics are issued by thread 0, like the counter increment).
__device__ __forceinline__ void em_finish_launch(const qrl_em_args& a) {
    if (a.finish_counter == nullptr) return;
    __syncthreads();                                    // every thread of the block has read dyn and is done
    if (threadIdx.x != 0) return;
    const bool publish_only = (a.flags & QRL_EM_PEER_PUBLISH) && a.peer_flags;
    // this block's min/max atomics, dyn reads and (peer barrier) its stores into the learner's buffer come first.
    // Publish-only exchange: the outputs are local, a device-scope fence per block is enough — the last block's
    // system-scope release below is cumulative over what it has observed through the counter
    // Completion word (ABI 5): the rows this block stored to pinned host memory must be ordered before the word the
    // last block stores there, so the block fences at system scope as well.
    if ((a.peer_flags && !publish_only && !(a.flags & 512)) || a.done_flag) __threadfence_system(); else __threadfence();
    const unsigned prev = atomicAdd(a.finish_counter, 1u);
    if (prev != gridDim.x - 1) return;
    if (publish_only) {
        __threadfence();               // acquire side of the counter hand-over (device scope: the other blocks' rows are local)
        if (a.minmax_out && a.obs_minmax) {
            volatile double* acc = a.obs_minmax;
            a.minmax_out[0] = acc[0];
            a.minmax_out[1] = acc[1];
            acc[0] = INFINITY;
            acc[1] = -INFINITY;
        }
        if (a.dyn && a.dyn_advance != 0) const_cast<int64_t*>(a.dyn)[0] += a.dyn_advance;
        *a.finish_counter = 0u;
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.peer_flags[a.peer_dst]) + a.peer_rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"((unsigned long long)a.peer_publish_value) : "memory");
        if (a.done_flag)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.done_flag), "l"((unsigned long long)a.done_value) : "memory");
        return;
    }
    // last block.  Cross-GPU barrier, part 1: publish this rank's new epoch in every rank's flag array FIRST — the
    // release fence then has nothing of this thread's own bookkeeping to wait for, and the flags travel over NVLink
    // while the local bookkeeping below is done.  (The other blocks' stores were fenced at system scope before their
    // counter increments, which this thread has observed.)
    unsigned long long epoch = 0ull;
    if (a.peer_flags) {
        epoch = *a.peer_epoch + 1ull;
        __threadfence_system();
        for (int p = 0; p < a.peer_world; ++p)
            reinterpret_cast<volatile unsigned long long*>(a.peer_flags[p])[a.peer_rank] = epoch;
        if (a.flags & 2048) __threadfence_system();   // diagnostic (A/B): wait for the flag stores before polling
    } else {
        __threadfence();
    }
    if (a.minmax_out && a.obs_minmax) {
        volatile double* acc = a.obs_minmax;
        a.minmax_out[0] = acc[0];
        a.minmax_out[1] = acc[1];
        acc[0] = INFINITY;
        acc[1] = -INFINITY;
    }
    if (a.dyn && a.dyn_advance != 0) const_cast<int64_t*>(a.dyn)[0] += a.dyn_advance;
    *a.finish_counter = 0u;
    if (a.peer_flags) {
        *a.peer_epoch = epoch;
        // part 2: wait for everybody's epoch.  Acquire loads at system scope (what the peers' data stores were released
        // to), so no separate acquire fence after the loop; diagnostic flag 4096 = plain volatile polling + a system fence
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(a.peer_flags[a.peer_rank]);
        const bool acq = !(a.flags & 4096);
        const unsigned long long t0 = global_timer_ns();
        for (int q = 0; q < a.peer_world && !(a.flags & 1024); ++q) {
            while (true) {
                unsigned long long seen;
                if (acq) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine + q) : "memory");
                else seen = *reinterpret_cast<const volatile unsigned long long*>(mine + q);
                if (seen >= epoch) break;
                if (global_timer_ns() - t0 > 2000000000ull) {   // a peer never arrived: fail loudly, do not hang
                    if (a.nan_flag) atomicOr(a.nan_flag, QRL_STATUS_PEER_TIMEOUT);
                    q = a.peer_world;
                    break;
                }
            }
        }
        if (!acq) __threadfence_system();
    }
    if (a.done_flag)     // ABI 5: everything above (and, through the counter, every block's rows) is released with it
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.done_flag), "l"((unsigned long long)a.done_value) : "memory");
}

}  // namespace qrl
