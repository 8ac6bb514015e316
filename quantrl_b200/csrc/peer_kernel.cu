// Learner-side gather of the env-sharded path (qrl_peer_pull) — sm_100a, NVLink peer memory.
//
// The reference has no multi-GPU path; this is the one exchange step of ours: per action step the learner
// rank needs [obs | reward] of every rank's environments.  Instead of an NCCL all_gather launch per step (latency bound:
// 40-160 KB per rank), every rank's step kernel leaves its last rows in its OWN memory (a ring of slices, symmetric
// memory mapped into the learner's process) and stores one epoch word into the learner's flag array when it is done
// (qrl_em_args, QRL_EM_PEER_PUBLISH).  This kernel runs on the learner, usually on a side stream next to the following
// step: the blocks that copy rank p's slice poll flags[p] first (one thread, nanosleep back-off), so the slices of ranks
// that have finished move over NVLink while the others still integrate; 16-byte loads from peer memory, 16-byte stores to
// the gathered block.  The last block advances the learner's epoch counter and hands every rank the credit (`pulled`)
// that lets it reuse the ring slice.
#include <chrono>
#if defined(__linux__)
#include <sys/prctl.h>
#endif
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace qrl {

// Small on purpose.  The gather of step i is enqueued right behind step i and polls while step i+1 already runs, so a
// gather block must be CO-RESIDENT with a step CTA: em_tile3_kernel takes 768 threads x 80 registers = 61440 of the SM's
// 65536 registers and ~200 KB of its shared memory, which leaves room for exactly 128 threads x 32 registers.  (A block
// that does not fit waits for the step CTA to exit, the next step's CTA then waits for the gather, and the gather waits
// for the other ranks: measured 50 instead of 27 us per step with 40 registers per thread.)
constexpr int PULL_THREADS = 128;

__global__ void __launch_bounds__(PULL_THREADS, 16) peer_pull_kernel(const qrl_peer_pull_args a, const int bpr) {
    const int p = blockIdx.x / bpr, b = blockIdx.x - p * bpr;
    const unsigned long long epoch = *a.pull_epoch + 1ull;
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        const unsigned long long t0 = global_timer_ns();
        while (true) {
            unsigned long long seen;       // relaxed polling (L2), one acquire fence once the epoch has arrived
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.flags + p) : "memory");
            if (seen >= epoch) { __threadfence_system(); break; }
            if (global_timer_ns() - t0 > 2000000000ull) {   // a rank never published: fail loudly, do not hang
                if (a.status) atomicOr(a.status, QRL_STATUS_PEER_TIMEOUT);
                ok = 0;
                break;
            }
            __nanosleep(200);
        }
        s_ok = ok;
    }
    __syncthreads();
    if (s_ok) {
        const int64_t cnt = a.env_count[p], off = a.env_offset[p];
        const double* src = a.peer_ring[p] + (int64_t)(epoch % (unsigned long long)a.out_slots) * a.slot_stride[p];
        // slice = [obs (cnt, n_obs) | reward (cnt,)]; both parts are whole numbers of doubles, copied as 16-byte pieces
        // where the alignment allows (ring slices and the gathered blocks are 16-byte aligned when cnt * n_obs is even)
        const int64_t n_o = cnt * a.n_obs, n_r = cnt;
        double* dst_o = a.obs_all + off * a.n_obs;
        double* dst_r = a.reward_all + off;
        const int64_t stride = (int64_t)bpr * PULL_THREADS;
        const int64_t t = (int64_t)b * PULL_THREADS + threadIdx.x;
        const bool vec_o = ((n_o & 1) == 0) && (((uintptr_t)src | (uintptr_t)dst_o) % 16 == 0);
        if (vec_o) {
            const double2* s2 = reinterpret_cast<const double2*>(src);
            double2* d2 = reinterpret_cast<double2*>(dst_o);
            const int64_t n2 = n_o >> 1;
            int64_t i = t;
            for (; i + 3 * stride < n2; i += 4 * stride) {      // 4 independent 16-byte loads in flight per thread
                const double2 v0 = s2[i], v1 = s2[i + stride], v2 = s2[i + 2 * stride], v3 = s2[i + 3 * stride];
                d2[i] = v0; d2[i + stride] = v1; d2[i + 2 * stride] = v2; d2[i + 3 * stride] = v3;
            }
            for (; i < n2; i += stride) d2[i] = s2[i];
        } else {
            for (int64_t i = t; i < n_o; i += stride) dst_o[i] = src[i];
        }
        for (int64_t i = t; i < n_r; i += stride) dst_r[i] = src[n_o + i];
    }
    // last block: advance the epoch, hand out the credits
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    const unsigned prev = atomicAdd(a.finish_counter, 1u);
    if (prev != gridDim.x - 1) return;
    __threadfence_system();
    *a.finish_counter = 0u;
    *a.pull_epoch = epoch;
    for (int q = 0; q < a.world; ++q)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_pulled[q]), "l"(epoch) : "memory");
}

// one warp: lane p polls flags[p] until it has reached `epoch`
__global__ void __launch_bounds__(32) peer_wait_kernel(const uint64_t* flags, int world, unsigned long long epoch, int* status) {
    const unsigned long long t0 = global_timer_ns();
    for (int p = threadIdx.x; p < world; p += 32) {
        while (true) {
            unsigned long long seen;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + p) : "memory");
            if (seen >= epoch) break;
            if (global_timer_ns() - t0 > 2000000000ull) {
                if (status) atomicOr(status, QRL_STATUS_PEER_TIMEOUT);
                break;
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

}  // namespace qrl

extern "C" int qrl_peer_push(const qrl_peer_push_args* args, void* main_stream, void* side_stream) {
    using namespace qrl;
    QRL_REQUIRE(args != nullptr, "qrl_peer_push: args is NULL");
    const qrl_peer_push_args& a = *args;
    QRL_REQUIRE(a.src && a.dst && a.bytes >= 0 && a.flag_src && a.flag_dst && a.step_done && a.copy_done,
                "qrl_peer_push: NULL argument");
    QRL_REQUIRE(main_stream != side_stream, "qrl_peer_push: the copies need a stream of their own");
    cudaStream_t ms = (cudaStream_t)main_stream, ss = (cudaStream_t)side_stream;
    QRL_CUDA(cudaEventRecord((cudaEvent_t)a.step_done, ms));
    QRL_CUDA(cudaStreamWaitEvent(ss, (cudaEvent_t)a.step_done, 0));
    if (a.bytes > 0) QRL_CUDA(cudaMemcpyAsync(a.dst, a.src, (size_t)a.bytes, cudaMemcpyDefault, ss));
    if (a.src2 && a.dst2 && a.bytes2 > 0) QRL_CUDA(cudaMemcpyAsync(a.dst2, a.src2, (size_t)a.bytes2, cudaMemcpyDefault, ss));
    QRL_CUDA(cudaMemcpyAsync(a.flag_dst, a.flag_src, 8, cudaMemcpyDefault, ss));
    QRL_CUDA(cudaEventRecord((cudaEvent_t)a.copy_done, ss));
    return 0;
}

// ---- worker-thread context of the push exchange -------------------------------------------------------------------
// Single producer (the stepping thread) / single consumer (the worker) ring of jobs, lock free on the producer's side.
// While a step loop is running the worker polls the ring with short timed sleeps (~25-70 us): a condition-variable
// wake-up per step costs the stepping thread a futex system call (measured 36 instead of 27 us per step on two of eight
// ranks), a busy-polling worker takes a core that eight ranks on one box do not have to spare (34.6 us) — and the
// transfer is not latency critical (the ring of slices is 16 steps deep).  After ~20 ms without a job the worker sleeps on
// a condition variable and the next push wakes it.
namespace qrl {
struct PeerCtx {
    static constexpr uint64_t RING = 256;
    int device = 0;
    cudaStream_t side = nullptr;
    qrl_peer_push_args ring[RING];
    std::atomic<uint64_t> head{0};             // jobs produced (sequence number of the newest job)
    std::atomic<uint64_t> issued{0};           // jobs whose driver calls have been made
    std::atomic<int> error{0};
    std::atomic<bool> sleeping{false}, stop{false};
    std::mutex mu;
    std::condition_variable cv;
    std::thread worker;

    void run() {
        cudaSetDevice(device);
#if defined(__linux__)
        prctl(PR_SET_TIMERSLACK, 1000UL, 0UL, 0UL, 0UL);   // 1 us instead of the default 50 us of timer slack for this thread
#endif
        uint64_t done = 0;
        int idle = 0;
        for (;;) {
            if (done < head.load(std::memory_order_acquire)) {
                idle = 0;
                const uint64_t seq = done + 1;
                const qrl_peer_push_args a = ring[seq % RING];
                cudaError_t e = cudaStreamWaitEvent(side, (cudaEvent_t)a.step_done, 0);
                if (e == cudaSuccess && a.bytes > 0) e = cudaMemcpyAsync(a.dst, a.src, (size_t)a.bytes, cudaMemcpyDefault, side);
                if (e == cudaSuccess && a.src2 && a.dst2 && a.bytes2 > 0) e = cudaMemcpyAsync(a.dst2, a.src2, (size_t)a.bytes2, cudaMemcpyDefault, side);
                if (e == cudaSuccess) e = cudaMemcpyAsync(a.flag_dst, a.flag_src, 8, cudaMemcpyDefault, side);
                if (e == cudaSuccess) e = cudaEventRecord((cudaEvent_t)a.copy_done, side);
                if (e != cudaSuccess && error.load() == 0) error.store((int)e);
                done = seq;
                issued.store(seq, std::memory_order_release);
                continue;
            }
            if (stop.load()) return;
            if (++idle < 400) {                // ~20 ms of timed polling
                std::this_thread::sleep_for(std::chrono::microseconds(20));
                continue;
            }
            std::unique_lock<std::mutex> lk(mu);
            sleeping.store(true);
            if (done >= head.load(std::memory_order_acquire) && !stop.load())
                cv.wait_for(lk, std::chrono::milliseconds(50));
            sleeping.store(false);
            idle = 0;
        }
    }
};
}  // namespace qrl

extern "C" int qrl_peer_ctx_create(void** ctx, void* side_stream) {
    using namespace qrl;
    QRL_REQUIRE(ctx != nullptr && side_stream != nullptr, "qrl_peer_ctx_create: NULL argument");
    PeerCtx* c = new PeerCtx();
    QRL_CUDA(cudaGetDevice(&c->device));
    c->side = (cudaStream_t)side_stream;
    c->worker = std::thread([c] { c->run(); });
    *ctx = c;
    return 0;
}

extern "C" uint64_t qrl_peer_ctx_push(void* ctx, const qrl_peer_push_args* args, void* main_stream) {
    using namespace qrl;
    if (ctx == nullptr || args == nullptr || !args->src || !args->dst || !args->flag_src || !args->flag_dst || !args->step_done ||
        !args->copy_done) {
        fail(QRL_E_INVALID, "%s", "qrl_peer_ctx_push: NULL argument");
        return 0;
    }
    PeerCtx* c = static_cast<PeerCtx*>(ctx);
    if (c->error.load() != 0) { cuda_fail((cudaError_t)c->error.load(), "qrl_peer_ctx worker"); return 0; }
    // in stream order, right behind the step launch: this one call stays on the caller's thread
    cudaError_t e = cudaEventRecord((cudaEvent_t)args->step_done, (cudaStream_t)main_stream);
    if (e != cudaSuccess) { cuda_fail(e, "cudaEventRecord(step_done)"); return 0; }
    const uint64_t seq = c->head.load(std::memory_order_relaxed) + 1;
    while (seq - c->issued.load(std::memory_order_acquire) >= PeerCtx::RING) std::this_thread::yield();   // (never in practice)
    c->ring[seq % PeerCtx::RING] = *args;
    c->head.store(seq, std::memory_order_release);
    if (c->sleeping.load()) {
        std::lock_guard<std::mutex> lk(c->mu);
        c->cv.notify_one();
    }
    return seq;
}

extern "C" int qrl_peer_ctx_wait(void* ctx, uint64_t seq) {
    using namespace qrl;
    QRL_REQUIRE(ctx != nullptr, "qrl_peer_ctx_wait: ctx is NULL");
    if (seq == 0) return 0;
    PeerCtx* c = static_cast<PeerCtx*>(ctx);
    QRL_REQUIRE(seq <= c->head.load(), "qrl_peer_ctx_wait: unknown sequence number");
    while (c->issued.load(std::memory_order_acquire) < seq) {
        if (c->error.load() != 0) return cuda_fail((cudaError_t)c->error.load(), "qrl_peer_ctx worker");
        std::this_thread::yield();
    }
    if (c->error.load() != 0) return cuda_fail((cudaError_t)c->error.load(), "qrl_peer_ctx worker");
    // the slot of job `seq` still holds its arguments unless RING newer jobs have been produced since (then its copy has
    // long finished and its event has been re-recorded for a later job)
    if (c->head.load() - seq < PeerCtx::RING - 1) QRL_CUDA(cudaEventSynchronize((cudaEvent_t)c->ring[seq % PeerCtx::RING].copy_done));
    return 0;
}

extern "C" int qrl_peer_ctx_destroy(void* ctx) {
    using namespace qrl;
    if (ctx == nullptr) return 0;
    PeerCtx* c = static_cast<PeerCtx*>(ctx);
    c->stop.store(true);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->cv.notify_all();
    }
    if (c->worker.joinable()) c->worker.join();
    delete c;
    return 0;
}

extern "C" int qrl_peer_event_wait(void* event) {
    using namespace qrl;
    QRL_REQUIRE(event != nullptr, "qrl_peer_event_wait: event is NULL");
    QRL_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return 0;
}

extern "C" int qrl_peer_wait(const uint64_t* flags, int32_t world, uint64_t epoch, int32_t* status, void* stream) {
    using namespace qrl;
    QRL_REQUIRE(flags != nullptr && world >= 1, "qrl_peer_wait: bad arguments");
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, world, (unsigned long long)epoch, status);
    return after_launch("peer_wait_kernel");
}

extern "C" int qrl_peer_pull(const qrl_peer_pull_args* args, void* stream) {
    using namespace qrl;
    QRL_REQUIRE(args != nullptr, "qrl_peer_pull: args is NULL");
    const qrl_peer_pull_args& a = *args;
    QRL_REQUIRE(a.world >= 1 && a.n_obs >= 1 && a.out_slots >= 1 && a.blocks_per_rank >= 0, "qrl_peer_pull: bad sizes");
    QRL_REQUIRE(a.flags && a.peer_ring && a.slot_stride && a.env_offset && a.env_count && a.obs_all && a.reward_all &&
                a.pull_epoch && a.peer_pulled && a.finish_counter, "qrl_peer_pull: NULL argument");
    const int bpr = a.blocks_per_rank > 0 ? a.blocks_per_rank : 8;
    peer_pull_kernel<<<a.world * bpr, PULL_THREADS, 0, (cudaStream_t)stream>>>(a, bpr);
    return after_launch("peer_pull_kernel");
}
