// C-ABI plumbing: version, error string, launch counter, device count and the FP64 peak micro-benchmark.
#include "common.cuh"

namespace qrl {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

// 8 independent DFMA chains per thread: saturates the FP64 pipe (64 lanes/clk/SM) with enough warps resident.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}
// one warp, one dependent DFMA chain: cycles per dependent FP64 FMA (pipeline latency)
__global__ void dfma_latency_kernel(double* out, long long* cycles, int n) {
    double a = 1.0 + threadIdx.x * 1e-9;
    const double m = 0.999999, b = 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = fma(a, m, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; out[0] = a; }
}
}  // namespace qrl

extern "C" int qrl_abi_version(void) { return QRL_ABI_VERSION; }
extern "C" const char* qrl_last_error(void) { return qrl::g_err; }
extern "C" uint64_t qrl_launch_count(void) { return qrl::g_launches.load(); }
extern "C" int qrl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int qrl_release(void) { return 0; }
extern "C" int qrl_host_device_pointer(const void* h_ptr, void** device_ptr) {
    using namespace qrl;
    QRL_REQUIRE(h_ptr != nullptr && device_ptr != nullptr, "qrl_host_device_pointer: NULL argument");
    cudaPointerAttributes at;
    QRL_CUDA(cudaPointerGetAttributes(&at, h_ptr));
    QRL_REQUIRE(at.type == cudaMemoryTypeHost && at.devicePointer != nullptr,
                "qrl_host_device_pointer: not page-locked host memory mapped for the current device");
    *device_ptr = at.devicePointer;
    return 0;
}

extern "C" int qrl_measure_fp64_latency(double* cycles_per_dfma) {
    using namespace qrl;
    QRL_REQUIRE(cycles_per_dfma != nullptr, "qrl_measure_fp64_latency: NULL output");
    double* d_out = nullptr;
    long long* d_cyc = nullptr;
    QRL_CUDA(cudaMalloc(&d_out, sizeof(double)));
    QRL_CUDA(cudaMalloc(&d_cyc, sizeof(long long)));
    const int n = 1 << 14;
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
        dfma_latency_kernel<<<1, 32>>>(d_out, d_cyc, n);
        int rc = after_launch("dfma_latency_kernel");
        if (rc) return rc;
        QRL_CUDA(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    }
    *cycles_per_dfma = (double)h / n;
    cudaFree(d_out);
    cudaFree(d_cyc);
    return 0;
}

extern "C" int qrl_measure_fp64_peak(int iters, double* tflops) {
    using namespace qrl;
    QRL_REQUIRE(tflops != nullptr && iters > 0, "qrl_measure_fp64_peak: bad arguments");
    int dev = 0, sms = 0;
    QRL_CUDA(cudaGetDevice(&dev));
    QRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* d_out = nullptr;
    QRL_CUDA(cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    QRL_CUDA(cudaEventCreate(&e0));
    QRL_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    double best_ms = 1e30;
    for (int rep = 0; rep < 5; ++rep) {  // rep 0 is the warm-up
        QRL_CUDA(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0);
        int rc = after_launch("dfma_peak_kernel");
        if (rc) return rc;
        QRL_CUDA(cudaEventRecord(e1, 0));
        QRL_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        QRL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best_ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return 0;
}
