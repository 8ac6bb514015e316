// Issue-rate micro-benchmark (sm_100a): how many warp instructions per cycle can one SM sub-partition issue for
// INT32 (LOP3/IADD), IMAD, FP64 and mixes of them?  Used to explain the EM kernel's instruction-issue bound.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void k(uint32_t* out, double* outd, int iters, long long* cyc) {
    uint32_t a[4] = {threadIdx.x + 1u, threadIdx.x + 2u, threadIdx.x + 3u, threadIdx.x + 4u};
    uint32_t m[4] = {threadIdx.x * 3u + 1u, 5u, 7u, 9u};
    double d[4] = {1.0 + threadIdx.x, 2.0, 3.0, 4.0};
    const double c = 0.999999, e = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 6) a[j] = (a[j] ^ m[j]) + (a[j] >> 3);           // LOP3/SHF/IADD (ALU)
                if (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 6) m[j] = m[j] * 0x9E3779B9u + a[j];             // IMAD
                if (MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6) d[j] = fma(d[j], c, e);                       // DFMA
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = a[0] ^ a[1] ^ a[2] ^ a[3] ^ m[0] ^ m[1] ^ m[2] ^ m[3];
    if (s == 0x12345678u) out[0] = s;
    if (d[0] + d[1] + d[2] + d[3] == 1.2345) outd[0] = d[0];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, int warps_per_smsp) {
    uint32_t* o; double* od; long long* c; cudaMalloc(&o, 4); cudaMalloc(&od, 8); cudaMalloc(&c, 8);
    int iters = 2000;
    k<MODE><<<148, 128 * warps_per_smsp>>>(o, od, iters, c); cudaDeviceSynchronize();
    k<MODE><<<148, 128 * warps_per_smsp>>>(o, od, iters, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SMSP %d: %.1f cycles per 32-op group per warp -> %.3f groups/cycle/SMSP\n", name, warps_per_smsp,
           (double)h / (iters * 8.0), warps_per_smsp * iters * 8.0 / (double)h);
    cudaFree(o); cudaFree(od); cudaFree(c);
}
int main() {
    for (int w : {1, 2, 4, 8}) {
        run<0>("ALU (xor+shift+add x4)", w);
        run<1>("IMAD x4", w);
        run<2>("DFMA x4", w);
        run<3>("ALU + IMAD", w);
        run<4>("ALU + DFMA", w);
        run<5>("IMAD + DFMA", w);
        run<6>("ALU + IMAD + DFMA", w);
    }
    return 0;
}
