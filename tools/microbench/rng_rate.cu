// Throughput of the two halves of the in-kernel noise generator at saturation (8+ warps per scheduler):
// Philox4x32-10 alone, the Box-Muller polynomial kernels alone, and both together.
#include <cstdio>
#include "../../quantrl_b200/csrc/philox.cuh"
using namespace qrl;
template <int MODE>
__global__ void k(double* out, int iters, long long* cyc, PhiloxKeys keys) {
    uint4 c = make_uint4(threadIdx.x, blockIdx.x, 7u, 1u);
    double acc = 0.0;
    uint32_t acci = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        c.x += 1;
        uint4 b;
        if (MODE == 0 || MODE == 2) b = philox4x32_10(c, keys); else b = make_uint4(c.x * 2654435761u, c.x ^ 0x9E3779B9u, c.x + 12345u, c.x * 40503u);
        if (MODE == 1 || MODE == 2) { double z0, z1; normal_pair(b, z0, z1); acc += z0 * z1; } else acci ^= b.x ^ b.y ^ b.z ^ b.w;
    }
    long long t1 = clock64();
    if (acc == 1.2345 || acci == 0x1234567u) out[0] = acc + acci;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, int w) {
    double* o; long long* c; cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    int iters = 2000;
    PhiloxKeys keys = make_philox_keys(0x123456789ULL);
    k<MODE><<<148, 128 * w>>>(o, iters, c, keys); cudaDeviceSynchronize();
    k<MODE><<<148, 128 * w>>>(o, iters, c, keys); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-22s warps/SMSP %d: %.1f cycles per call per warp -> %.1f cycles per warp-call per SMSP\n", name, w, (double)h / iters, (double)h / iters / w);
}
int main() {
    for (int w : {1, 2, 4, 8}) { run<0>("philox only", w); run<1>("box-muller only", w); run<2>("philox + box-muller", w); }
    return 0;
}
