// Throughput of the in-kernel noise generator at saturation: Philox4x32-10 alone, the normal transform alone, and both —
// round 1's polynomial Box-Muller (philox_r1.cuh, kept here for the comparison) against round 2's table-driven one with
// the tables in shared memory (the tiled kernel) and behind L1 (the other kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o rng_rate rng_rate.cu && ./rng_rate
#include <cstdio>
#include "../../quantrl_b200/csrc/philox.cuh"
#include "philox_r1.cuh"
using namespace qrl;
// GEN: 0 = round 1, 1 = round 2 / shared-memory tables, 2 = round 2 / L1 tables;  MODE: 0 philox, 1 transform, 2 both
template <int GEN, int MODE>
__global__ void k(double* out, int iters, long long* cyc, PhiloxKeys keys) {
    __shared__ __align__(16) unsigned char tabs[8192];
    RngTabSmem ts = rng_tables_to_smem(tabs, threadIdx.x, blockDim.x);
    __syncthreads();
    uint4 c = make_uint4(threadIdx.x, blockIdx.x, 7u, 1u);
    double acc = 0.0;
    uint32_t acci = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        c.x += 1;
        uint4 b;
        if (MODE == 0 || MODE == 2) b = philox4x32_10(c, keys); else b = make_uint4(c.x * 2654435761u, c.x ^ 0x9E3779B9u, c.x + 12345u, c.x * 40503u);
        if (MODE == 1 || MODE == 2) {
            double z0, z1;
            if (GEN == 0) qrl_r1::normal_pair(b, z0, z1);
            else if (GEN == 1) normal_pair(b, ts, z0, z1);
            else normal_pair(b, RngTabL1(), z0, z1);
            acc += z0 * z1;
        } else acci ^= b.x ^ b.y ^ b.z ^ b.w;
    }
    long long t1 = clock64();
    if (acc == 1.2345 || acci == 0x1234567u) out[0] = acc + acci;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int GEN, int MODE> void run(const char* name, int w) {
    double* o; long long* c; cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    int iters = 2000;
    PhiloxKeys keys = make_philox_keys(0x123456789ULL);
    k<GEN, MODE><<<148, 128 * w>>>(o, iters, c, keys); cudaDeviceSynchronize();
    k<GEN, MODE><<<148, 128 * w>>>(o, iters, c, keys); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP %d: %7.1f cycles per call per warp -> %6.1f cycles per warp-call per SMSP\n", name, w, (double)h / iters, (double)h / iters / w);
    cudaFree(o); cudaFree(c);
}
int main() {
    for (int w : {1, 2, 4, 6, 8}) {
        run<0, 0>("philox only", w);
        run<0, 1>("r1 transform only", w);
        run<1, 1>("r2 transform only (smem tables)", w);
        run<2, 1>("r2 transform only (L1 tables)", w);
        run<0, 2>("r1 philox + transform", w);
        run<1, 2>("r2 philox + transform (smem)", w);
        run<2, 2>("r2 philox + transform (L1)", w);
    }
    return 0;
}
