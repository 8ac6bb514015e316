// Where does the third-generation tiled EM kernel spend its time?  Builds em_tile3_kernel.cu with -DQRL_T3_TIMELINE (per-warp
// cycle accumulators around every phase and mbarrier wait) and prints, for a config-3-sized launch, the per-role
// averages.    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DQRL_T3_TIMELINE -o t3_timeline t3_timeline.cu
//              ./t3_timeline [E] [n] [K]          (knobs: QRL_EM_T3_C / QRL_EM_T3_D / QRL_EM_T3_THREADS)
#ifndef T3_SRC
#define T3_SRC "../../quantrl_b200/csrc/em_tile3_kernel.cu"
#endif
#include T3_SRC
#include <vector>
namespace qrl {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    using namespace qrl;
    const int E = argc > 1 ? atoi(argv[1]) : 4096, n = argc > 2 ? atoi(argv[2]) : 4, K = argc > 3 ? atoi(argv[3]) : 100;
    double *A, *nz, *std_, *x, *S, *O, *R, *rl, *rc, *xl, *ol, *w, *mm, *cache;
    int* nan_;
    int* sel;
    CK(cudaMalloc(&A, sizeof(double) * E * n * n)); CK(cudaMalloc(&nz, 8 * n)); CK(cudaMalloc(&std_, 8 * n));
    CK(cudaMalloc(&x, 8 * E * n)); CK(cudaMalloc(&S, 8ll * (K + 1) * E * n)); CK(cudaMalloc(&O, 8ll * (K + 1) * E * n));
    CK(cudaMalloc(&R, 8ll * (K + 1) * E)); CK(cudaMalloc(&rl, 8 * E)); CK(cudaMalloc(&rc, 8 * E)); CK(cudaMalloc(&xl, 8 * E * n));
    CK(cudaMalloc(&ol, 8 * E * n)); CK(cudaMalloc(&w, 8 * n)); CK(cudaMalloc(&mm, 16)); CK(cudaMalloc(&nan_, 4)); CK(cudaMalloc(&sel, 4));
    CK(cudaMalloc(&cache, 8ll * E * (K + 1) * 2));
    std::vector<double> hA(E * n * n), hn(n, 0.05), hs(n, 0.01), hx(E * n, 1.0), hw(n, 1.0);
    for (size_t i = 0; i < hA.size(); ++i) hA[i] = 0.3 * sin(0.37 * i);
    double hmm[2] = {INFINITY, -INFINITY};
    int hsel = -1;
    CK(cudaMemcpy(A, hA.data(), 8 * hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(nz, hn.data(), 8 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(std_, hs.data(), 8 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(x, hx.data(), 8 * E * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w, hw.data(), 8 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(mm, hmm, 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sel, &hsel, 4, cudaMemcpyHostToDevice)); CK(cudaMemset(nan_, 0, 4)); CK(cudaMemset(rc, 0, 8 * E));
    qrl_em_args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.n_envs = E; a.K = K; a.t_ssz = 0.01; a.A = A; a.A_stride_e = n * n; a.nz = nz; a.rng_mode = QRL_RNG_PHILOX; a.seed = 1;
    a.obs_std = std_; a.x0 = x; a.states = S; a.observations = O; a.x_last = xl; a.obs_last = ol; a.reward_kind = QRL_REWARD_NEG_WSQ_OBS;
    a.reward_w = w; a.reward = R; a.reward_last = rl; a.rewards_cum = rc; a.obs_minmax = mm; a.nan_flag = nan_;
    a.packed = cache; a.packed_stride_e = (K + 1) * 2; a.packed_row_pitch = 2; a.T_step = S; a.actions = x; a.n_actions = 1;
    a.sel_idxs = sel; a.n_sel = 1;
    for (int i = 0; i < 5; ++i) if (em_tile3_launch(a, 0) != 0) { printf("launch failed: %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 200; ++i) em_tile3_launch(a, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const T3Plan p = t3_plan(E, n, K, 1, device_sm_count());
    printf("E=%d n=%d K=%d: %.2f us per launch; plan G=%d C=%d D=%d nphase=%d rec threads=%d threads=%d grid=%d smem=%zu\n", E, n, K,
           ms * 1e3 / 200, p.G, p.C, p.D, p.nphase, p.nrt, p.threads, p.grid, p.smem);
#ifndef QRL_T3_TIMELINE
    return 0;
#else
    static unsigned long long tl[256][32][8];
    CK(cudaMemcpyFromSymbol(tl, t3_tl, sizeof(tl)));
    const int nw = p.threads / 32, nrw = p.nrt / 32;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cyc_us = 1e3 / (double)clk_khz * 1.0;   // us per cycle at the nominal clock (indicative)
    const int nb = p.grid < 256 ? p.grid : 256;
    double pro = 0, loop = 0, epi = 0, wprod = 0, wout = 0, wwx = 0, wwe = 0, rcomp = 0, rwait = 0;
    for (int b = 0; b < nb; ++b) {
        for (int wi = 0; wi < nw; ++wi) {
            const unsigned long long* t = tl[b][wi];
            pro += (double)(t[1] - t[0]); loop += (double)(t[6] - t[1]); epi += (double)(t[7] - t[6]);
            if (wi < nrw) { rcomp += t[2]; rwait += t[4]; }
            else { wprod += t[2]; wout += t[3]; wwx += t[4]; wwe += t[5]; }
        }
    }
    const double nwk = (double)nb * (nw - nrw), nrc = (double)nb * nrw, nall = (double)nb * nw;
    printf("absolute (globaltimer, mean over warps): prologue %.2f us, main loop %.2f us, epilogue up to the finish protocol %.2f us\n",
           pro / nall * 1e-3, loop / nall * 1e-3, epi / nall * 1e-3);
    printf("workers (cycles -> us at %.0f MHz): produce %.2f, rows out %.2f, wait full_x %.2f, wait empty %.2f\n", clk_khz * 1e-3,
           wprod / nwk * cyc_us, wout / nwk * cyc_us, wwx / nwk * cyc_us, wwe / nwk * cyc_us);
    printf("recurrence warps: compute %.2f us, wait full_w %.2f us\n", rcomp / nrc * cyc_us, rwait / nrc * cyc_us);
    if (getenv("T3_TRACE")) {
        static unsigned long long tr[2][256];
        CK(cudaMemcpyFromSymbol(tr, t3_trace, sizeof(tr)));
        const unsigned long long t0 = tl[0][0][0];
        const int NC = (K + p.C - 1) / p.C;
        printf("trace of CTA 0 (us since kernel entry): prologue done %.2f\n", (tr[0][252] - t0) * 1e-3);
        for (int it = 0; it < NC + 2; ++it)
            printf("  worker it %2d: rows wait %.2f -> %.2f .. produce start %.2f end %.2f\n", it, it >= 2 ? (tr[1][4 * it] - t0) * 1e-3 : 0.0,
                   it >= 2 ? (tr[1][4 * it + 1] - t0) * 1e-3 : 0.0, it < NC ? (tr[1][4 * it + 2] - t0) * 1e-3 : 0.0, it < NC ? (tr[1][4 * it + 3] - t0) * 1e-3 : 0.0);
        for (int c = 0; c < NC; ++c) printf("  rec chunk %2d: wait from %.2f, start %.2f\n", c, (tr[0][2 * c] - t0) * 1e-3, (tr[0][2 * c + 1] - t0) * 1e-3);
        printf("  loop end: rec %.2f worker %.2f, after sync %.2f, kernel end mark %.2f\n", (tr[0][254] - t0) * 1e-3, (tr[1][254] - t0) * 1e-3,
               (tr[0][255] - t0) * 1e-3, (tl[0][0][7] - t0) * 1e-3);
    }
    return 0;
#endif
}
