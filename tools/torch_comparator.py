"""Non-reference comparator: what the reference's stochastic path looks like when it is merely moved to
batched torch-CUDA ops on the same B200 — one Python loop per sub-step, ~8 small kernels each (randn x2, bmm, mul, add,
noise add, reward), exactly the shape of quantrl/envs/stochastic.py:218-242 + envs/base.py:462 vectorised over envs.
Reported next to the fused kernel in DESIGN.md; NOT used by the product path."""
import json, sys, time
import torch

def run(E=4096, n=4, K=100, steps=20, graph=False):
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(0)
    A = 0.3 * torch.randn((E, n, n), dtype=torch.float64, device=dev, generator=g)
    M = torch.eye(n, dtype=torch.float64, device=dev) + A * 0.01
    nz = torch.full((E, n), 0.05, dtype=torch.float64, device=dev)
    std = torch.full((E, n), 0.01, dtype=torch.float64, device=dev)
    S = torch.empty((K + 1, E, n), dtype=torch.float64, device=dev)
    O = torch.empty_like(S)
    R = torch.empty((K + 1, E), dtype=torch.float64, device=dev)
    S[0] = 1.0
    sq = 0.01 ** 0.5
    def interval():
        O[0] = S[0] + std * torch.randn((E, n), dtype=torch.float64, device=dev)
        for i in range(K):
            W = sq * torch.randn((E, n), dtype=torch.float64, device=dev)
            S[i + 1] = torch.bmm(M, S[i].unsqueeze(-1)).squeeze(-1) + nz * W
            O[i + 1] = S[i + 1] + std * torch.randn((E, n), dtype=torch.float64, device=dev)
        torch.sum(O * O, dim=-1, out=R).neg_()
        S[0].copy_(S[K])
    for _ in range(3):
        interval()
    torch.cuda.synchronize()
    if graph:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            interval()
        fn = gr.replay
    else:
        fn = interval
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"E": E, "n": n, "K": K, "cuda_graph": graph, "ms_per_interval": round(ms, 3), "env_substeps_per_s": f"{E * K / (ms * 1e-3):.3e}"}

if __name__ == "__main__":
    for graph in (False, True):
        print(json.dumps(run(graph=graph)))
