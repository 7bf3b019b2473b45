"""How long does one weighted (A-Res) row selection take?  Star graphs whose rows all have the same
degree; few enough seeds that every warp gets at most one row => kernel time ~ single-row latency."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch, dgs
dev = torch.device("cuda", 0)
K = int(os.environ.get("K", "25"))
DEGS = [int(x) for x in os.environ.get("DEGS", "64,256,512,2048,8192").split(",")]
SEEDS = [int(x) for x in os.environ.get("SEEDS", "256,2048,16384").split(",")]
for deg in DEGS:
    for S in SEEDS:
        n = S
        indptr = (torch.arange(n + 1, dtype=torch.int64) * deg).to(dev)
        indices = torch.randint(0, n, (n * deg,), dtype=torch.int64, device=dev)
        probs = torch.rand(n * deg, device=dev) + 0.05
        seeds = torch.arange(S, device=dev)
        for rep in (False,):
            for _ in range(3):
                dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, K, rep, rng_seed=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(10):
                dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, K, rep, rng_seed=i)
            e1.record(); torch.cuda.synchronize()
            print(f"deg {deg:5d} seeds {S:6d} k {K}: {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us per call (plan + pick + host sync)")
