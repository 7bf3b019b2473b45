"""cProfile of the host side of the e2e step (run on the GPU box)."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch, dgs, dgs_synth
dev = torch.device("cuda", 0)
N, E, D, dt = dgs_synth.SHAPES["products"]
ip, ix, _ = dgs_synth.make_csr(N, E, device=dev)
ft = dgs_synth.make_features(N, D, dt, device=dev)
labels = (torch.arange(N, device=dev) % 47)
sampler = dgs.classes.CSRSampler(ip, ix)
seeds_pin = dgs_synth.seed_batches(N, 1024, 260).pin_memory()
lab_host = torch.empty(1024, dtype=torch.int64).pin_memory()

def step(i):
    s = seeds_pin[i].to(dev, non_blocking=True)
    blocks = sampler._CAPI_sample_node_classifiction(s, [15, 10, 5], False)
    x = dgs.ops._CAPI_cuda_index_select(ft, blocks[-1][1])
    lab = dgs.ops._CAPI_cuda_index_select(labels, s)
    lab_host.copy_(lab, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return blocks, x

for i in range(20):
    step(i)
t0 = time.perf_counter()
for i in range(20, 120):
    step(i)
print("us/step un-profiled:", (time.perf_counter() - t0) / 100 * 1e6)
pr = cProfile.Profile()
pr.enable()
for i in range(120, 250):
    step(i)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
