"""Host-side breakdown of one end-to-end step through BatchLoader.load (DGS_LOADER_TRACE=1 prints the
native call's internal laps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch, dgs, dgs_synth
dev = torch.device("cuda", 0)
N, E, D, dt = dgs_synth.SHAPES["products"]
ip, ix, _ = dgs_synth.make_csr(N, E, device=dev)
ft = dgs_synth.make_features(N, D, dt, device=dev)
labels = (torch.arange(N, device=dev) % 47)
smp = dgs.classes.CSRSampler(ip, ix)
loader = dgs.classes.BatchLoader(smp, ft, labels)
seeds = dgs_synth.seed_batches(N, 1024, 700).pin_memory()
lab = torch.empty(1024, dtype=torch.int64).pin_memory()
for i in range(50):
    loader.load(seeds[i], [15, 10, 5], labels_out=lab)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50, 650):
    loader.load(seeds[i], [15, 10, 5], labels_out=lab)
print("us/step BatchLoader.load:", (time.perf_counter() - t0) / 600 * 1e6)
t0 = time.perf_counter()
for i in range(100):
    a = torch.empty(3_000_000, dtype=torch.int64, device=dev)
    b = torch.empty((200_000, 100), device=dev)
    c = torch.empty(1024, dtype=torch.int64, device=dev)
print("us for the 3 torch.empty:", (time.perf_counter() - t0) / 100 * 1e6)
