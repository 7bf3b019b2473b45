"""load_many in a loop vs iter_many (sampling of the next group overlapped with the extracts of the
current one on a second stream), products shape, one GPU."""
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
fan = [15, 10, 5]
for B in (4, 8, 16):
    G = 30
    seeds = dgs_synth.seed_batches(N, 1024, B * (G + 4), seed=B).pin_memory()
    groups = [seeds[j * B:(j + 1) * B] for j in range(G + 4)]
    for j in range(4):
        loader.load_many(groups[j], fan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(4, G + 4):
        res = loader.load_many(groups[j], fan)
    torch.cuda.synchronize()
    t_seq = (time.perf_counter() - t0) / (G * B) * 1e6
    for _ in loader.iter_many(groups[:4], fan):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for res in loader.iter_many(groups[4:], fan):
        pass
    torch.cuda.synchronize()
    t_ovl = (time.perf_counter() - t0) / (G * B) * 1e6
    print(f"B={B:2d}  load_many {t_seq:7.1f} us/batch   iter_many {t_ovl:7.1f} us/batch", flush=True)
