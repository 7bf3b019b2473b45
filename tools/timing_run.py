"""Per-kernel warm timings of the fused pipeline (DGS_BLOCKS_TIMING=1 must be set)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch, dgs, dgs_synth
dev = torch.device("cuda", 0)
N, E, D, dt = dgs_synth.SHAPES["products"]
BIAS = os.environ.get("BIAS", "0") == "1"
ip, ix, pr = dgs_synth.make_csr(N, E, device=dev, weights=BIAS)
sampler = dgs.classes.CSRSampler(ip, ix, pr if BIAS else None)
B = int(os.environ.get("BATCH", "1024"))
FAN = [int(x) for x in os.environ.get("FAN", "15,10,5").split(",")]
seeds = dgs_synth.seed_batches(N, B, 12, device=dev)
for i in range(12):
    out = sampler._CAPI_sample_node_classifiction(seeds[i], FAN)
print("edges", sum(b[2].numel() for b in out), "frontier", out[-1][1].numel(), [b[0].numel() for b in out])
