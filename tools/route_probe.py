"""Route kernels of the id-exchange extract on one GPU: dgs.ops.route_ids for 192 k and 1 M random
ids, world 8 (run under `ncu --metrics gpu__time_duration.sum -k regex:route` for per-kernel times,
or plain for the event-timed total)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402

import dgs  # noqa: E402

dev = torch.device("cuda", 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (192_355, 1_000_000):
    q = torch.randint(0, 2_449_029, (n,), device=dev)
    for _ in range(3):
        dgs.ops.route_ids(q, 8)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        keep = dgs.ops.route_ids(q, 8)
    e1.record()
    torch.cuda.synchronize()
    print(f"route_ids n={n} world=8: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (3 launches)", flush=True)
