"""BASELINE config 3: biased (edge-weight) sampling fan-out [25,10] on the products-shaped graph with
the GPU location-table feature / structure cache, cache-ratio sweep 0-100 % (ratio 0 = un-cached ops
path over pinned host memory).  One GPU.  Writes one JSON object per ratio.

  python tools/run_configs.py [--shape products] [--ratios 0,0.01,0.05,0.1,0.25,0.5,1.0]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402

import dgs  # noqa: E402
import dgs_synth  # noqa: E402
from DistGNN.cache import get_cache_nids_by_degree  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fan-out", default="25,10")
    ap.add_argument("--ratios", default="0,0.01,0.05,0.1,0.25,0.5,1.0")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--uniform", action="store_true")
    args = ap.parse_args()
    fan = [int(x) for x in args.fan_out.split(",")]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    ip, ix, pr = dgs_synth.make_csr(N, E, device=dev, weights=not args.uniform)
    ft = dgs_synth.make_features(N, D, dt, device=dev)
    ipc, ixc, ftc = ip.cpu().pin_memory(), ix.cpu().pin_memory(), ft.cpu().pin_memory()
    prc = pr.cpu().pin_memory() if pr is not None else torch.Tensor()
    del ip, ix, pr, ft
    torch.cuda.empty_cache()
    row_bytes = D * ftc.element_size()
    seeds = dgs_synth.seed_batches(N, args.batch, args.steps + 5, device=dev)
    for r in [float(x) for x in args.ratios.split(",")]:
        if r == 0:
            smp = dgs.classes.CSRSampler(ipc, ixc, prc if prc.numel() else None, device=dev)
            extract = lambda nids: dgs.ops._CAPI_cuda_index_select(ftc, nids)
        else:
            cache = get_cache_nids_by_degree(ipc, r)
            smp = dgs.classes.P2PCacheSampler(ipc, ixc, prc, cache, 0)
            fs = dgs.classes.P2PCacheFeatureServer(ftc, cache, 0)
            extract = fs._CAPI_get_feature
        for i in range(5):
            extract(smp._CAPI_sample_node_classifiction(seeds[i], fan)[-1][1])
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        edges = rows = 0
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            blocks = smp._CAPI_sample_node_classifiction(seeds[5 + i], fan)
            x = extract(blocks[-1][1])
            edges += sum(b[2].numel() for b in blocks)
            rows += x.shape[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"config": "biased" if not args.uniform else "uniform", "shape": args.shape,
                          "fan_out": fan, "batch": args.batch, "cache_ratio": r,
                          "ms_per_step": ms / args.steps, "batches_per_sec": args.steps / (ms * 1e-3),
                          "sampled_edges_per_sec": edges / (ms * 1e-3),
                          "extract_gbps": rows * (2 * row_bytes + 8) / (ms * 1e-3) / 1e9,
                          "rows_per_step": rows / args.steps, "edges_per_step": edges / args.steps,
                          "wall_ms_per_step": (time.perf_counter() - t0) / args.steps * 1e3}), flush=True)
        if r != 0:
            smp.close()
            fs.close()
            del smp, fs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
