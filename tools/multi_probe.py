"""Per-batch cost of the multi-batch sampling kernel: products shape, [15,10,5], batch 1024,
B = 1, 2, 4, 8, 16 mini-batches per cooperative launch, launches issued back to back.

  python tools/multi_probe.py [--bias] [--fan-out 15,10,5] [--shape products]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402

import dgs  # noqa: E402
import dgs_synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--fan-out", default="15,10,5")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--bias", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--bs", default="1,2,4,8,16")
    args = ap.parse_args()
    fan = [int(x) for x in args.fan_out.split(",")]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    ip, ix, pr = dgs_synth.make_csr(N, E, device=dev, weights=args.bias)
    smp = dgs.classes.CSRSampler(ip, ix, pr)
    pipe = smp._pipe
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for B in [int(b) for b in args.bs.split(",")]:
        seeds = dgs_synth.seed_batches(N, args.batch, B * (args.reps + 3), seed=B, device=dev)
        groups = [seeds[j * B:(j + 1) * B].contiguous() for j in range(args.reps + 3)]
        # untimed pass with as many arenas alive as the timed one: the caching allocator then owns
        # them and no cudaMalloc lands in the timed region
        keep = [pipe.enqueue_many(groups[j % 3], fan, False, list(range(B)), deliver_counts=False)[1]
                for j in range(args.reps)]
        del keep
        torch.cuda.synchronize()
        e0.record()
        keep = [pipe.enqueue_many(groups[3 + j], fan, False, list(range(j, j + B)), deliver_counts=False)[1]
                for j in range(args.reps)]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        out[B] = {"ms_per_launch": ms, "us_per_batch": ms * 1e3 / B}
        del keep
        print(f"B={B:2d}  {ms * 1e3:8.1f} us / launch  {ms * 1e3 / B:7.1f} us / batch", file=sys.stderr, flush=True)
    # the single-batch entry for comparison
    seeds = dgs_synth.seed_batches(N, args.batch, args.reps + 3, seed=99, device=dev)
    keep = [pipe.enqueue_only(seeds[j % 3], fan) for j in range(args.reps)]
    del keep
    torch.cuda.synchronize()
    e0.record()
    keep = [pipe.enqueue_only(seeds[3 + j], fan, rng_seed=j + 1) for j in range(args.reps)]
    e1.record()
    torch.cuda.synchronize()
    out["single_call"] = {"us_per_batch": e0.elapsed_time(e1) / args.reps * 1e3}
    print(json.dumps({"shape": args.shape, "fan_out": fan, "batch": args.batch, "bias": args.bias,
                      "result": out}))


if __name__ == "__main__":
    main()
