"""Where does a mini-batch step spend its time?  (run on the GPU box; writes a small JSON)

  python tools/breakdown.py [--shape products] [--batch 1024] [--fan-out 15,10,5]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402

import dgs  # noqa: E402
import dgs_synth  # noqa: E402
from dgs import _lib  # noqa: E402
from dgs._util import stream  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fan-out", default="15,10,5")
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    fan = [int(x) for x in args.fan_out.split(",")]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    ip, ix, _ = dgs_synth.make_csr(N, E, device=dev)
    ft = dgs_synth.make_features(N, D, dt, device=dev)
    sampler = dgs.classes.CSRSampler(ip, ix)
    seeds = dgs_synth.seed_batches(N, args.batch, args.reps + 5, device=dev)
    pipe = sampler._pipe
    l = _lib.lib()
    res = {}

    # full API call, wall clock (includes the one host sync)
    for i in range(5):
        blocks = sampler._CAPI_sample_node_classifiction(seeds[i], fan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.reps):
        blocks = sampler._CAPI_sample_node_classifiction(seeds[5 + i], fan)
    torch.cuda.synchronize()
    res["api_sample_us"] = (time.perf_counter() - t0) / args.reps * 1e6

    # raw C call: CPU enqueue time and GPU time
    S = args.batch
    pl = pipe._plan(S, fan)
    L = len(fan)
    arena = torch.empty(pl["total"], dtype=torch.int64, device=dev)
    fr, rows, cols, off = [], [], [], 0
    for u, n in zip(pl["ubs"], pl["nnz_ubs"]):
        fr.append(arena[off:off + u + n]); off += u + n
        rows.append(arena[off:off + n]); off += n
        cols.append(arena[off:off + n]); off += n
    counts_dev = torch.empty(2 * L, dtype=torch.int64, device=dev)
    a_fr = _lib.vp_array([t.data_ptr() for t in fr])
    a_r = _lib.vp_array([t.data_ptr() for t in rows])
    a_c = _lib.vp_array([t.data_ptr() for t in cols])

    def raw(i):
        _lib.check(l.dgs_sample_blocks(C.byref(pipe._graph), seeds[i].data_ptr(), S, L, pl["fo"], 0,
                                       C.c_uint64(i + 1), a_fr, a_r, a_c, pl["cap_edges"], pl["cap_front"],
                                       counts_dev.data_ptr(), pl["ws"].data_ptr(), pl["ws_bytes"], pl["epoch"], None, stream()))
        pl["epoch"] += 1

    raw(0)
    torch.cuda.synchronize()
    cpu = gpu = 0.0
    for i in range(args.reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        raw(5 + i)
        cpu += time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        gpu += e0.elapsed_time(e1)
    res["raw_sample_cpu_enqueue_us"] = cpu / args.reps * 1e6
    res["raw_sample_gpu_latency_us"] = gpu / args.reps * 1e3
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.reps):
        raw(5 + i)
    e1.record()
    torch.cuda.synchronize()
    res["raw_sample_gpu_back_to_back_us"] = e0.elapsed_time(e1) / args.reps * 1e3

    # extract
    frontier = blocks[-1][1]
    x = dgs.ops._CAPI_cuda_index_select(ft, frontier)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.reps):
        x = dgs.ops._CAPI_cuda_index_select(ft, frontier)
    res["api_extract_cpu_enqueue_us"] = (time.perf_counter() - t0) / args.reps * 1e6
    torch.cuda.synchronize()
    res["frontier_rows"] = frontier.numel()
    res["edges"] = sum(b[2].numel() for b in blocks)
    # empty-stream sync cost and an empty kernel-free API call
    t0 = time.perf_counter()
    for i in range(200):
        torch.cuda.current_stream().synchronize()
    res["idle_stream_sync_us"] = (time.perf_counter() - t0) / 200 * 1e6
    t0 = time.perf_counter()
    for i in range(200):
        torch.empty(1000, dtype=torch.int64, device=dev)
    res["torch_empty_us"] = (time.perf_counter() - t0) / 200 * 1e6
    t0 = time.perf_counter()
    for i in range(200):
        arena[10:20]
    res["torch_slice_us"] = (time.perf_counter() - t0) / 200 * 1e6
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
