"""BASELINE configs 4 / 5: a large synthetic graph (papers100M- or friendster-shaped) whose CSR and
feature rows are sharded over the GPUs of one box (node n on GPU n % P), built per shard directly on
the device, sampled and extracted with in-kernel NVLink peer loads.  Launch with torchrun:

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 \
      tools/run_sharded.py --shape papers100M --batch 1024 --fan-out 15,10,5

Rank 0 prints one JSON object: sampled edges/s, extract GB/s (algorithmic), peer-load GB/s per GPU
(remote fraction (P-1)/P of the extracted row bytes / kernel time), batches/s, all summed over ranks
with the slowest rank's time.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dgs  # noqa: E402
import dgs_synth  # noqa: E402
from DistGNN.dist import create_communicator  # noqa: E402


def main():
    # only the JSON object goes to stdout: libraries (NCCL's version banner) write to fd 1 as well
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="papers100M")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fan-out", default="15,10,5")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--bias", action="store_true")
    ap.add_argument("--extract-rows", type=int, default=1_000_000)
    ap.add_argument("--replicate-hot", type=float, default=0.0,
                    help="also cache the feature rows of the top FRAC nodes by degree on EVERY GPU "
                         "(the reference's hot-node cache idea): local hits replace NVLink reads; the "
                         "location map is then the hash table (local copy wins)")
    args = ap.parse_args()
    fan = [int(x) for x in args.fan_out.split(",")]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        create_communicator(world)
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    t0 = time.time()
    nids, sp, si, spr = dgs_synth.make_shard(N, E, rank, world, device=dev, weights=args.bias)
    feat = dgs_synth.make_features(N, D, dt, device=dev, nids=nids)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    smp = dgs.classes.P2PCacheSampler.from_device_shards(sp, si, spr, nids, N, rank)
    hot_extra = 0
    if args.replicate_hot > 0 and world > 1:
        deg = dgs_synth.degrees(N, E, device=dev)
        hot = torch.topk(deg, int(N * args.replicate_hot)).indices
        del deg
        extra = hot[hot % world != rank]                 # hot nodes this rank does not own
        hot_extra = extra.numel()
        feat = torch.cat([feat, dgs_synth.make_features(N, D, dt, device=dev, nids=extra)])
        fnids = torch.cat([nids, extra.to(nids.dtype)])
        fs = dgs.classes.P2PCacheFeatureServer.from_device_shard(feat, fnids, N, rank)
    else:
        fs = dgs.classes.P2PCacheFeatureServer.from_device_shard(feat, nids, N, rank)
    del sp, si, spr, feat
    torch.cuda.empty_cache()
    t_build = time.time() - t0 - t_gen
    row_bytes = D * torch.empty(0, dtype=dt).element_size()
    seeds = dgs_synth.seed_batches(N, args.batch, args.steps + 5, seed=rank, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    for i in range(5):
        fs._CAPI_get_feature(smp._CAPI_sample_node_classifiction(seeds[i], fan)[-1][1])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = rows = 0
    t_s = t_x = 0.0
    evs = []
    e0.record()
    for i in range(args.steps):
        ta = time.perf_counter()
        blocks = smp._CAPI_sample_node_classifiction(seeds[5 + i], fan)
        tb = time.perf_counter()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        x = fs._CAPI_get_feature(blocks[-1][1])
        eb.record()
        evs.append((ea, eb))
        tc = time.perf_counter()
        t_s += tb - ta
        t_x += tc - tb
        edges += sum(b[2].numel() for b in blocks)
        rows += x.shape[0]
    e1.record()
    barrier()
    ms = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX)
    host_sample_us = reduce(t_s / args.steps * 1e6, dist.ReduceOp.MAX)
    host_extract_us = reduce(t_x / args.steps * 1e6, dist.ReduceOp.MAX)
    gpu_extract_us = reduce(sum(a.elapsed_time(b) for a, b in evs) / args.steps * 1e3, dist.ReduceOp.MAX)
    tot_edges, tot_rows = reduce(edges, dist.ReduceOp.SUM), reduce(rows, dist.ReduceOp.SUM)
    # host-side diagnostics of the extract call (max over ranks)
    fr = blocks[-1][1].clone()
    torch.cuda.synchronize()
    ta = time.perf_counter()
    for i in range(20):
        tmp = torch.empty((fr.numel(), D), dtype=dt, device=dev)
    diag_empty_us = reduce((time.perf_counter() - ta) / 20 * 1e6, dist.ReduceOp.MAX)
    from dgs import _lib
    from dgs._util import stream as _stream
    out = torch.empty((fr.numel(), D), dtype=dt, device=dev)
    torch.cuda.synchronize()
    e0.record()
    ta = time.perf_counter()
    for i in range(20):
        _lib.check(_lib.lib().dgs_extract_sharded(fs.gpu_features_._handle, fs._row_bytes, 1, fr.data_ptr(),
                                                  fr.numel(), out.data_ptr(), 0, _stream()))
    diag_call_us = reduce((time.perf_counter() - ta) / 20 * 1e6, dist.ReduceOp.MAX)
    e1.record()
    torch.cuda.synchronize()
    diag_gpu_us = reduce(e0.elapsed_time(e1) / 20 * 1e3, dist.ReduceOp.MAX)
    # extract-only microbench: R random ids per rank, back-to-back launches
    g = torch.Generator().manual_seed(rank)
    q = [torch.randint(0, N, (args.extract_rows,), generator=g).to(dev) for _ in range(4)]
    xms_algo = {}
    for algo in (1, 2):
        outs = [fs._CAPI_get_feature(q[i % 4], algo) for i in range(8)]   # untimed: allocator pools
        del outs
        barrier()
        e0.record()
        outs = [fs._CAPI_get_feature(q[i % 4], algo) for i in range(8)]
        e1.record()
        barrier()
        xms_algo[algo] = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX) / 8
        del outs
    xms = xms_algo[1]
    # sampling-only, back to back
    keep = [smp._pipe.enqueue_only(seeds[i], fan) for i in range(3)]
    barrier()
    e0.record()
    keep = [smp._pipe.enqueue_only(seeds[5 + i], fan, rng_seed=i + 1) for i in range(args.steps)]
    e1.record()
    barrier()
    sms = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX) / args.steps
    del keep
    if rank == 0:
        R = args.extract_rows
        os.write(json_fd, (json.dumps({
            "shape": args.shape, "n_gpus": world, "batch": args.batch, "fan_out": fan,
            "bias": args.bias, "replicate_hot": args.replicate_hot, "hot_rows_replicated_per_gpu": hot_extra,
            "loc_mode": "modulo" if fs._mod_world else "hash", "gen_s": t_gen, "build_s": t_build,
            "ms_per_step": ms / args.steps, "batches_per_sec": world * args.steps / (ms * 1e-3),
            "sampled_edges_per_sec": tot_edges / (ms * 1e-3),
            "extract_gbps_in_step": tot_rows * (2 * row_bytes + 8) / (ms * 1e-3) / 1e9,
            "rows_per_step_per_gpu": tot_rows / world / args.steps,
            "edges_per_step_per_gpu": tot_edges / world / args.steps,
            "sample_kernel_ms": sms, "host_sample_call_us_max": host_sample_us,
            "host_extract_call_us_max": host_extract_us, "gpu_extract_in_step_us_max": gpu_extract_us, "host_cpus": len(os.sched_getaffinity(0)),
            "diag": {"torch_empty_us": diag_empty_us, "extract_c_call_us": diag_call_us,
                     "extract_gpu_frontier_us": diag_gpu_us},
            "extract_only": {"rows_per_gpu": R, "ms": xms, "ms_tma": xms_algo[2],
                             "algorithmic_gbps_per_gpu": R * (2 * row_bytes + 8) / (xms * 1e-3) / 1e9,
                             "peer_load_gbps_per_gpu": R * row_bytes * (world - 1) / world / (xms * 1e-3) / 1e9,
                             "nvlink_peak_gbps": 900, "nvlink_measured_peer_copy_gbps": 770},
            "mem_allocated_gb": torch.cuda.max_memory_allocated() / 1e9}) + "\n").encode())
    if world > 1:
        dist.barrier()
    smp.close(barrier=False)
    fs.close(barrier=False)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
