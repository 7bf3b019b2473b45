#!/usr/bin/env python
"""bench.py - the mini-batch data path of BASELINE.json on B200.

A step = one mini-batch: multi-hop fan-out sampling (+ relabel) of `--batch` seeds and the feature
extract of the input frontier, on synthetic graphs of the named shapes (dist-gnn_b200/dgs_synth.py).

  python bench.py [--gpus N] [--steps K] [--warmup W]        this repo (sm_100a kernels)
  python bench.py --impl reference ...                       CPU baseline arm (DGL-semantics
                                                             sample_neighbors + index_select port,
                                                             oracle/dgs_oracle.c, all host threads)

N = 1  : BASELINE configs[1] - products-shaped graph resident in HBM, uniform [15,10,5], batch
         1024, un-cached path (CSRSampler + _CAPI_cuda_index_select).
N > 1  : the same shape sharded over the N GPUs of the box (node n lives on GPU n mod N: CSR rows
         and feature rows), every rank samples its own seed batches through P2PCacheSampler /
         P2PCacheFeatureServer; remote rows are NVLink peer loads issued inside the kernels, no
         collective on the data path ("scaling": "weak").
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "dist-gnn_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "sampled_edges_per_sec"
UNIT = "edges/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="products", choices=["tiny", "small", "products"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fan-out", default="15,10,5")
    ap.add_argument("--bias", action="store_true")
    ap.add_argument("--extract-algo", type=int, default=0)
    ap.add_argument("--layout", default="policy", choices=["policy", "sharded"],
                    help="N > 1: where graph + feature rows live. policy = placement computed by "
                         "DistGNN.cache (the reference's selfish / selfless / auto model) for the free "
                         "HBM; sharded = node n on GPU n mod N, remote rows over NVLink")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(args, fan_out):
    import dgs_synth
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    return (f"{args.shape}-shaped synthetic graph ({N} nodes, ~{E} edges, {D}-dim "
            f"{str(dt).replace('torch.', '')} feats), {'biased' if args.bias else 'uniform'} fan-out "
            f"{fan_out}, batch {args.batch}")


# ---------------------------------------------------------------------------------- CPU arm
def cpu_baseline(args, fan_out, host_graph, seconds, steps=None, warmup=2):
    """DGL-semantics CPU sample_neighbors + relabel + index_select (oracle port), all host threads."""
    import numpy as np
    import oracle
    import dgs_synth
    indptr, indices, probs, feat = host_graph
    N = len(indptr) - 1
    runner = oracle.CpuBatchRunner(indptr, indices, probs, feat, args.batch, fan_out)
    nb = (steps or 256) + warmup
    seeds = dgs_synth.seed_batches(N, args.batch, nb, seed=99).numpy()
    for w in range(warmup):
        runner.run(seeds[w], w)
    edges = rows = done = 0
    t0 = time.perf_counter()
    while True:
        e, r = runner.run(seeds[warmup + done % (nb - warmup)], 1000 + done)
        edges += e
        rows += r
        done += 1
        el = time.perf_counter() - t0
        if steps is not None:
            if done >= steps:
                break
        elif el >= seconds:
            break
    row_bytes = feat.dtype.itemsize * feat.shape[1]
    return {"value": edges / el, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{done} batches of the same workload ({el:.1f} s), DGL-semantics "
                      "sample_neighbors + to_block relabel + index_select restated in C/OpenMP "
                      "(DGL itself is not installed)",
            "batches_per_sec": done / el, "extract_gbps": rows * (2 * row_bytes + 8) / el / 1e9,
            "ms_per_step": el / done * 1e3, "steps": done}


def host_graph_from(args, device):
    """Generate the graph once (on the GPU when there is one) and hand numpy copies to the CPU arm."""
    import dgs_synth
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    indptr, indices, probs = dgs_synth.make_csr(N, E, seed=0, device=device, weights=args.bias)
    feat = dgs_synth.make_features(N, D, dt, seed=0, device=device)
    return indptr, indices, probs, feat


def run_reference(args, fan_out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ip, ix, pr, ft = host_graph_from(args, dev)
    host = (ip.cpu().numpy(), ix.cpu().numpy(), pr.cpu().numpy() if pr is not None else None,
            ft.cpu().numpy())
    del ip, ix, pr, ft
    cb = cpu_baseline(args, fan_out, host, None, steps=args.steps, warmup=max(1, args.warmup))
    line = {
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps"],
        "warmup": max(1, args.warmup), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "impl": "reference", "config": {"workload": workload_name(args, fan_out)},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "batches_per_sec": cb["batches_per_sec"], "extract_gbps": cb["extract_gbps"],
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------- B200 arm
def run_b200(args, fan_out):
    import torch.distributed as dist
    import dgs
    import dgs_synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from DistGNN.dist import create_communicator
        create_communicator(world)

    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    K, W = args.steps, args.warmup
    ip, ix, pr, ft = host_graph_from(args, dev)
    row_bytes = D * ft.element_size()
    labels = (torch.arange(N, device=dev) % 47).to(torch.int64)
    host_graph = None
    if rank == 0 and not args.no_cpu_baseline:
        host_graph = (ip.cpu().numpy(), ix.cpu().numpy(), pr.cpu().numpy() if pr is not None else None,
                      ft.cpu().numpy())

    if world == 1:
        sampler = dgs.classes.CSRSampler(ip, ix, pr)

        def extract(nids):
            return dgs.ops._CAPI_cuda_index_select(ft, nids, args.extract_algo)
        layout = "graph + features resident in HBM, un-cached path"
    else:
        # The class API wants the whole graph as pinned CPU tensors (src/sampling/sampler.cc:64-86);
        # the products shape is small enough for that.
        ipc, ixc, ftc = ip.cpu().pin_memory(), ix.cpu().pin_memory(), ft.cpu().pin_memory()
        prc = pr.cpu().pin_memory() if pr is not None else torch.Tensor()
        del ip, ix, ft
        torch.cuda.empty_cache()
        shard_nids = torch.arange(rank, N, world, dtype=torch.int64)   # node n -> GPU n mod world
        policy = None
        if args.layout == "policy":
            # what the reference's training script does (example/graphsage/node_classification.py:
            # 73-167): heat of every node -> value per byte -> knapsack over the free device memory
            from DistGNN.cache import choose_cache_policy, get_available_memory, get_node_heat
            graph = {"indptr": ipc, "indices": ixc, "features": ftc}
            if pr is not None:
                graph["probs"] = prc
            sh, fh = get_node_heat(ipc, ixc, torch.arange(N), fan_out,
                                   probs=prc if pr is not None else None, mode="uva")
            free = get_available_memory(local_rank, 7 << 30)
            name, s_nids, f_nids = choose_cache_policy(graph, sh, fh, free, world,
                                                       probs="probs" if pr is not None else None)
            del sh, fh
            policy = {"chosen": name, "free_hbm_gb": free / 1e9,
                      "structure_nodes_cached": int(s_nids.numel()),
                      "feature_rows_cached": int(f_nids.numel())}
            s_nids, f_nids = torch.sort(s_nids.cpu())[0], torch.sort(f_nids.cpu())[0]
        else:
            s_nids = f_nids = shard_nids
        sampler = dgs.classes.P2PCacheSampler(ipc, ixc, prc, s_nids, rank)
        fserver = dgs.classes.P2PCacheFeatureServer(ftc, f_nids, rank)

        def extract(nids):
            return fserver._CAPI_get_feature(nids, args.extract_algo)
        if sampler._mod_world < 0 and fserver._mod_world < 0:
            layout = (f"placement by the cache policy ({policy['chosen']}): everything fits the free HBM, "
                      f"so every GPU holds a full replica (CSR + features); no remote reads")
        elif sampler._mod_world > 0:
            layout = f"CSR + features sharded nid mod {world} over {world} GPUs, NVLink peer loads in-kernel"
        else:
            layout = f"placement by the cache policy ({policy['chosen']}), location table, NVLink peer loads"

    # distinct seed batches per rank and per step
    seeds_all = dgs_synth.seed_batches(N, args.batch, 3 * (K + W) + 2, seed=rank)
    seeds_dev = seeds_all.to(dev)
    seeds_pin = seeds_all.pin_memory()

    def step_device(i):
        blocks = sampler._CAPI_sample_node_classifiction(seeds_dev[i], fan_out, False)
        x = extract(blocks[-1][1])
        return blocks, x

    lab_host = torch.empty(args.batch, dtype=torch.int64).pin_memory()

    def step_e2e(i):
        s = seeds_pin[i].to(dev, non_blocking=True)                      # H2D of the step's input
        blocks = sampler._CAPI_sample_node_classifiction(s, fan_out, False)
        x = extract(blocks[-1][1])
        lab = dgs.ops._CAPI_cuda_index_select(labels, s)                  # node_classification.py:228
        lab_host.copy_(lab, non_blocking=True)                           # D2H of the step's result
        torch.cuda.current_stream().synchronize()
        return blocks, x

    loader = dgs.classes.BatchLoader(sampler, ft if world == 1 else fserver, labels)

    def step_fused(i):
        # extension: same batch through ONE call / one host round trip (dgs.classes.BatchLoader)
        blocks, x, _ = loader.load(seeds_pin[i], fan_out, False, algo=args.extract_algo,
                                   labels_out=lab_host)
        return blocks, x

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing ("value")
    for i in range(2):   # set-up (allocator pools, lazily enabled peer mappings) - not a timed step
        step_device(i)
        step_e2e(i)
        step_fused(i)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for i in range(W):
        step_device(i)
    barrier()
    launches0 = dgs.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    edges = rows = hop_seeds = uniq = 0
    e0.record()
    for i in range(W, W + K):
        blocks, x = step_device(i)
        edges += sum(b[2].numel() for b in blocks)
        hop_seeds += sum(b[0].numel() for b in blocks)
        uniq += sum(b[1].numel() for b in blocks)
        rows += x.shape[0]
    e1.record()
    barrier()
    launches = dgs.launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    total_edges = sum_over_ranks(edges)
    total_rows = sum_over_ranks(rows)
    # ---- roofline of the dominant kernel (the extract gather): the K launches of the timed region
    # are re-issued back to back on the same stream between two CUDA events, so the figure is the
    # kernel's own average duration (no host gaps); every launch gathers a different random
    # 75+ MB row set out of the 0.98 GB table into its own output buffer.
    # (the frontiers are re-sampled here and copied out of their arenas: holding arena views inside
    # the timed loop would force a fresh cudaMalloc per step)
    frontiers = [sampler._CAPI_sample_node_classifiction(seeds_dev[W + i], fan_out, False)[-1][1].clone()
                 for i in range(K)]
    outs = [extract(f) for f in frontiers]   # untimed pass: the caching allocator now owns K outputs
    del outs
    barrier()
    x0 = torch.cuda.Event(enable_timing=True)
    x1 = torch.cuda.Event(enable_timing=True)
    outs = []
    x0.record()
    for f in frontiers:
        outs.append(extract(f))
    x1.record()
    torch.cuda.synchronize()
    ex_ms = x0.elapsed_time(x1)
    ex_bytes = sum(f.numel() * (2 * row_bytes + 8) for f in frontiers)
    del outs
    # same for the sampling kernel (one cooperative launch per batch): K batches back to back
    pipe = sampler._pipe
    keep = [pipe.enqueue_only(seeds_dev[i], fan_out) for i in range(min(3, K))]
    barrier()
    x0.record()
    keep = [pipe.enqueue_only(seeds_dev[W + i], fan_out, rng_seed=i + 1) for i in range(K)]
    x1.record()
    torch.cuda.synchronize()
    sm_ms = x0.elapsed_time(x1)
    del keep
    # SURVEY 8d, summed over hops: sampling 24 (S + nnz) + relabel 8 (S + nnz) + 32 nnz + 8 U
    sm_bytes = 24.0 * (hop_seeds + edges) + 8.0 * (hop_seeds + edges) + 32.0 * edges + 8.0 * uniq

    # ---- end-to-end timing through the plugin API with host seeds / host result ("e2e")
    for i in range(W):
        step_e2e(K + W + i)
    barrier()
    t_edges = 0
    e0.record()
    for i in range(K + 2 * W, 2 * K + 2 * W):
        blocks, x = step_e2e(i)
        t_edges += sum(b[2].numel() for b in blocks)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_edges = sum_over_ranks(t_edges)

    # ---- the same end-to-end step through the one-call BatchLoader extension ("e2e_fused")
    for i in range(W):
        step_fused(2 * K + 2 * W + i)
    barrier()
    f_edges = 0
    e0.record()
    for i in range(2 * K + 3 * W, 3 * K + 3 * W):
        blocks, x = step_fused(i)
        f_edges += sum(b[2].numel() for b in blocks)
    e1.record()
    barrier()
    ms_fused = max_over_ranks(e0.elapsed_time(e1))
    fused_edges = sum_over_ranks(f_edges)
    clk = clocks.stop() if rank == 0 else None   # sampled from the warm-up through both timed regions

    # ---- N > 1, policy placement: the same steps over the modulo-sharded layout as well (north star
    # item 4: remote CSR rows and feature rows read by NVLink peer loads inside the kernels)
    sharded = None
    if world > 1 and args.layout == "policy":
        sampler2 = dgs.classes.P2PCacheSampler(ipc, ixc, prc, shard_nids, rank)
        fserver2 = dgs.classes.P2PCacheFeatureServer(ftc, shard_nids, rank)
        for i in range(W + 2):
            b2 = sampler2._CAPI_sample_node_classifiction(seeds_dev[i], fan_out, False)
            fserver2._CAPI_get_feature(b2[-1][1], args.extract_algo)
        barrier()
        s_edges = s_rows = 0
        e0.record()
        for i in range(W, W + K):
            b2 = sampler2._CAPI_sample_node_classifiction(seeds_dev[i], fan_out, False)
            x2 = fserver2._CAPI_get_feature(b2[-1][1], args.extract_algo)
            s_edges += sum(b[2].numel() for b in b2)
            s_rows += x2.shape[0]
        e1.record()
        barrier()
        ms2 = max_over_ranks(e0.elapsed_time(e1))
        fr2 = [f.clone() for f in frontiers]
        outs = [fserver2._CAPI_get_feature(f, args.extract_algo) for f in fr2]
        del outs
        barrier()
        outs = []
        x0.record()
        for f in fr2:
            outs.append(fserver2._CAPI_get_feature(f, args.extract_algo))
        x1.record()
        torch.cuda.synchronize()
        ex2_ms = max_over_ranks(x0.elapsed_time(x1))
        del outs
        tot_e2, tot_r2 = sum_over_ranks(s_edges), sum_over_ranks(s_rows)
        remote = sum(f.numel() for f in fr2) * row_bytes * (world - 1) / world
        sharded = {"layout": f"CSR + features sharded nid mod {world}, NVLink peer loads in-kernel",
                   "value": tot_e2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / K,
                   "batches_per_sec": world * K / (ms2 * 1e-3),
                   "extract_avg_launch_ms": ex2_ms / K,
                   "extract_peer_load_gbps_per_gpu": remote / (ex2_ms * 1e-3) / 1e9,
                   "nvlink_peak_gbps": 900}
        sampler2.close()
        fserver2.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    ex_gbs = ex_bytes / (ex_ms * 1e-3) / 1e9
    sm_gbs = sm_bytes / (sm_ms * 1e-3) / 1e9
    traffic = {}
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")))
        for name, key in (("gather_rows", "extract"), ("fused_batch", "sample")):
            d = prof[name][0]
            traffic[key] = (float(d["dram__bytes_read.sum"].split()[0]) +
                            float(d["dram__bytes_write.sum"].split()[0])) * 1e6
    except Exception:
        pass
    roof_extract = {"bound": "hbm", "kernel": "gather_rows_kernel (feature extract)",
                    "achieved": ex_gbs, "peak": peak, "unit": "GB/s", "frac": ex_gbs / peak,
                    "peak_source": peak_src, "traffic": traffic.get("extract"),
                    "algorithmic_bytes": "rows * (2 * row_bytes + 8)", "avg_launch_ms": ex_ms / K,
                    "rows_per_launch": rows / K}
    roof_sample = {"bound": "hbm", "kernel": "fused_batch_kernel (all hops: sample + relabel, one "
                                             "cooperative launch per batch)",
                   "achieved": sm_gbs, "peak": peak, "unit": "GB/s", "frac": sm_gbs / peak,
                   "peak_source": peak_src, "traffic": traffic.get("sample"),
                   "algorithmic_bytes": "sum over hops of 24 (S + nnz) [sample] + 8 (S + nnz) + 32 nnz + 8 U [relabel]",
                   "avg_launch_ms": sm_ms / K,
                   "note": "not bandwidth-bound at batch %d: limited by the issue rate of uncoalesced "
                           "accesses / random atomics and by 9 dependent phases with grid barriers "
                           "(DESIGN.md section 5, profiles/r01_rt_probe.txt)" % args.batch}
    dominant, other = (roof_sample, roof_extract) if sm_ms >= ex_ms else (roof_extract, roof_sample)
    dominant = dict(dominant)
    dominant["share_of_gpu_time"] = max(sm_ms, ex_ms) / (sm_ms + ex_ms)
    line = {
        "metric": METRIC, "value": total_edges / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args, fan_out), "layout": layout,
                   "l2": "inputs larger than L2 (feature table %.2f GB, CSR %.2f GB, fresh seeds "
                         "every step)" % (N * row_bytes / 1e9, (E * 8 + N * 8) / 1e9),
                   "extract_algo": args.extract_algo},
        "batches_per_sec": world * K / (ms * 1e-3),
        "extract_gbps": total_rows * (2 * row_bytes + 8) / (ms * 1e-3) / 1e9,
        "clocks": clk,
        "e2e": {"value": e2e_edges / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": args.batch * 8, "d2h_bytes_per_step": args.batch * 8 + 48,
                "ms_per_step": ms_e2e / K, "batches_per_sec": world * K / (ms_e2e * 1e-3),
                "note": "seeds from pinned host memory, blocks + features stay on the device (the "
                        "plugin API returns CUDA tensors), labels of the batch + hop sizes read back"},
        "placement_policy": policy if world > 1 else None,
        "sharded_layout": sharded,
        "e2e_fused": {"value": fused_edges / (ms_fused * 1e-3), "unit": UNIT,
                      "ms_per_step": ms_fused / K, "batches_per_sec": world * K / (ms_fused * 1e-3),
                      "note": "extension, not the reference-facing API: dgs.classes.BatchLoader enqueues "
                              "sample -> extract (frontier size read on the device) -> labels and "
                              "syncs once; same inputs, outputs and copies as e2e"},
        "gpu_launches": launches,
        "roofline": dominant,
        "roofline_other": other,
    }
    if host_graph is not None:
        cb = cpu_baseline(args, fan_out, host_graph, args.cpu_baseline_seconds)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["batches_per_sec"] = cb["batches_per_sec"]
        line["cpu_baseline"]["extract_gbps"] = cb["extract_gbps"]
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _OnlyJsonOnStdout:
    """Libraries (NCCL's version banner, torch warnings) sometimes write to fd 1; the contract is
    ONE JSON line on stdout, so fd 1 is pointed at stderr for the whole run and the JSON line is
    written to the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def emit(obj):
    OUT.emit(json.dumps(obj))


def main():
    global OUT
    args = parse_args()
    fan_out = [int(x) for x in args.fan_out.split(",")]
    with _OnlyJsonOnStdout() as out:
        OUT = out
        if args.impl == "reference":
            run_reference(args, fan_out)
        else:
            run_b200(args, fan_out)


if __name__ == "__main__":
    main()
