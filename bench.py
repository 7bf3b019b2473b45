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
N > 1  : the north-star layout - the same shape SHARDED over the N GPUs of the box (node n lives on
         GPU n mod N: CSR rows and feature rows), every rank samples its own seed batches through
         P2PCacheSampler / P2PCacheFeatureServer; remote rows are NVLink peer loads issued inside
         the kernels, no collective on the data path ("scaling": "weak").  The full-replica layout
         the cache policy would pick on a 180 GB GPU is measured too and reported as the extra key
         `replica_layout` (--layout policy / replica make it the headline instead).
Every BASELINE config is reachable: --shape papers100M|friendster (shards are generated on their
owner GPU), --bias --fan-out 25,10 --cache-ratio r (config 3), --replicate-hot f.
Before anything is timed every rank checks one full-neighbour batch against the CPU oracle and one
extract against the closed-form feature rows ("parity_checked").
One JSON line is printed by rank 0.
"""
import argparse
import glob
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arm runs with every host core while the other ranks of a torchrun launch wait for it:
# OpenMP workers must not spin at their barriers when a core is taken by somebody else's wait loop
# (measured: 3.4 M instead of 52 M edges/s with the default active wait policy at N = 2).
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "dist-gnn_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "sampled_edges_per_sec"
UNIT = "edges/s"
BIG_SHAPES = ("papers100M", "friendster")   # never materialised on the host / on one rank's CPU


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="products",
                    choices=["tiny", "small", "products", "papers100M", "friendster"])
    ap.add_argument("--batch", type=int, default=None, help="default 1024 (4096 for friendster)")
    ap.add_argument("--fan-out", default=None,
                    help="default 15,10,5 (25,10 with --bias, 20,15,10 for friendster)")
    ap.add_argument("--bias", action="store_true", help="edge-weight biased sampling")
    ap.add_argument("--cache-ratio", type=float, default=None,
                    help="N = 1: cache the top r*N nodes by degree on the GPU (location table), read "
                         "the rest from pinned host memory; 0 = everything from pinned host memory; "
                         "default: graph resident in HBM, un-cached ops path")
    ap.add_argument("--extract-algo", type=int, default=0)
    ap.add_argument("--layout", default="sharded", choices=["sharded", "policy", "replica"],
                    help="N > 1: where graph + feature rows live.  sharded = node n on GPU n mod N, "
                         "remote rows over NVLink (north star); policy = placement computed by "
                         "DistGNN.cache for the free HBM; replica = everything on every GPU")
    ap.add_argument("--replicate-hot", type=float, default=0.0,
                    help="sharded layout: additionally cache the feature rows of the top FRAC nodes by "
                         "degree on every GPU (local hits replace NVLink reads; location table)")
    ap.add_argument("--prefetch", type=int, default=8,
                    help="mini-batches per launch of the pipelined leg (e2e_pipelined)")
    ap.add_argument("--overlap", action="store_true",
                    help="also time BatchLoader.iter_many (sampling of group g + 1 on the main stream "
                         "overlapped with the extracts of group g on a side stream): e2e_overlapped")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-extra-layout", action="store_true")
    ap.add_argument("--nvtx", action="store_true",
                    help="wrap the step's two plugin calls in the reference's NVTX ranges "
                         "('sampling', 'loading') for ncu --nvtx-include")
    ap.add_argument("--no-exchange", action="store_true",
                    help="N > 1: skip the NCCL id-exchange variant in the extract-only leg")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 4096 if args.shape == "friendster" else 1024
    if args.fan_out is None:
        args.fan_out = "20,15,10" if args.shape == "friendster" else ("25,10" if args.bias else "15,10,5")
    return args


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


def config_of(args, fan_out):
    """The `config` object - a pure function of the command line, so both arms print the same one."""
    import dgs_synth
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    row_bytes = D * torch.empty(0, dtype=dt).element_size()
    w = (f"{args.shape}-shaped synthetic graph ({N} nodes, ~{E} edges, {D}-dim "
         f"{str(dt).replace('torch.', '')} feats), {'biased' if args.bias else 'uniform'} fan-out "
         f"{fan_out}, batch {args.batch}")
    if args.cache_ratio is not None:
        w += f", GPU cache of the top {args.cache_ratio:g} of the nodes by degree, rest in pinned host memory"
    return {"workload": w,
            "l2": "inputs larger than L2 (feature table %.2f GB, CSR %.2f GB, fresh seeds every step)"
                  % (N * row_bytes / 1e9, (E * 8 + N * 8) / 1e9)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------- CPU arm
def cpu_baseline(args, fan_out, host_graph, seconds, steps=None, warmup=2):
    """DGL-semantics CPU sample_neighbors + relabel + index_select (oracle port), all host threads
    (asked for explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    import oracle
    import dgs_synth
    oracle.set_num_threads(0)
    torch.set_num_threads(host_cores())
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


def full_graph(args, device):
    """The whole graph on one device (small shapes) - also the source of the CPU arm's numpy copies."""
    import dgs_synth
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    indptr, indices, probs = dgs_synth.make_csr(N, E, seed=0, device=device, weights=args.bias)
    feat = dgs_synth.make_features(N, D, dt, seed=0, device=device)
    return indptr, indices, probs, feat


def to_host_numpy(ip, ix, pr, ft):
    f = ft.cpu()
    if f.dtype == torch.bfloat16:      # numpy has no bf16: rows are moved as bytes anyway
        f = f.view(torch.int16)
    return (ip.cpu().numpy(), ix.cpu().numpy(), pr.cpu().numpy() if pr is not None else None, f.numpy())


def host_ram_ok(args):
    import dgs_synth
    N, E, D, dt = dgs_synth.SHAPES[args.shape]
    need = N * D * torch.empty(0, dtype=dt).element_size() + E * (12 if args.bias else 8) + N * 8
    try:
        import psutil
        return psutil.virtual_memory().available > 1.5 * need, need
    except Exception:
        return False, need


def run_reference(args, fan_out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.shape in BIG_SHAPES:
        ok, need = host_ram_ok(args)
        if not ok:
            emit({"impl": "reference", "unavailable": f"{args.shape}: the CPU arm needs the whole graph "
                                                     f"({need / 1e9:.0f} GB) in host memory"})
            return
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    host = to_host_numpy(*full_graph(args, dev))
    if dev == "cuda":
        torch.cuda.empty_cache()
    cb = cpu_baseline(args, fan_out, host, None, steps=args.steps, warmup=max(1, args.warmup))
    line = {
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps"],
        "warmup": max(1, args.warmup), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "impl": "reference", "config": config_of(args, fan_out),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "batches_per_sec": cb["batches_per_sec"], "extract_gbps": cb["extract_gbps"],
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------- reference kernels
def load_reference_module():
    """oracle/_ref: the UNMODIFIED reference compiled for sm_100a (oracle/build_ref.sh).  Its module
    is also called `dgs`, so this repo's package is hidden from sys.modules while it loads."""
    cands = glob.glob(os.path.join(ROOT, "oracle", "_ref", "dgs.cpython-*.so"))
    if not cands:
        raise RuntimeError("oracle/_ref not built (oracle/build_ref.sh needs /root/reference)")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "dgs" or k.startswith("dgs.")}
    try:
        spec = importlib.util.spec_from_file_location("dgs", cands[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in list(sys.modules):
            if k == "dgs" or k.startswith("dgs."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return mod


def reference_gpu_leg(args, fan_out, ipc, ixc, prc, ftc, seeds_dev, steps, dev):
    """The reference's own kernels on this B200, same inputs, same plugin calls (P2PCacheSampler +
    P2PCacheFeatureServer with every node cached on the GPU = the best case for the reference).
    Outside every timed region of this repo's arm; reported like cpu_baseline."""
    ref = load_reference_module()
    ref.ops._CAPI_set_nccl(1, ref.ops._CAPI_get_unique_id(), 0)
    N = ipc.numel() - 1
    allnodes = torch.arange(N)
    smp = ref.classes.P2PCacheSampler(ipc, ixc, prc, allnodes, 0)
    fs = ref.classes.P2PCacheFeatureServer(ftc, allnodes.to(dev), 0)
    for i in range(3):
        b = smp._CAPI_sample_node_classifiction(seeds_dev[i], fan_out, False)
        fs._CAPI_get_feature(b[-1][1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = rows = 0
    e0.record()
    for i in range(steps):
        b = smp._CAPI_sample_node_classifiction(seeds_dev[3 + i], fan_out, False)
        x = fs._CAPI_get_feature(b[-1][1])
        edges += sum(t[2].numel() for t in b)
        rows += x.shape[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    row_bytes = ftc.shape[1] * ftc.element_size()
    del smp, fs
    torch.cuda.empty_cache()
    return {"value": edges / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "batches_per_sec": steps / (ms * 1e-3),
            "extract_gbps": rows * (2 * row_bytes + 8) / (ms * 1e-3) / 1e9,
            "what": "oracle/_ref = the unmodified reference sources compiled for sm_100a, same B200, "
                    "same seeds; sampler._CAPI_sample_node_classifiction + feature_server._CAPI_get_"
                    "feature with every node cached on the GPU (src/sampling/sampler.cc:146-166, "
                    "src/feature/cuda/feature_ops.cu:75-138)"}


# ---------------------------------------------------------------------------------- parity (untimed)
def parity_check(args, sampler, extract, N, D, dt, dev, rank, host_csr, shard_info):
    """One full-neighbour batch against the CPU oracle and one extract against the closed-form
    feature rows, on every rank, before anything is timed."""
    import numpy as np
    import oracle
    import dgs_synth
    g = torch.Generator().manual_seed(4242 + rank)
    seeds = torch.randint(0, N, (48,), generator=g).unique()
    if host_csr is not None:
        ip, ix = host_csr
        hops = 2
    else:
        # big shapes: rebuild the CSR rows of the seeds from the closed-form generators and hand the
        # oracle a CSR in which only those rows are populated (one hop)
        E = dgs_synth.SHAPES[args.shape][1]
        deg = dgs_synth.degrees(N, E, device=dev)
        gptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=gptr[1:])
        order = torch.sort(seeds).values.to(dev)
        d = deg[order]
        starts = gptr[order]
        sparse = torch.zeros(N, dtype=torch.int64, device=dev)
        sparse[order] = d
        ipd = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(sparse, 0, out=ipd[1:])
        rows = [dgs_synth.edge_targets(N, 0, int(s), int(c), device=dev) for s, c in
                zip(starts.tolist(), d.tolist())]
        ix = torch.cat(rows).cpu().numpy() if rows else np.zeros(0, np.int64)
        ip = ipd.cpu().numpy()
        del deg, gptr, sparse, ipd
        hops = 1
    exp = oracle.sample_blocks_all_neighbors(seeds.numpy(), ip, ix, hops)
    got = sampler._CAPI_sample_node_classifiction(seeds.to(dev), [-1] * hops, False)
    ok = len(got) == len(exp)
    for a, e in zip(got, exp):
        for x, z in zip(a, e):
            ok = ok and np.array_equal(x.cpu().numpy(), z)
    q = torch.randint(0, N, (20000,), generator=g).to(dev)
    x = extract(q)
    ok = ok and torch.equal(x, dgs_synth.feature_rows(q, D, dt))
    return bool(ok)


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
    row_bytes = D * torch.empty(0, dtype=dt).element_size()
    big = args.shape in BIG_SHAPES
    labels = (torch.arange(N, device=dev) % 47).to(torch.int64)
    host_graph = host_csr = None
    policy = None
    closers = []
    ipc = ixc = prc = ftc = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM,
                               "min": dist.ReduceOp.MIN}[op])
        return float(t.item())

    # ------------------------------------------------------------------ data + samplers
    feature_source = None     # tensor or P2PCacheFeatureServer handed to BatchLoader
    if big:
        # shards are generated on their owner GPU; nothing whole-graph ever exists on the host
        t0 = time.time()
        nids, sp, si, spr = dgs_synth.make_shard(N, E, rank, world, device=dev, weights=args.bias)
        feat = dgs_synth.make_features(N, D, dt, device=dev, nids=nids)
        torch.cuda.synchronize()
        sampler = dgs.classes.P2PCacheSampler.from_device_shards(sp, si, spr, nids, N, rank)
        fnids = nids
        if args.replicate_hot > 0 and world > 1:
            deg = dgs_synth.degrees(N, E, device=dev)
            hot = torch.topk(deg, int(N * args.replicate_hot)).indices
            del deg
            extra = hot[hot % world != rank]
            feat = torch.cat([feat, dgs_synth.make_features(N, D, dt, device=dev, nids=extra)])
            fnids = torch.cat([nids, extra.to(nids.dtype)])
        fserver = dgs.classes.P2PCacheFeatureServer.from_device_shard(feat, fnids, N, rank)
        del sp, si, spr, feat
        torch.cuda.empty_cache()
        closers += [sampler, fserver]
        feature_source = fserver
        layout = (f"CSR + features sharded nid mod {world} over {world} GPU(s), shards generated on "
                  f"the device ({time.time() - t0:.0f} s)" +
                  (", NVLink peer loads in-kernel" if world > 1 else ""))
        if args.replicate_hot > 0 and world > 1:
            layout += f", feature rows of the top {args.replicate_hot:g} nodes by degree replicated"

        def extract(nids_):
            return fserver._CAPI_get_feature(nids_, args.extract_algo)
    else:
        ip, ix, pr, ft = full_graph(args, dev)
        need_host = (rank == 0 and not args.no_cpu_baseline) or world > 1 or args.cache_ratio is not None
        if need_host:
            ipc, ixc, ftc = ip.cpu().pin_memory(), ix.cpu().pin_memory(), ft.cpu().pin_memory()
            prc = pr.cpu().pin_memory() if pr is not None else torch.Tensor()
            host_csr = (ipc.numpy(), ixc.numpy())
            if rank == 0 and not args.no_cpu_baseline:
                f = ftc.view(torch.int16) if ftc.dtype == torch.bfloat16 else ftc
                host_graph = (ipc.numpy(), ixc.numpy(), prc.numpy() if pr is not None else None, f.numpy())
        else:
            host_csr = (ip.cpu().numpy(), ix.cpu().numpy())
        if world == 1 and args.cache_ratio is None:
            sampler = dgs.classes.CSRSampler(ip, ix, pr)
            feature_source = ft
            layout = "graph + features resident in HBM, un-cached path"

            def extract(nids_):
                return dgs.ops._CAPI_cuda_index_select(ft, nids_, args.extract_algo)
        elif world == 1:
            del ip, ix, ft
            torch.cuda.empty_cache()
            if args.cache_ratio <= 0:
                sampler = dgs.classes.CSRSampler(ipc, ixc, prc if pr is not None else None, device=dev)
                feature_source = ftc
                layout = "graph + features in pinned host memory (UVA reads over PCIe), nothing cached"

                def extract(nids_):
                    return dgs.ops._CAPI_cuda_index_select(ftc, nids_, args.extract_algo)
            else:
                from DistGNN.cache import get_cache_nids_by_degree
                cache = get_cache_nids_by_degree(ipc, args.cache_ratio)
                sampler = dgs.classes.P2PCacheSampler(ipc, ixc, prc, cache, 0)
                fserver = dgs.classes.P2PCacheFeatureServer(ftc, cache, 0)
                closers += [sampler, fserver]
                feature_source = fserver
                layout = (f"top {args.cache_ratio:g} of the nodes by degree ({cache.numel()}) cached in HBM "
                          f"behind the location table, the rest read from pinned host memory")

                def extract(nids_):
                    return fserver._CAPI_get_feature(nids_, args.extract_algo)
        else:
            # The class API wants the whole graph as pinned CPU tensors (src/sampling/sampler.cc:64-86)
            del ip, ix, ft
            torch.cuda.empty_cache()
            shard_nids = torch.arange(rank, N, world, dtype=torch.int64)   # node n -> GPU n mod world
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
            elif args.layout == "replica":
                s_nids = f_nids = torch.arange(N, dtype=torch.int64)
            else:
                s_nids = f_nids = shard_nids
                if args.replicate_hot > 0:
                    from DistGNN.cache import get_cache_nids_by_degree
                    hot = get_cache_nids_by_degree(ipc, args.replicate_hot)
                    f_nids = torch.cat([shard_nids, hot[hot % world != rank]])
            sampler = dgs.classes.P2PCacheSampler(ipc, ixc, prc, s_nids, rank)
            fserver = dgs.classes.P2PCacheFeatureServer(ftc, f_nids, rank)
            closers += [sampler, fserver]
            feature_source = fserver

            def extract(nids_):
                return fserver._CAPI_get_feature(nids_, args.extract_algo)
            if sampler._mod_world < 0 and fserver._mod_world < 0:
                layout = ("every GPU holds a full replica (CSR + features)" +
                          (f" - placement by the cache policy ({policy['chosen']}): everything fits the "
                           f"free HBM" if policy else "") + "; no remote reads")
            elif sampler._mod_world > 0:
                layout = f"CSR + features sharded nid mod {world} over {world} GPUs, NVLink peer loads in-kernel"
                if args.replicate_hot > 0:
                    layout += (f", feature rows of the top {args.replicate_hot:g} nodes by degree "
                               f"replicated on every GPU (location table)")
            else:
                layout = f"placement by the cache policy ({policy['chosen']}), location table, NVLink peer loads"

    # ------------------------------------------------------------------ parity before timing
    ok = parity_check(args, sampler, extract, N, D, dt, dev, rank, host_csr, None)
    parity_checked = reduce(1.0 if ok else 0.0, "min") == 1.0
    assert parity_checked, "parity check failed: blocks / extract differ from the oracle"

    # distinct seed batches per rank and per step
    n_batches = 3 * (K + W) + 2 + max(1, args.prefetch) * (K + W + 2)
    seeds_all = dgs_synth.seed_batches(N, args.batch, n_batches, seed=rank)
    seeds_dev = seeds_all.to(dev)
    seeds_pin = seeds_all.pin_memory()

    if args.nvtx:
        # the reference's ranges (scripts/ncu_sampling.py:38,48, scripts/ncu_feature.py:17-20), for
        #   ncu --nvtx --nvtx-include "sampling/" ... / --nvtx-include "loading/" ...
        rng_push, rng_pop = torch.cuda.nvtx.range_push, torch.cuda.nvtx.range_pop
    else:
        rng_push = rng_pop = lambda *a: None

    def step_device(i):
        rng_push("sampling")
        blocks = sampler._CAPI_sample_node_classifiction(seeds_dev[i], fan_out, False)
        rng_pop()
        rng_push("loading")
        x = extract(blocks[-1][1])
        rng_pop()
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

    loader = dgs.classes.BatchLoader(sampler, feature_source, labels)

    def step_fused(i):
        # extension: same batch through ONE call / one host round trip (dgs.classes.BatchLoader)
        blocks, x, _ = loader.load(seeds_pin[i], fan_out, False, algo=args.extract_algo,
                                   labels_out=lab_host)
        return blocks, x

    def timed(step, first, count):
        """`count` steps between two CUDA events + barriers; max over ranks, sums over ranks."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        edges = rows = hop_seeds = uniq = 0
        barrier()
        e0.record()
        for i in range(first, first + count):
            blocks, x = step(i)
            edges += sum(b[2].numel() for b in blocks)
            hop_seeds += sum(b[0].numel() for b in blocks)
            uniq += sum(b[1].numel() for b in blocks)
            rows += x.shape[0]
        e1.record()
        barrier()
        return {"ms": reduce(e0.elapsed_time(e1), "max"), "edges": reduce(edges, "sum"),
                "rows": reduce(rows, "sum"), "local": (edges, rows, hop_seeds, uniq)}

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
    launches0 = dgs.launch_count()
    t_dev = timed(step_device, W, K)
    launches = dgs.launch_count() - launches0
    ms = t_dev["ms"]
    edges, rows, hop_seeds, uniq = t_dev["local"]

    # ---- roofline of the two kernels: the K launches of the timed region are re-issued back to back
    # on the same stream between two CUDA events, so the figure is the kernel's own average duration
    # (no host gaps); every launch gathers a different random 75+ MB row set out of the table into
    # its own output buffer.  (the frontiers are re-sampled here and copied out of their arenas:
    # holding arena views inside the timed loop would force a fresh cudaMalloc per step)
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
    ex_ms_max = reduce(ex_ms, "max")
    ex_rows = sum(f.numel() for f in frontiers)
    ex_bytes = ex_rows * (2 * row_bytes + 8)
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
    # (+ 4 bytes per weight of every sampled row when biased: not counted - conservative)
    sm_bytes = 24.0 * (hop_seeds + edges) + 8.0 * (hop_seeds + edges) + 32.0 * edges + 8.0 * uniq

    # ---- end-to-end timing through the plugin API with host seeds / host result ("e2e")
    for i in range(W):
        step_e2e(K + W + i)
    t_e2e = timed(step_e2e, K + 2 * W, K)

    # ---- the same end-to-end step through the one-call BatchLoader extension ("e2e_fused")
    for i in range(W):
        step_fused(2 * K + 2 * W + i)
    t_fused = timed(step_fused, 2 * K + 3 * W, K)

    # ---- B mini-batches per launch ("e2e_pipelined"): the loader takes B seed batches at once
    pipelined = None
    B = args.prefetch
    many_ok = B > 1 and pipe._plan_many(B, args.batch, fan_out)["ws"] is not None
    if many_ok:
        base = 3 * (K + W) + 2
        groups = [seeds_pin[base + j * B: base + (j + 1) * B] for j in range(K + W + 2)]

        def step_many(j):
            res = loader.load_many(groups[j], fan_out, False, algo=args.extract_algo)
            blocks = [b for r in res for b in r[0]]
            x = torch.empty(sum(r[1].shape[0] for r in res), 0)
            return blocks, x
        for j in range(W + 2):
            step_many(j)
        t_many = timed(step_many, W + 2, K)
        pipelined = {"value": t_many["edges"] / (t_many["ms"] * 1e-3), "unit": UNIT,
                     "batches_per_launch": B, "ms_per_batch": t_many["ms"] / (K * B),
                     "batches_per_sec": world * K * B / (t_many["ms"] * 1e-3),
                     "extract_gbps": t_many["rows"] * (2 * row_bytes + 8) / (t_many["ms"] * 1e-3) / 1e9,
                     "note": "extension: BatchLoader.load_many - B seed batches from pinned host memory "
                             "sampled by ONE cooperative launch (the hops of all B batches share every "
                             "phase and grid barrier), then B extracts; results bit-identical to B "
                             "single calls with the same RNG seeds"}
        if args.overlap:
            # N > 1: the NVLink-bound gathers keep the link busy with 2 CTAs per SM, which leaves the
            # rest of every SM to the next group's sampling kernels (measured at 2 GPUs: 0.114 -> 0.095
            # ms per batch; without the cap 0.108; on one GPU the HBM-bound gather and the
            # latency-bound sampling only slow each other down, so no cap and no gain there)
            gc = int(os.environ.get("DGS_BENCH_GATHER_CTAS", "0")) or (2 if world > 1 else None)
            for _ in loader.iter_many(groups[:W + 2], fan_out, False, None, args.extract_algo, gc):
                pass
            barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            o_edges = 0
            e0.record()
            for res in loader.iter_many(groups[W + 2:W + 2 + K], fan_out, False, None, args.extract_algo, gc):
                o_edges += sum(b[2].numel() for r in res for b in r[0])
            e1.record()
            barrier()
            o_ms = reduce(e0.elapsed_time(e1), "max")
            pipelined["overlapped"] = {"ms_per_batch": o_ms / (K * B),
                                       "value": reduce(o_edges, "sum") / (o_ms * 1e-3), "unit": UNIT,
                                       "note": "BatchLoader.iter_many: two streams, the next group's "
                                               "sampling overlaps this group's extracts"}
        # the multi-batch sampling kernel alone, back to back
        sd = [g.to(dev) for g in groups[:K + 2]]
        for j in range(2):
            keep = loader.enqueue_many_only(sd[j], fan_out)
        barrier()
        x0.record()
        for j in range(K):     # (each arena is released as the next one is allocated: stream-ordered reuse)
            keep = loader.enqueue_many_only(sd[2 + j], fan_out)
        x1.record()
        torch.cuda.synchronize()
        pipelined["sample_kernel_ms_per_batch"] = x0.elapsed_time(x1) / (K * B)
        pipelined["sample_kernel_frac_of_hbm"] = (sm_bytes / K) / (pipelined["sample_kernel_ms_per_batch"]
                                                                  * 1e-3) / 1e9 / peaks()[0]
        pipelined["sample_kernels"] = ("mb_pick_kernel / mb_rank_kernel / mb_emit_kernel: one launch per "
                                       "phase and hop, shared by the B batches (3 L launches)")
        del keep, sd
    # ---- N > 1: extract-only leg, 1 M random rows per GPU on all GPUs at once (the NVLink roofline
    # at a launch size that is not dominated by the fixed cost of a mini-batch-sized launch)
    extract_only = None
    if world > 1 and getattr(feature_source, "_mod_world", -1) >= 0:
        R = 1_000_000
        g = torch.Generator().manual_seed(777 + rank)
        qs = [torch.randint(0, N, (R,), generator=g).to(dev) for _ in range(2)]
        outs = [extract(qs[i % 2]) for i in range(4)]
        del outs
        barrier()
        x0.record()
        outs = [extract(qs[i % 2]) for i in range(6)]
        x1.record()
        barrier()
        xo_ms = reduce(x0.elapsed_time(x1), "max") / 6
        del outs
        payload = R * row_bytes * (world - 1) / world / (xo_ms * 1e-3) / 1e9
        extract_only = {"rows_per_gpu": R, "ms": xo_ms, "peer_payload_gbps_per_gpu": payload,
                        "algorithmic_gbps_per_gpu": R * (2 * row_bytes + 8) / (xo_ms * 1e-3) / 1e9}
        # the north star's alternative, measured on the same requests: the ids travel to the owners
        # over NCCL (all-to-all), the owners gather from their own shard, a second all-to-all brings
        # the rows back (P2PCacheFeatureServer.get_feature_exchange) - against in-kernel peer loads
        if getattr(feature_source, "_mod_world", -1) > 0 and hasattr(feature_source, "get_feature_exchange") \
                and not args.no_exchange:
            ex = feature_source.get_feature_exchange(qs[0])
            same = torch.equal(ex, extract(qs[0]))
            assert reduce(1.0 if same else 0.0, "min") == 1.0, "id-exchange extract differs from the peer-load extract"
            del ex
            feature_source.get_feature_exchange(qs[1])
            barrier()
            x0.record()
            for i in range(4):
                keep = feature_source.get_feature_exchange(qs[i % 2])
            x1.record()
            barrier()
            ex_ms = reduce(x0.elapsed_time(x1), "max") / 4
            del keep
            extract_only.update({
                "exchange_ms": ex_ms,
                "exchange_payload_gbps_per_gpu": R * row_bytes * (world - 1) / world / (ex_ms * 1e-3) / 1e9,
                "exchange_vs_peer_loads": xo_ms / ex_ms,
                "exchange_what": "route kernel + all_gather(counts) + host read + NCCL all-to-all(ids) + "
                                 "local gather + NCCL all-to-all(rows) + gather back to request order; "
                                 "rows equal to the peer-load extract (asserted)"})
        del qs
    clk = clocks.stop() if rank == 0 else None   # sampled from the warm-up through the timed regions

    # ---- N > 1, sharded headline: the full-replica layout (what the cache policy picks when
    # everything fits the HBM of every GPU) in the same run, as an extra key
    extra_layout = None
    if world > 1 and not big and args.layout == "sharded" and not args.no_extra_layout:
        everything = torch.arange(N, dtype=torch.int64)
        sampler2 = dgs.classes.P2PCacheSampler(ipc, ixc, prc, everything, rank)
        fserver2 = dgs.classes.P2PCacheFeatureServer(ftc, everything, rank)

        def step2(i):
            b2 = sampler2._CAPI_sample_node_classifiction(seeds_dev[i], fan_out, False)
            return b2, fserver2._CAPI_get_feature(b2[-1][1], args.extract_algo)
        for i in range(W + 2):
            step2(i)
        t2 = timed(step2, W, K)
        extra_layout = {"layout": "full replica of CSR + features on every GPU, no remote reads",
                        "value": t2["edges"] / (t2["ms"] * 1e-3), "unit": UNIT, "ms_per_step": t2["ms"] / K,
                        "batches_per_sec": world * K / (t2["ms"] * 1e-3)}
        sampler2.close()
        fserver2.close()

    # ---- the reference's own kernels on this GPU (N = 1, small shapes), outside the timed regions
    ref_gpu = None
    if world == 1 and not big and not args.no_reference_gpu and args.cache_ratio is None:
        try:
            if ipc is None:
                ipc, ixc, ftc = host_csr_to_pinned(host_csr, ft)
                prc = pr.cpu().pin_memory() if pr is not None else torch.Tensor()
            ref_gpu = reference_gpu_leg(args, fan_out, ipc, ixc, prc, ftc, seeds_dev, min(K, 20), dev)
        except Exception as e:   # not built / not loadable on this box: say so, do not fail the bench
            ref_gpu = {"unavailable": repr(e)[:300]}

    def quiet_barrier():
        """End-of-run barrier that does not burn a host core while rank 0 times the CPU baseline."""
        torch.cuda.synchronize()
        w = dist.barrier(async_op=True)
        while not w.is_completed():
            time.sleep(0.05)

    if rank != 0:
        if world > 1:
            quiet_barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    remote_frac = 0.0
    if world > 1 and feature_source is not None and getattr(feature_source, "_mod_world", -1) >= 0:
        remote_frac = (world - 1) / world
    ex_gbs = ex_bytes / (ex_ms * 1e-3) / 1e9
    sm_gbs = sm_bytes / (sm_ms * 1e-3) / 1e9
    traffic, traffic_src = {}, None
    for prof_name in ("r02_ncu_full_summary.json", "r01_ncu_full_summary.json"):
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from a committed ncu capture
            prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))
            for name, key in (("gather_rows", "extract"), ("fused_batch", "sample")):
                d = prof[name][0]
                traffic[key] = (float(d["dram__bytes_read.sum"].split()[0]) +
                                float(d["dram__bytes_write.sum"].split()[0])) * 1e6
            traffic_src = (f"profiles/{prof_name} (ncu --set full capture of this kernel on the products "
                           f"workload, N = 1; a constant read from the file, not measured in this run)")
            break
        except Exception:
            continue
    if world > 1 or args.shape != "products" or args.bias or args.cache_ratio is not None:
        traffic, traffic_src = {}, None     # the capture is for the default workload only
    roof_extract = {"bound": "hbm", "kernel": "gather_rows_kernel (feature extract)",
                    "achieved": ex_gbs, "peak": peak, "unit": "GB/s", "frac": ex_gbs / peak,
                    "peak_source": peak_src, "traffic": traffic.get("extract"),
                    "traffic_source": traffic_src if traffic.get("extract") else None,
                    "algorithmic_bytes": "rows * (2 * row_bytes + 8)", "avg_launch_ms": ex_ms / K,
                    "rows_per_launch": ex_rows / K}
    # bytes on the wire per payload byte of a peer row read, from ncu nvlrx__bytes.sum /
    # (rows * row_bytes) (profiles/r02_nvlink_ncu.csv): whole 32-byte sectors + ~16 B of protocol per
    # 128-byte line
    wire = {400: 1.2562, 512: 1.125}.get(row_bytes)
    if remote_frac > 0:
        peer = ex_rows * row_bytes * remote_frac / (ex_ms_max * 1e-3) / 1e9
        roof_extract.update({"bound": "nvlink", "peer_load_gbps_per_gpu": peer, "nvlink_peak_gbps": 900,
                             "nvlink_frac": peer / 900.0,
                             "nvlink_wire_gbps_per_gpu": peer * wire if wire else None,
                             "nvlink_wire_frac": peer * wire / 900.0 if wire else None,
                             "note": "remote fraction (P-1)/P of the gathered row bytes / slowest "
                                     "rank's kernel time, against 900 GB/s NVLink ingress per GPU; "
                                     "wire = payload x the ncu-measured nvlrx__bytes ratio for this "
                                     "row size (profiles/r02_nvlink_ncu.csv)"})
    if extract_only is not None:
        extract_only["nvlink_frac"] = extract_only["peer_payload_gbps_per_gpu"] / 900.0
        if wire:
            extract_only["nvlink_wire_gbps_per_gpu"] = extract_only["peer_payload_gbps_per_gpu"] * wire
            extract_only["nvlink_wire_frac"] = extract_only["peer_payload_gbps_per_gpu"] * wire / 900.0
    roof_sample = {"bound": "hbm", "kernel": "multi_batch_kernel (all hops: sample + relabel, one "
                                             "cooperative launch per batch)",
                   "achieved": sm_gbs, "peak": peak, "unit": "GB/s", "frac": sm_gbs / peak,
                   "peak_source": peak_src, "traffic": traffic.get("sample"),
                   "traffic_source": traffic_src if traffic.get("sample") else None,
                   "algorithmic_bytes": "sum over hops of 24 (S + nnz) [sample] + 8 (S + nnz) + 32 nnz + 8 U [relabel]",
                   "avg_launch_ms": sm_ms / K,
                   "note": "not bandwidth-bound at batch %d: limited by the issue rate of uncoalesced "
                           "accesses / random atomics and by dependent phases with grid barriers "
                           "(DESIGN.md section 5); e2e_pipelined puts B batches in one launch" % args.batch}
    dominant, other = (roof_sample, roof_extract) if sm_ms >= ex_ms else (roof_extract, roof_sample)
    dominant = dict(dominant)
    dominant["share_of_gpu_time"] = max(sm_ms, ex_ms) / (sm_ms + ex_ms)
    line = {
        "metric": METRIC, "value": t_dev["edges"] / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": config_of(args, fan_out),
        "setup": {"layout": layout, "extract_algo": args.extract_algo, "host_cores": host_cores()},
        "parity_checked": parity_checked,
        "batches_per_sec": world * K / (ms * 1e-3),
        "extract_gbps": t_dev["rows"] * (2 * row_bytes + 8) / (ms * 1e-3) / 1e9,
        "clocks": clk,
        "e2e": {"value": t_e2e["edges"] / (t_e2e["ms"] * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": args.batch * 8, "d2h_bytes_per_step": args.batch * 8 + 16 * len(fan_out),
                "ms_per_step": t_e2e["ms"] / K, "batches_per_sec": world * K / (t_e2e["ms"] * 1e-3),
                "note": "seeds from pinned host memory, blocks + features stay on the device (the "
                        "plugin API returns CUDA tensors), labels of the batch + hop sizes read back"},
        "placement_policy": policy,
        "replica_layout": extra_layout,
        "e2e_fused": {"value": t_fused["edges"] / (t_fused["ms"] * 1e-3), "unit": UNIT,
                      "ms_per_step": t_fused["ms"] / K,
                      "batches_per_sec": world * K / (t_fused["ms"] * 1e-3),
                      "note": "extension, not the reference-facing API: dgs.classes.BatchLoader.load = one "
                              "native call: seeds H2D -> labels + D2H -> sample -> extract (frontier size "
                              "read on the device); returns when the hop sizes and the labels are on the "
                              "host, WITHOUT draining the stream (the extract of step i overlaps the host "
                              "side of step i + 1; everything is synchronised at the end of the K steps); "
                              "same inputs, outputs and copies as e2e"},
        "e2e_pipelined": pipelined,
        "gpu_launches": launches,
        "roofline": dominant,
        "roofline_other": other,
        "extract_only": extract_only,
        "reference_gpu": ref_gpu,
    }
    if host_graph is not None:
        cb = cpu_baseline(args, fan_out, host_graph, args.cpu_baseline_seconds)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["batches_per_sec"] = cb["batches_per_sec"]
        line["cpu_baseline"]["extract_gbps"] = cb["extract_gbps"]
    elif big:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port",
                                "sample": f"skipped: the CPU arm would need the whole {args.shape} graph "
                                          f"in host memory (run `bench.py --impl reference --shape "
                                          f"{args.shape}` on a box with enough RAM)"}
    emit(line)
    if world > 1:
        quiet_barrier()
        dist.destroy_process_group()


def host_csr_to_pinned(host_csr, ft):
    ip, ix = host_csr
    return (torch.from_numpy(ip).pin_memory(), torch.from_numpy(ix).pin_memory(), ft.cpu().pin_memory())


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
