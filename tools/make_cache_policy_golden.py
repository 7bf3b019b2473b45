"""Generates tests/golden/ref_cache_policy.npz by running the REFERENCE's cache policy
(/root/reference/python/DistGNN/cache/cache_value.py, imported, not copied) on seeded inputs in a
2-rank gloo group on the CPU.  The reference hard-codes device="cuda" in a few tensor constructors;
a proxy `torch` handed to that module drops the argument, nothing else is touched.  Run in the
build container (needs /root/reference):   python tools/make_cache_policy_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
REF = "/root/reference/python/DistGNN/cache/cache_value.py"
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cache_policy_inputs import COST, MEMS, WORLD, cost_args, inputs  # noqa: E402


class _TorchNoCuda:
    """getattr-forwarding view of torch whose tensor factories ignore device="cuda"."""

    def __getattr__(self, name):
        obj = getattr(torch, name)
        if name in ("tensor", "zeros", "zeros_like", "full_like"):
            def factory(*a, **kw):
                if kw.get("device") == "cuda":
                    kw.pop("device")
                return obj(*a, **kw)
            return factory
        return obj


def load_reference():
    sys.modules.setdefault("dgs", types.ModuleType("dgs"))     # only get_node_heat uses it (not called)
    spec = importlib.util.spec_from_file_location("ref_cache_value", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.torch = _TorchNoCuda()
    return mod


def worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    ref = load_reference()
    res = {}
    for bias in (False, True):
        for mem in MEMS:
            graph, sh, fh, probs = inputs(rank, bias)
            a = cost_args()
            tag = f"r{rank}_b{int(bias)}_m{mem}"
            s1, f1 = ref.get_cache_nids_selfish(graph, sh, fh, mem, *a, probs=probs)
            v1 = ref.compute_total_value_selfish(graph, sh, fh, s1, f1, *a, probs=probs)
            s2, f2 = ref.get_cache_nids_selfless(graph, sh, fh, mem, *a, probs=probs)
            v2 = ref.compute_total_value_selfless(graph, sh, fh, s2, f2, COST["bandwidth_gpu"],
                                                  COST["bandwidth_nvlink"], WORLD, *a[1:], probs=probs)
            hs, hf = ref.get_hot_nids_p2p_global(sh, fh)
            for k, v in (("selfish_s", s1), ("selfish_f", f1), ("selfless_s", s2), ("selfless_f", f2),
                         ("global_s", hs), ("global_f", hf)):
                res[f"{tag}_{k}"] = v.numpy()
            res[f"{tag}_values"] = np.array([v1, v2], dtype=np.float64)
    np.savez_compressed(out + f".rank{rank}.npz", **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "ref_cache_policy")
    mp.spawn(worker, args=(29577, out), nprocs=WORLD, join=True)
    merged = {}
    for r in range(WORLD):
        with np.load(out + f".rank{r}.npz") as z:
            merged.update({k: z[k] for k in z.files})
        os.remove(out + f".rank{r}.npz")
    np.savez_compressed(out + ".npz", **merged)
    print("wrote", out + ".npz", len(merged), "arrays")
