"""Seeded inputs of the cache-policy golden vectors, shared by the generator
(tools/make_cache_policy_golden.py, runs the reference) and tests/test_cache_policy.py."""
import torch

import dgs_synth

WORLD = 2
MEMS = (40_000, 200_000, 10_000_000)
# the constants of example/graphsage/node_classification.py:79-85
COST = dict(bandwidth_gpu=120.62, bandwidth_host=8.32, bandwidth_nvlink=9.25, sampling_read_bytes_gpu=480,
            sampling_read_bytes_host=480, feature_read_bytes_gpu=480, feature_read_bytes_host=512)


def cost_args():
    return (COST["bandwidth_gpu"], COST["sampling_read_bytes_gpu"], COST["feature_read_bytes_gpu"],
            COST["bandwidth_host"], COST["sampling_read_bytes_host"], COST["feature_read_bytes_host"])


def inputs(rank, bias):
    n, d = 3000, 24
    indptr, indices, probs = dgs_synth.make_csr(n, 50000, seed=77, weights=True, classes=6)
    graph = {"indptr": indptr, "indices": indices, "probs": probs,
             "features": torch.zeros(n, d, dtype=torch.float32)}
    g = torch.Generator().manual_seed(1000 + rank)
    s_heat = torch.rand(n, generator=g) * (torch.rand(n, generator=g) < 0.4)
    f_heat = s_heat + torch.rand(n, generator=g) * (torch.rand(n, generator=g) < 0.5)
    return graph, s_heat.float(), f_heat.float(), ("probs" if bias else None)
