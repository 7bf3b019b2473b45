"""Cache-policy helpers (subset of python/DistGNN/cache/cache_value.py; the selfish / selfless /
auto cost model is SURVEY.md 8f-3, a later row).  What the data-path benchmarks need is here:
heat propagation over the sampling fan-out and simple hot-set selection."""
import torch


def get_node_heat(indptr, indices, seeds, fan_out, probs=None, capi=None):
    """Expected visit count of every node when `seeds` are sampled with `fan_out`
    (cache_value.py:6-53): heat flows seed -> neighbour with min(1, heat * k / deg) per edge, one
    round per hop (fan_out walked from the back).  All tensors CUDA (or pinned host for the CSR)."""
    if capi is None:
        import dgs as capi
    num_nodes = indptr.numel() - 1
    dev = seeds.device
    heat = torch.zeros(num_nodes, dtype=torch.float32, device=dev)
    heat[seeds] = 1.0
    total = heat.clone()
    cur = seeds
    for k in reversed(list(fan_out)):
        if probs is None:
            nxt = capi.ops._CAPI_compute_frontier_heat(cur, indptr, indices, heat, int(k), 0)
        else:
            nxt = capi.ops._CAPI_compute_frontier_heat_with_bias(cur, indptr, indices, probs, heat,
                                                                 int(k), 0)
        total += nxt
        heat = nxt
        cur = torch.nonzero(nxt > 0).reshape(-1).to(seeds.dtype)
    return total


def get_cache_nids_by_degree(indptr, ratio, rank=0, world_size=1):
    """Hot set used by the cache-ratio sweep (BASELINE config 3): the top `ratio` fraction of the
    nodes by degree, dealt round-robin to the ranks."""
    deg = indptr[1:] - indptr[:-1]
    n = int(round(float(ratio) * deg.numel()))
    order = torch.argsort(deg, descending=True, stable=True)[:n]
    return order[rank::world_size].contiguous()


def get_structure_space(indptr, indices, probs=None):
    """Bytes per node of cached structure (cache_value.py:412-417 style accounting)."""
    per_edge = indices.element_size() + (probs.element_size() if probs is not None else 0)
    deg = (indptr[1:] - indptr[:-1]).to(torch.int64)
    return deg * per_edge + indptr.element_size()


def get_feature_space(features):
    stride = 1
    for d in features.shape[1:]:
        stride *= d
    return stride * features.element_size()


def get_available_memory(device, reserve_bytes=7 << 30):
    """Free device memory minus a reserve (7 GiB in example/graphsage/node_classification.py:73)."""
    free, _ = torch.cuda.mem_get_info(device)
    return max(0, free - reserve_bytes)
