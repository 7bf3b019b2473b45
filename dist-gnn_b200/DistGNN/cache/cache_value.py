"""Cache-policy helpers (subset of python/DistGNN/cache/cache_value.py; the selfish / selfless /
auto cost model is SURVEY.md 8f-3, a later row).  What the data-path benchmarks need is here:
get_node_heat with the reference's semantics (over the B200 heat kernels) and a simple
degree-ranked hot set for the cache-ratio sweep."""
import torch


def get_node_heat(indptr, indices, node_ids, fan_outs, probs=None, mode="uva", capi=None):
    """Expected visit counts of every node when `node_ids` are sampled with `fan_outs`
    (python/DistGNN/cache/cache_value.py:6-53).  Returns (sampling_heat, feature_heat), float32
    CUDA tensors over all nodes: heat flows seed -> neighbour with min(1, heat * k / deg) per edge
    (weight-proportional when `probs` is given), one round per hop with fan_outs walked from the
    back.  mode "uva": the CPU CSR is pinned for the duration and read in place; "cuda": moved."""
    assert mode in ["uva", "cuda"]
    if capi is None:
        import dgs as capi
    ops = capi.ops
    if mode == "uva":
        for t in (indptr, indices, probs):
            if t is not None:
                ops._CAPI_tensor_pin_memory(t)
    else:
        indptr, indices = indptr.cuda(), indices.cuda()
        if probs is not None:
            probs = probs.cuda()
    num_nodes = indptr.shape[0] - 1
    sampling_heat = torch.zeros(num_nodes).cuda()
    seeds_heat = torch.zeros(num_nodes).cuda()
    seeds_heat[node_ids] = 1
    seeds = node_ids.cuda()
    frontier_heat = torch.zeros(num_nodes).cuda()
    for num_picks in reversed(list(fan_outs)):
        if probs is None:
            frontier_heat = ops._CAPI_compute_frontier_heat(seeds, indptr, indices, seeds_heat,
                                                            num_picks, 0)
        else:
            frontier_heat = ops._CAPI_compute_frontier_heat_with_bias(seeds, indptr, indices, probs,
                                                                      seeds_heat, num_picks, 0)
        sampling_heat += seeds_heat
        seeds_heat += frontier_heat
        seeds = torch.nonzero(seeds_heat > 0).squeeze(1).to(indices.dtype)
    feature_heat = sampling_heat + frontier_heat
    if mode == "uva":
        for t in (indptr, indices, probs):
            if t is not None:
                ops._CAPI_tensor_unpin_memory(t)
    return sampling_heat, feature_heat


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
