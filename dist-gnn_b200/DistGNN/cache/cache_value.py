"""Cache policy (python/DistGNN/cache/cache_value.py): node heat over the B200 heat
kernels, and the selfish / selfless placement model with the reference's function names, argument
order and results.  Everything below get_node_heat is device-agnostic tensor code (it follows the
device of the heat vectors instead of hard-coding "cuda"), so the 2-rank tests run on gloo.

Differences from the reference, none visible in the results:
* get_hot_nids_p2p_global needs no root: two N-length all-reduces (max heat, then lowest rank that
  holds the max) replace gather-to-root + argmax + scatter + send/recv (cache_value.py:65-150);
* B200_COST_MODEL holds re-measured constants for the cost-model arguments the reference's example
  hard-codes for its own machine (example/graphsage/node_classification.py:79-85)."""
import torch
import torch.distributed as dist

# Arguments of the placement model, re-measured on this pool's B200s (profiles/r01_results.md):
# GB/s for random row reads and bytes moved per visit of a node, in the units the reference uses
# (only ratios matter).  bandwidth_gpu: local HBM gather; bandwidth_nvlink: peer loads per GPU when
# every GPU reads every other; bandwidth_host: pinned host rows over PCIe Gen5.
B200_COST_MODEL = {
    "bandwidth_gpu": 4600.0, "bandwidth_nvlink": 624.0, "bandwidth_host": 50.0,
    "sampling_read_bytes_gpu": 480, "sampling_read_bytes_host": 480,
    "feature_read_bytes_gpu": 480, "feature_read_bytes_host": 512,
}


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
    pinned_here = []
    if mode == "uva":
        for t in (indptr, indices, probs):
            if t is not None and not t.is_pinned():    # already-pinned inputs are left alone
                ops._CAPI_tensor_pin_memory(t)
                pinned_here.append(t)
    else:
        indptr, indices = indptr.cuda(), indices.cuda()
        if probs is not None:
            probs = probs.cuda()
    num_nodes = indptr.shape[0] - 1
    sampling_heat = torch.zeros(num_nodes).cuda()
    seeds_heat = torch.zeros(num_nodes).cuda()
    seeds_heat[node_ids] = 1
    seeds = node_ids.cuda().to(indices.dtype)   # the heat kernel reads seeds with indices' id type
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
    torch.cuda.current_stream().synchronize()
    for t in pinned_here:
        ops._CAPI_tensor_unpin_memory(t)
    return sampling_heat, feature_heat


def get_cache_nids_by_degree(indptr, ratio, rank=0, world_size=1):
    """Hot set used by the cache-ratio sweep (BASELINE config 3): the top `ratio` fraction of the
    nodes by degree, dealt round-robin to the ranks."""
    deg = indptr[1:] - indptr[:-1]
    n = int(round(float(ratio) * deg.numel()))
    order = torch.argsort(deg, descending=True, stable=True)[:n]
    return order[rank::world_size].contiguous()


def get_hot_nids_local(sampling_heat, feature_heat):
    """Nodes this rank ever visits (cache_value.py:57-61)."""
    return torch.nonzero(sampling_heat).flatten(), torch.nonzero(feature_heat).flatten()


def _hottest_rank_owns(heat, group):
    """Ids of the nodes whose heat is highest on this rank (ties -> lowest rank, like argmax over
    the stacked heats at cache_value.py:92-93), ascending, restricted to heat > 0 (:147)."""
    size, rank = dist.get_world_size(group), dist.get_rank(group)
    top = heat.clone()
    dist.all_reduce(top, dist.ReduceOp.MAX, group)
    claim = torch.where(heat == top, rank, size).to(torch.int32)
    dist.all_reduce(claim, dist.ReduceOp.MIN, group)
    return torch.nonzero((claim == rank) & (heat > 0)).flatten()


def get_hot_nids_p2p_global(sampling_heat, feature_heat, group=None):
    """Deal every node to the rank of its p2p group on which it is hottest
    (cache_value.py:65-150).  Collective over `group`."""
    return _hottest_rank_owns(sampling_heat, group), _hottest_rank_owns(feature_heat, group)


def get_structure_space(nids, graph, probs=None):
    """Bytes each of `nids` costs in the structure cache: its row of indices (+ weights `graph[probs]`)
    and one indptr entry (cache_value.py:153-165)."""
    indptr = graph["indptr"].to(nids.device)
    per_edge = graph["indices"].element_size()
    if probs is not None:
        per_edge += graph[probs].element_size()
    return (indptr[nids + 1] - indptr[nids]) * per_edge + indptr.element_size()


def get_feature_space(graph):
    """Bytes of one feature row (cache_value.py:168-173)."""
    feat = graph["features"]
    return int(feat.element_size() * feat.numel() / (graph["indptr"].numel() - 1))


def get_node_value(heat, space_bytes, reduced_time):
    """Time saved per cached byte (cache_value.py:176-180)."""
    assert isinstance(space_bytes, (int, torch.Tensor))
    return heat / space_bytes * reduced_time


def get_cache_nids_local(sampling_nids, sampling_space, sampling_value, feature_nids, feature_space,
                         feature_value, free_capacity_bytes):
    """Greedy knapsack over structure rows and feature rows together: take candidates by
    descending value until `free_capacity_bytes` is used (cache_value.py:183-206).
    -> (structure nids, feature nids, bytes used)."""
    ns = sampling_nids.numel()
    order = torch.argsort(torch.cat([sampling_value, feature_value]), descending=True)
    used = torch.cumsum(torch.cat([sampling_space, feature_space])[order], 0)
    cut = torch.searchsorted(used, torch.tensor([free_capacity_bytes], device=used.device))
    take = order[:cut]
    is_structure = take < ns
    return sampling_nids[take[is_structure]], feature_nids[take[~is_structure] - ns], used[cut - 1].item()


def _reduced_times(bandwidth_gpu, sampling_read_bytes_gpu, feature_read_bytes_gpu, bandwidth_host,
                   sampling_read_bytes_host, feature_read_bytes_host):
    return (sampling_read_bytes_host / bandwidth_host - sampling_read_bytes_gpu / bandwidth_gpu,
            feature_read_bytes_host / bandwidth_host - feature_read_bytes_gpu / bandwidth_gpu)


def _place(graph, sampling_heat, feature_heat, sampling_hot, feature_hot, available_mem, times, probs):
    s_space = get_structure_space(sampling_hot, graph, probs=probs)
    s_value = get_node_value(sampling_heat[sampling_hot], s_space, times[0])
    row = get_feature_space(graph)
    f_value = get_node_value(feature_heat[feature_hot], row, times[1])
    return get_cache_nids_local(sampling_hot, s_space, s_value, feature_hot,
                                torch.full_like(feature_hot, row), f_value, available_mem)


def get_cache_nids_selfish(graph, sampling_heat, feature_heat, available_mem, bandwidth_gpu,
                           sampling_read_bytes_gpu, feature_read_bytes_gpu, bandwidth_host,
                           sampling_read_bytes_host, feature_read_bytes_host, probs=None):
    """Every rank fills its own memory with what IT visits most (replication across ranks allowed)
    (cache_value.py:210-240)."""
    times = _reduced_times(bandwidth_gpu, sampling_read_bytes_gpu, feature_read_bytes_gpu,
                           bandwidth_host, sampling_read_bytes_host, feature_read_bytes_host)
    s_hot, f_hot = get_hot_nids_local(sampling_heat, feature_heat)
    s, f, _ = _place(graph, sampling_heat, feature_heat, s_hot, f_hot, available_mem, times, probs)
    return s, f


def get_cache_nids_selfless(graph, sampling_heat, feature_heat, available_mem, bandwidth_gpu,
                            sampling_read_bytes_gpu, feature_read_bytes_gpu, bandwidth_host,
                            sampling_read_bytes_host, feature_read_bytes_host, probs=None, group=None):
    """Partition first - a node goes to the rank where it is hottest - then spend what memory is
    left selfishly on the rest, and order both lists by descending heat
    (cache_value.py:244-310).  Collective over `group`."""
    times = _reduced_times(bandwidth_gpu, sampling_read_bytes_gpu, feature_read_bytes_gpu,
                           bandwidth_host, sampling_read_bytes_host, feature_read_bytes_host)
    s_hot, f_hot = get_hot_nids_p2p_global(sampling_heat, feature_heat, group=group)
    s, f, used = _place(graph, sampling_heat, feature_heat, s_hot, f_hot, available_mem, times, probs)
    if available_mem - used > 0:
        # second round over the nodes not yet placed here (their heat masked out, not modified)
        s_rest, f_rest = sampling_heat.clone(), feature_heat.clone()
        s_rest[s] = 0
        f_rest[f] = 0
        s2, f2 = get_cache_nids_selfish(graph, s_rest, f_rest, available_mem - used, bandwidth_gpu,
                                        sampling_read_bytes_gpu, feature_read_bytes_gpu, bandwidth_host,
                                        sampling_read_bytes_host, feature_read_bytes_host, probs=probs)
        s, f = torch.cat([s, s2]), torch.cat([f, f2])
        s = s[torch.argsort(sampling_heat[s], descending=True)]
        f = f[torch.argsort(feature_heat[f], descending=True)]
    return s, f


def compute_total_value_selfish(graph, sampling_heat, feature_heat, sampling_cache_nids,
                                feature_cache_nids, bandwidth_gpu, sampling_read_bytes_gpu,
                                feature_read_bytes_gpu, bandwidth_host, sampling_read_bytes_host,
                                feature_read_bytes_host, probs=None):
    """Sum of the values of a placement (cache_value.py:314-344)."""
    times = _reduced_times(bandwidth_gpu, sampling_read_bytes_gpu, feature_read_bytes_gpu,
                           bandwidth_host, sampling_read_bytes_host, feature_read_bytes_host)
    s_space = get_structure_space(sampling_cache_nids, graph, probs=probs)
    total = torch.sum(get_node_value(sampling_heat[sampling_cache_nids], s_space, times[0])).item()
    total += torch.sum(get_node_value(feature_heat[feature_cache_nids], get_feature_space(graph),
                                      times[1])).item()
    return total


def compute_total_value_selfless(graph, sampling_heat, feature_heat, sampling_cache_nids,
                                 feature_cache_nids, bandwidth_gpu, bandwidth_nvlink, num_gpu,
                                 sampling_read_bytes_gpu, feature_read_bytes_gpu, bandwidth_host,
                                 sampling_read_bytes_host, feature_read_bytes_host, probs=None,
                                 group=None):
    """Value of a partitioned placement: own nodes at the local bandwidth left once the peers'
    reads are served, plus the nodes cached on any OTHER rank at NVLink bandwidth
    (cache_value.py:347-409).  Collective over `group`."""
    rest = (sampling_read_bytes_gpu, feature_read_bytes_gpu, bandwidth_host, sampling_read_bytes_host,
            feature_read_bytes_host)
    local = compute_total_value_selfish(graph, sampling_heat, feature_heat, sampling_cache_nids,
                                        feature_cache_nids, bandwidth_gpu - (num_gpu - 1) * bandwidth_nvlink,
                                        *rest, probs=probs)
    n = graph["indptr"].numel() - 1
    remote = []
    for mine in (sampling_cache_nids, feature_cache_nids):
        held = torch.zeros(n, dtype=torch.int32, device=sampling_heat.device)
        held[mine] = 1
        dist.all_reduce(held, dist.ReduceOp.SUM, group)     # cached by how many ranks
        held[mine] = 0
        remote.append(torch.nonzero(held).flatten())
    return local + compute_total_value_selfish(graph, sampling_heat, feature_heat, remote[0], remote[1],
                                               bandwidth_nvlink, *rest, probs=probs)


def choose_cache_policy(graph, sampling_heat, feature_heat, available_mem, num_gpu, probs=None,
                        group=None, cost_model=None):
    """The "auto" policy of example/graphsage/node_classification.py:86-167 as one call: compute
    both placements, sum their values over all ranks, keep the better one.
    -> (policy name, sampling_cache_nids, feature_cache_nids).  Collective."""
    cm = dict(B200_COST_MODEL if cost_model is None else cost_model)
    nv = cm.pop("bandwidth_nvlink")
    args = (cm["bandwidth_gpu"], cm["sampling_read_bytes_gpu"], cm["feature_read_bytes_gpu"],
            cm["bandwidth_host"], cm["sampling_read_bytes_host"], cm["feature_read_bytes_host"])
    selfish = get_cache_nids_selfish(graph, sampling_heat, feature_heat, available_mem, *args, probs=probs)
    selfless = get_cache_nids_selfless(graph, sampling_heat, feature_heat, available_mem, *args,
                                       probs=probs, group=group)
    v = torch.tensor([
        compute_total_value_selfish(graph, sampling_heat, feature_heat, *selfish, *args, probs=probs),
        compute_total_value_selfless(graph, sampling_heat, feature_heat, *selfless, cm["bandwidth_gpu"],
                                     nv, num_gpu, *args[1:], probs=probs, group=group)],
        dtype=torch.float64, device=sampling_heat.device)
    dist.all_reduce(v, dist.ReduceOp.SUM, group)
    if v[0].item() > v[1].item():
        return ("selfish",) + tuple(selfish)
    return ("selfless",) + tuple(selfless)


def get_available_memory(device, reserved_mem):
    """Device memory not yet allocated by torch minus a reserve (cache_value.py:412-417)."""
    total = torch.cuda.mem_get_info(device)[1]
    return max(int(total - torch.cuda.memory_allocated(device=device) - reserved_mem), 0)
