"""NCCL bootstrap for the data-path plugin (python/DistGNN/dist/communicator.py:5-17)."""
import torch.distributed as dist


def create_communicator(group_size, group=None, capi=None):
    """Rank 0 of every `group_size`-rank group creates the NCCL unique id, it is broadcast to the
    group with torch.distributed (works over gloo as well as nccl), and each rank initialises the
    process-global dgs NCCL context with (group_size, id, rank-in-group).

    `capi` (default: the `dgs` module) only needs `ops._CAPI_get_unique_id` / `ops._CAPI_set_nccl`;
    tests pass a stub to check the broadcast logic without a GPU.  As in the reference, `group`
    must be the process group made of exactly those group_size ranks (or None when
    group_size == world size); the reference's own tests pass the rank here by mistake
    (tests/test_nccl.py:12) - that is rejected."""
    if capi is None:
        import dgs as capi
    if group is not None and not isinstance(group, dist.ProcessGroup):
        raise TypeError("group must be a torch.distributed ProcessGroup or None")
    rank = dist.get_rank()
    group_rank = rank % group_size
    group_root = rank - group_rank
    if group_rank == 0:
        broadcast_list = [capi.ops._CAPI_get_unique_id()]
    else:
        broadcast_list = [None]
    dist.broadcast_object_list(broadcast_list, group_root, group)
    unique_ids = broadcast_list[0]
    capi.ops._CAPI_set_nccl(group_size, unique_ids, group_rank)
    return unique_ids


def owner_of(nids, world_size):
    """Shard owner used by the synthetic benchmarks: node n lives on GPU n mod P."""
    return nids % world_size


def partition_seeds(seeds, rank, world_size):
    """Data-parallel split of the training seeds: rank r takes the r-th contiguous slice
    (example/graphsage/node_classification.py:309-321 splits train_nids the same way)."""
    n = seeds.shape[0]
    per = (n + world_size - 1) // world_size
    return seeds[rank * per:min(n, (rank + 1) * per)]
