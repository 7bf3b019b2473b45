"""Id-exchange variant of the sharded feature extract (north star (4): "NCCL used only for seed/ID
exchange"; SURVEY 8e: "an optional alternative to be measured against pure peer loads").

The default extract reads remote rows with NVLink peer loads issued inside the gather kernel
(dgs_extract_sharded, replacing _IndexP2PCacheKernel, src/feature/cuda/feature_ops.cu:38-73).
Here the requests travel instead: every rank routes its ids to their owners with one NCCL
all-to-all, the owners gather the rows from their OWN shard (HBM speed, no remote latency) and a
second all-to-all carries the rows back.  One host round trip per extract (the split sizes).
The reference has no such path; its only id exchange is the build-time all-gather-v of cache id
lists (src/nccl/nccl_context.cc:65-112).  bench.py times both on the same requests
(`extract_only.exchange_*`); DESIGN.md section 6 has the numbers.
"""
import torch
import torch.distributed as dist


def exchange_extract(nids, world, rank, local_rows, group=None, route=None, gather=None):
    """out[i] = feature row of node nids[i] for the modulo layout (node n = row n // world of the
    shard on rank n % world; `local_rows` is this rank's shard).  Collective over `group`.

    route(nids, world) -> (send_idx, inv, counts) and gather(table, idx) -> rows default to the
    native kernels (dgs.ops.route_ids, dgs.ops._CAPI_cuda_index_select) - they fail loudly without
    the CUDA library; the gloo tests inject torch restatements to check the collective logic."""
    if route is None or gather is None:
        import dgs
        route = route or dgs.ops.route_ids
        gather = gather or dgs.ops._CAPI_cuda_index_select
    if dist.get_world_size(group) != world or dist.get_rank(group) != rank:
        raise RuntimeError("exchange_extract: (world, rank) must be those of the process group")
    nids = nids.contiguous()
    n = nids.numel()
    send_idx, inv, counts = route(nids, world)
    # requests of every rank for every owner: m[r][d]; ONE host round trip per extract
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    m = torch.stack(all_counts).cpu()
    send_splits = m[rank].tolist()
    recv_splits = m[:, rank].tolist()
    if sum(send_splits) != n:
        raise RuntimeError("exchange_extract: routed request counts do not add up")
    recv_idx = torch.empty(sum(recv_splits), dtype=nids.dtype, device=nids.device)
    dist.all_to_all_single(recv_idx, send_idx, recv_splits, send_splits, group=group)
    rows_send = gather(local_rows, recv_idx)                  # local shard only
    rows_recv = torch.empty((n,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype,
                            device=local_rows.device)
    dist.all_to_all_single(rows_recv, rows_send, send_splits, recv_splits, group=group)
    return gather(rows_recv, inv)                             # back to request order
