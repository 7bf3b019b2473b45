"""``dgs.ops`` - the free functions of the reference's pybind module (src/pybind.cc:47-77),
implemented over the sm_100a C-ABI library.  Names, argument order and error behaviour follow the
reference; every result is a freshly allocated CUDA tensor on the current device."""
import ctypes as C

import torch

from . import _lib
from ._util import (MAX_TWO_PASS_ELEMS, check_cpu, check_cuda, check_device_readable, contiguous,
                    itype, ptr, stream, zero_ws)

lib = _lib.lib
check = _lib.check


# ------------------------------------------------------------------ NCCL context
def _CAPI_get_unique_id():
    """nccl::GetUniqueId (src/nccl/nccl_context.cc:13-18): 16 int64 = the 128-byte ncclUniqueId."""
    buf = (C.c_int64 * 16)()
    check(lib().dgs_nccl_get_unique_id(buf), "_CAPI_get_unique_id")
    return [int(x) for x in buf]


def _CAPI_set_nccl(nranks, unique_id_array, rank):
    """nccl::SetNCCL (src/nccl/nccl_context.cc:20-23, 34-44)."""
    if len(unique_id_array) != 16:
        raise RuntimeError("unique id must be a list of 16 int64")
    buf = _lib.i64_array(unique_id_array)
    check(lib().dgs_nccl_set(int(nranks), buf, int(rank)), "_CAPI_set_nccl")


def _Test_GetLocalRank():
    return int(lib().dgs_nccl_rank())


def _Test_GetWorldSize():
    return int(lib().dgs_nccl_world())


def _Test_Randn():
    """ctx::randn_uint64 (src/context/context.cc:6)."""
    return int(lib().dgs_randn_uint64())


def seed(value):
    """Extension: re-seed the process-global random engine (the reference cannot be seeded)."""
    lib().dgs_seed(int(value) & 0xFFFFFFFFFFFFFFFF)


def _allgather_tensors(local):
    """Variable-length all-gather of a 1-D CUDA tensor -> list of per-rank tensors."""
    l = lib()
    world, rank = l.dgs_nccl_world(), l.dgs_nccl_rank()
    local = local.contiguous()
    if world == 1:
        return [local]
    sizes = (C.c_int64 * world)()
    check(l.dgs_nccl_allgather_i64(local.numel(), sizes), "allgather sizes")
    outs = []
    for r in range(world):
        outs.append(local if r == rank else torch.empty(int(sizes[r]), dtype=local.dtype,
                                                         device=local.device))
    es = local.element_size()
    recv = _lib.vp_array([ptr(t) for t in outs])
    nbytes = _lib.i64_array([t.numel() * es for t in outs])
    check(l.dgs_nccl_allgatherv(ptr(local), local.numel() * es, recv, nbytes), "allgatherv")
    return outs


def _Test_NCCLTensorAllGather(local_tensor):
    """NCCLContext::NCCLTensorAllGather_ (src/nccl/nccl_context.cc:52-112)."""
    check_cuda(local_tensor, "local_tensor")
    return _allgather_tensors(local_tensor.reshape(-1))


# ------------------------------------------------------------------ pin memory
def _CAPI_tensor_pin_memory(data):
    """TensorPinMemory (src/common/pin_memory.cc:7-12): in-place cudaHostRegister."""
    check_cpu(data, "data")
    if data.is_pinned() or data.numel() == 0:
        return
    check(lib().dgs_host_register(data.data_ptr(), data.numel() * data.element_size()),
          "_CAPI_tensor_pin_memory")


def _CAPI_tensor_unpin_memory(data):
    """TensorUnpinMemory (src/common/pin_memory.cc:14-19)."""
    check_cpu(data, "data")
    if data.numel() == 0:
        return
    check(lib().dgs_host_unregister(data.data_ptr()), "_CAPI_tensor_unpin_memory")


# ------------------------------------------------------------------ feature gather
def _CAPI_cuda_index_select(data, nids, algo=0):
    """GetFeaturesCUDA (src/feature/cuda/feature_ops.cu:173-210): out[i] = data[nids[i]] for a
    CUDA or pinned-host table; the result has data's rank with dim 0 = len(nids)."""
    check_cuda(nids, "nids")
    check_device_readable(data, "data")
    contiguous(data, "data")
    nids = nids.contiguous()
    n = nids.numel()
    stride = 1
    for d in data.shape[1:]:
        stride *= d
    out = torch.empty((n,) + tuple(data.shape[1:]), dtype=data.dtype, device=nids.device)
    if n and stride:
        check(lib().dgs_index_select(ptr(data), stride * data.element_size(), itype(nids, "nids"),
                                     ptr(nids), n, ptr(out), int(algo), stream()),
              "_CAPI_cuda_index_select")
    return out


def route_ids(nids, world):
    """Extension (north star (4): NCCL for the seed / ID exchange; no reference counterpart):
    partition the requested node ids by owner for ONE all-to-all.  Node n lives on GPU n mod world
    at slot n // world (the layout of the modulo-sharded extract).  Returns (send_idx, inv, counts):
    send_idx = slot numbers grouped by owner, inv[i] = position of request i in that grouped order
    (rows that come back in send order are put in request order by out = rows[inv]), counts =
    int64[world] on the device.  Enqueued on the current stream, no host round trip."""
    check_cuda(nids, "nids")
    nids = nids.contiguous()
    n = nids.numel()
    l = lib()
    it = itype(nids, "nids")
    wsb = l.dgs_route_ws_bytes(n, int(world))
    if wsb < 0:
        raise RuntimeError(f"route_ids: world {world} not supported")
    with torch.cuda.device(nids.device):
        ws = torch.empty(int(wsb), dtype=torch.uint8, device=nids.device)
        send_idx = torch.empty_like(nids)
        inv = torch.empty_like(nids)
        counts = torch.empty(int(world), dtype=torch.int64, device=nids.device)
        check(l.dgs_route_ids(it, ptr(nids), n, int(world), ptr(send_idx), ptr(inv), ptr(counts),
                              ptr(ws), wsb, stream()), "route_ids")
    return send_idx, inv, counts


# ------------------------------------------------------------------ sub-CSR extraction (test hooks)
def _Test_ExtractIndptr(nids, indptr):
    """ExtractIndptr (src/sampling/cuda/utils.cu:12-42)."""
    check_cuda(nids, "nids")
    check_device_readable(indptr, "indptr")
    nids = nids.contiguous()
    n = nids.numel()
    sub = torch.empty(n + 1, dtype=indptr.dtype, device=nids.device)
    l = lib()
    ws = zero_ws(l.dgs_extract_indptr_ws_bytes(n), nids.device)
    check(l.dgs_extract_indptr(itype(nids, "nids"), itype(indptr, "indptr"), ptr(nids), n,
                               ptr(indptr), ptr(sub), ptr(ws), stream()), "_Test_ExtractIndptr")
    return sub


def _Test_ExtractEdgeData(nids, indptr, sub_indptr, edge_data):
    """ExtractEdgeData (src/sampling/cuda/utils.cu:71-101)."""
    check_cuda(nids, "nids")
    check_cuda(sub_indptr, "sub_indptr")
    check_device_readable(indptr, "indptr")
    check_device_readable(edge_data, "edge_data")
    if edge_data.dtype not in (torch.int32, torch.int64, torch.float32):
        raise RuntimeError("Value can only be int32 or int64 or float32")
    nids = nids.contiguous()
    n = nids.numel()
    total = int(sub_indptr[n].item()) if n else 0
    out = torch.empty(total, dtype=edge_data.dtype, device=nids.device)
    if total:
        check(lib().dgs_extract_edge_data(itype(nids, "nids"), itype(indptr, "indptr"),
                                          edge_data.element_size(), ptr(nids), n, ptr(indptr),
                                          ptr(sub_indptr), ptr(edge_data), ptr(out), stream()),
              "_Test_ExtractEdgeData")
    return out


# ------------------------------------------------------------------ sampling
def _make_graph(indptr, indices, probs=None):
    g = _lib.Graph()
    g.itype = itype(indices, "indices")
    g.etype = itype(indptr, "indptr")
    g.indptr = ptr(indptr)
    g.indices = ptr(indices)
    g.probs = ptr(probs) if probs is not None else None
    return g


def _sample_one_hop(g, seeds, num_picks, replace, rng_seed=None, keep=()):
    """One hop through dgs_sample_neighbors; returns exactly-sized (coo_row, coo_col)."""
    l = lib()
    dev = seeds.device
    S = seeds.numel()
    k = int(num_picks)
    if rng_seed is None:
        rng_seed = l.dgs_randn_uint64()
    nnz_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = zero_ws(l.dgs_sample_ws_bytes(S), dev)
    st = stream()

    def run(row, col, cap):
        check(l.dgs_sample_neighbors(C.byref(g), ptr(seeds), S, None, k, int(bool(replace)),
                                     C.c_uint64(rng_seed), ptr(row), ptr(col), cap, ptr(nnz_dev),
                                     ptr(ws), st), "sample_neighbors")

    ub = S * k if k >= 0 else -1
    if 0 <= ub <= MAX_TWO_PASS_ELEMS:
        row = torch.empty(ub, dtype=seeds.dtype, device=dev)
        col = torch.empty(ub, dtype=seeds.dtype, device=dev)
        run(row, col, ub)
        nnz = int(nnz_dev.item())
        return row[:nnz], col[:nnz]
    # unknown / huge bound (num_picks = -1, or "num_picks >= max degree" as the reference needs for
    # full neighbourhoods): count first, then sample into exactly-sized outputs.  Same RNG key, so
    # both passes agree.
    run(None, None, 0)
    nnz = int(nnz_dev.item())
    row = torch.empty(nnz, dtype=seeds.dtype, device=dev)
    col = torch.empty(nnz, dtype=seeds.dtype, device=dev)
    if nnz:
        run(row, col, nnz)
    return row, col


def _check_sampling_args(seeds, indptr, indices, probs=None):
    check_cuda(seeds, "seeds")
    check_device_readable(indptr, "indptr")
    check_device_readable(indices, "indices")
    contiguous(indptr, "indptr")
    contiguous(indices, "indices")
    if seeds.dtype != indices.dtype:
        raise RuntimeError(f"seeds ({seeds.dtype}) and indices ({indices.dtype}) must share an id type")
    if probs is not None:
        check_device_readable(probs, "probs")
        contiguous(probs, "probs")
        if probs.dtype != torch.float32:
            raise RuntimeError("probs must be float32")
        if probs.numel() != indices.numel():
            raise RuntimeError("probs must have one weight per edge")


def _CAPI_cuda_sample_neighbors(seeds, indptr, indices, num_picks, replace, rng_seed=None):
    """RowWiseSamplingUniformCUDA (src/sampling/cuda/rowwise_sampling.cu:143-189).
    num_picks = -1 returns every neighbour (extension; DGL convention)."""
    _check_sampling_args(seeds, indptr, indices)
    seeds = seeds.contiguous()
    g = _make_graph(indptr, indices)
    return _sample_one_hop(g, seeds, num_picks, replace, rng_seed)


def _CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, num_picks, replace,
                                     rng_seed=None):
    """RowWiseSamplingBiasCUDA (src/sampling/cuda/rowwise_sampling_bias.cu:226-288); no
    num_picks <= 32 limit."""
    _check_sampling_args(seeds, indptr, indices, probs)
    seeds = seeds.contiguous()
    g = _make_graph(indptr, indices, probs)
    return _sample_one_hop(g, seeds, num_picks, replace, rng_seed)


def _Test_AresKeys(weights, rng_key, item):
    """Test hook: A-Res keys of one row for (rng_key, item); the k largest (ties: smaller position
    first) are the weighted sample without replacement every biased path must return."""
    check_cuda(weights, "weights")
    weights = weights.contiguous()
    if weights.dtype != torch.float32:
        raise RuntimeError("weights must be float32")
    out = torch.empty_like(weights)
    check(lib().dgs_debug_ares_keys(ptr(weights), weights.numel(),
                                    C.c_uint64(int(rng_key) & 0xFFFFFFFFFFFFFFFF),
                                    C.c_uint64(int(item)), ptr(out), stream()), "_Test_AresKeys")
    return out


# ------------------------------------------------------------------ relabel
def _relabel(mapping, to_relabel, table=None, capacity=None):
    l = lib()
    if len(mapping) == 0:
        raise RuntimeError("mapping tensors must not be empty")
    dt = mapping[0].dtype
    for t in list(mapping) + list(to_relabel):
        check_cuda(t, "relabel tensor")
        if t.dtype != dt:
            raise RuntimeError("all relabel tensors must share one id type")
    dev = mapping[0].device
    mapping = [t.contiguous().reshape(-1) for t in mapping]
    shapes = [t.shape for t in to_relabel]
    rel = [t.contiguous().reshape(-1) for t in to_relabel]
    if len(mapping) > 4:
        mapping = [torch.cat(mapping)]
    split = None
    if len(rel) > 4:
        split = [t.numel() for t in rel]
        rel = [torch.cat(rel)]
    n = sum(t.numel() for t in mapping)
    if table is None:
        capacity = l.dgs_relabel_table_capacity(n)
        table = torch.full((capacity * 2,), -1, dtype=torch.int64, device=dev)
    ws = zero_ws(l.dgs_relabel_ws_bytes(n), dev)
    unique = torch.empty(n, dtype=dt, device=dev)
    outs = [torch.empty_like(t) for t in rel]
    nuniq = torch.zeros(1, dtype=torch.int64, device=dev)
    check(l.dgs_relabel(itype(mapping[0]), len(mapping), _lib.vp_array([ptr(t) for t in mapping]),
                        _lib.i64_array([t.numel() for t in mapping]), None, len(rel),
                        _lib.vp_array([ptr(t) for t in rel]),
                        _lib.i64_array([t.numel() for t in rel]), None,
                        _lib.vp_array([ptr(t) for t in outs]), ptr(unique), ptr(nuniq), ptr(table),
                        capacity, ptr(ws), stream()), "relabel")
    u = int(nuniq.item())
    if split is not None:
        outs = list(outs[0].split(split))
    outs = [o.reshape(s) for o, s in zip(outs, shapes)]
    return unique[:u], outs


def _CAPI_cuda_sampled_tensor_relabel(mapping_tensors, requiring_relabel_tensors):
    """TensorRelabelCUDA (src/sampling/cuda/tensor_relabel.cu:182-205)."""
    return _relabel(list(mapping_tensors), list(requiring_relabel_tensors))


# ------------------------------------------------------------------ cache-policy heat
def _heat(seeds, indptr, indices, probs, seeds_heat, num_picks, indptr_diff):
    check_cuda(seeds, "seeds")
    check_cuda(seeds_heat, "seeds_heat")
    check_device_readable(indptr, "indptr")
    check_device_readable(indices, "indices")
    if seeds_heat.dtype != torch.float32:
        raise RuntimeError("seeds_heat must be float32")
    if seeds.dtype != indices.dtype:
        # the kernel reads both through one id type (the reference throws on data_ptr<IdType>)
        raise RuntimeError(f"seeds ({seeds.dtype}) and indices ({indices.dtype}) must share an id type")
    seeds = seeds.contiguous()
    out = torch.zeros_like(seeds_heat)
    check(lib().dgs_frontier_heat(itype(indices, "indices"), itype(indptr, "indptr"), ptr(seeds),
                                  seeds.numel(), ptr(indptr), ptr(indices),
                                  ptr(probs) if probs is not None else None, ptr(seeds_heat),
                                  ptr(out), int(num_picks), int(indptr_diff), stream()),
          "compute_frontier_heat")
    return out


def _CAPI_compute_frontier_heat(seeds, indptr, indices, seeds_heat, num_picks, indptr_diff):
    """ComputeFrontierHeat (src/cache/cuda/preprocess_heat.cu:35-56)."""
    return _heat(seeds, indptr, indices, None, seeds_heat, num_picks, indptr_diff)


def _CAPI_compute_frontier_heat_with_bias(seeds, indptr, indices, probs, seeds_heat, num_picks,
                                          indptr_diff):
    """ComputeFrontierHeatWithBias (src/cache/cuda/preprocess_heat.cu:100-121).  All seeds are
    processed (the reference skips the last one, :107 - a bug we do not replicate)."""
    check_device_readable(probs, "probs")
    return _heat(seeds, indptr, indices, probs, seeds_heat, num_picks, indptr_diff)


# ------------------------------------------------------------------ block construction (extension)
def coo_rows_to_indptr(coo_row, num_rows, check_sorted=False):
    """CSC row pointer (num_rows + 1 entries, id dtype) of a sampled hop whose `coo_row` is ascending.
    Every sampling entry point emits it so WHEN THE HOP'S SEEDS ARE DISTINCT (always true from the
    second hop on - the seeds are a frontier); a duplicate seed in hop 0 is relabelled to the id of
    its first occurrence, which breaks the order.  Replaces what dgl.create_block derives in the
    caller (example/graphsage/node_classification.py:18-28).  check_sorted=True verifies the
    precondition (one host sync) and raises if it does not hold; without it the result for
    unsorted rows is undefined (DistGNN.dataloading.Block checks and falls back to a sort)."""
    check_cuda(coo_row, "coo_row")
    coo_row = coo_row.contiguous()
    indptr = torch.empty(int(num_rows) + 1, dtype=coo_row.dtype, device=coo_row.device)
    flag = torch.zeros(1, dtype=torch.int32, device=coo_row.device) if check_sorted else None
    check(lib().dgs_coo_rows_to_indptr(itype(coo_row, "coo_row"), ptr(coo_row), coo_row.numel(),
                                       int(num_rows), ptr(indptr), ptr(flag) if check_sorted else None,
                                       stream()), "coo_rows_to_indptr")
    if check_sorted and int(flag.item()):
        raise RuntimeError("coo_rows_to_indptr: rows are not ascending in [0, num_rows)")
    return indptr
