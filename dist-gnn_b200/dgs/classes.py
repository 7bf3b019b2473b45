"""``dgs.classes`` - TensorP2PServer, P2PCacheSampler, P2PCacheFeatureServer with the reference's
constructor / method names (src/pybind.cc:17-45), implemented over the sm_100a C-ABI library."""
import ctypes as C

import torch

from . import _lib, ops
from ._util import (ID_DTYPES, check_cpu, check_cuda, check_device_readable, contiguous, itype, ptr,
                    stream, tensor_from_ptr)

lib = _lib.lib
check = _lib.check
MAX_FUSED_ELEMS = 1 << 28  # ids in the worst-case arena of the fused whole-batch path (2 GiB of int64)
MAX_MANY_ELEMS = 1 << 30   # same for one B-batch launch (8 GiB of int64: friendster, b 4096, B 8 = 2.9 GiB)


class _NoCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NOCTX = _NoCtx()


def _on_device(device):
    """torch.cuda.device(device), skipped (it costs ~5 us per entry) when already current."""
    if torch.cuda.current_device() == device.index:
        return _NOCTX
    return torch.cuda.device(device)


def _barrier():
    check(lib().dgs_nccl_barrier(), "barrier")


class TensorP2PServer:
    """cache::TensorP2PServer (src/cache/tensor_p2p_cache.h:43-73, tensor_p2p_cache.cc:11-132).

    Copies a CUDA tensor into a VMM shard (cuMemCreate; the fd is passed to the peers of the box and
    cuMemMap'ped there - legacy cudaMalloc + CUDA IPC only as a fallback) so that every rank holds a
    pointer table of all shards.  Construction is collective over the NCCL context.  Unlike the reference the destructor does NOT
    run a collective (tensor_p2p_cache.cc:115 barriers inside ~TensorP2PServer, which deadlocks under
    Python GC ordering); call ``close()`` on all ranks for a synchronised teardown."""

    def __init__(self, tensor, _uninitialised_shape=None, _dtype=None, _device=None):
        self._handle = None
        if _uninitialised_shape is not None:
            shape, dtype, device = tuple(_uninitialised_shape), _dtype, _device
            src = None
        else:
            check_cuda(tensor, "tensor")
            if tensor.dim() < 1 or tensor.shape[0] <= 0:
                raise RuntimeError("TensorP2PServer: first dimension must be > 0")
            tensor = tensor.contiguous()
            shape, dtype, device = tuple(tensor.shape), tensor.dtype, tensor.device
            src = ptr(tensor)
        self._shape = shape
        self._dtype = dtype
        self._device = device
        self._esize = torch.empty(0, dtype=dtype).element_size()
        stride = 1
        for d in shape[1:]:
            stride *= d
        self.item_stride_ = stride
        nbytes = shape[0] * stride * self._esize
        if nbytes <= 0:
            raise RuntimeError("TensorP2PServer: empty tensor")
        h = C.c_void_p()
        with torch.cuda.device(device):
            torch.cuda.current_stream().synchronize()  # the source was produced on this stream
            check(lib().dgs_p2p_server_create(src, nbytes, C.byref(h)), "TensorP2PServer")
        self._handle = h
        self.world = lib().dgs_p2p_server_world(h)
        self.rank = lib().dgs_p2p_server_rank(h)

    @classmethod
    def _empty(cls, shape, dtype, device):
        """Extension: allocate the shard without a staging copy (the caller fills it in place)."""
        return cls(None, _uninitialised_shape=shape, _dtype=dtype, _device=device)

    @classmethod
    def _virtual(cls, tensors, rank):
        """Extension (tests): adopt tensors of this process as the shards of `len(tensors)`
        emulated ranks, seen from emulated rank `rank`."""
        self = cls.__new__(cls)
        t0 = tensors[rank]
        self._shape, self._dtype, self._device = tuple(t0.shape), t0.dtype, t0.device
        self._esize = t0.element_size()
        stride = 1
        for d in t0.shape[1:]:
            stride *= d
        self.item_stride_ = stride
        self._keep = [t.contiguous() for t in tensors]
        ptrs = _lib.vp_array([t.data_ptr() for t in self._keep])
        nb = _lib.i64_array([t.numel() * t.element_size() for t in self._keep])
        h = C.c_void_p()
        check(lib().dgs_p2p_server_create_virtual(len(tensors), rank, ptrs, nb, C.byref(h)),
              "TensorP2PServer._virtual")
        self._handle = h
        self.world = len(tensors)
        self.rank = rank
        return self

    @property
    def _local_ptr(self):
        return lib().dgs_p2p_server_ptr(self._handle, self.rank)

    # -- reference API
    def _CAPI_get_local_device_tensor(self):
        """GetLocalDeviceTensor (tensor_p2p_cache.cc:120-125): local shard, original shape."""
        l = lib()
        return tensor_from_ptr(l.dgs_p2p_server_ptr(self._handle, self.rank),
                               l.dgs_p2p_server_nbytes(self._handle, self.rank), self._dtype,
                               self._shape, owner=self, device=self._device)

    def _CAPI_get_device_tensor(self, device_id):
        """GetDeviceTensor (tensor_p2p_cache.cc:127-132): flat 1-D view of device_id's shard."""
        l = lib()
        if not 0 <= device_id < self.world:
            raise RuntimeError(f"device_id {device_id} out of range [0, {self.world})")
        nb = l.dgs_p2p_server_nbytes(self._handle, device_id)
        return tensor_from_ptr(l.dgs_p2p_server_ptr(self._handle, device_id), nb, self._dtype,
                               (nb // self._esize,), owner=self, device=self._device)

    def close(self, barrier=True):
        if self._handle is not None:
            h, self._handle = self._handle, None
            check(lib().dgs_p2p_server_destroy(h, 1 if barrier else 0), "TensorP2PServer.close")

    def __del__(self):
        try:
            self.close(barrier=False)
        except Exception:
            pass


def _placement_mode(local_nids, num_nodes):
    """How the ranks' cache sets tile the node ids - decided collectively (one int64 all-gather):
      "modulo"  : rank r caches exactly [r, r + P, r + 2 P, ...] in that order: node n lives on
                  GPU n % P at shard slot n // P;
      "replica" : (P > 1) every rank caches every node in id order: node n is in the LOCAL shard at
                  slot n (the reference resolves ties to the local copy too, hashmap.cu:37-42);
      "hash"    : anything else."""
    l = lib()
    world, rank = l.dgs_nccl_world(), l.dgs_nccl_rank()
    mine = 0
    if local_nids.numel() == (num_nodes - rank + world - 1) // world:
        exp = torch.arange(rank, num_nodes, world, dtype=local_nids.dtype, device=local_nids.device)
        mine = int(torch.equal(local_nids, exp))
    if world > 1 and local_nids.numel() == num_nodes:
        exp = torch.arange(num_nodes, dtype=local_nids.dtype, device=local_nids.device)
        mine = 2 * int(torch.equal(local_nids, exp))
    flags = (C.c_int64 * world)()
    check(l.dgs_nccl_allgather_i64(mine, flags), "cache placement check")
    codes = {int(f) for f in flags}
    return {(1,): "modulo", (2,): "replica"}.get(tuple(codes), "hash")


def _is_modulo_sharded(local_nids, num_nodes):
    return _placement_mode(local_nids, num_nodes) == "modulo"


def _build_loc_table(local_nids, num_nodes, device, allow_modulo=True):
    """All-gather the per-rank cached id lists and build the packed location table
    (CreateNidsP2PCacheHashMapCUDA, src/hashmap/cuda/hashmap.cu:15-77).  Returns
    (table int64[2*cap] or None, capacity, n_unique, mod_world).  No table is built when the owner
    is arithmetic: an exact modulo sharding of all nodes (mod_world = P) or a full replica on every
    rank (mod_world = -1)."""
    l = lib()
    world, rank = l.dgs_nccl_world(), l.dgs_nccl_rank()
    mode = _placement_mode(local_nids, num_nodes) if allow_modulo else "hash"
    if mode != "hash":
        return None, l.dgs_loc_table_capacity(num_nodes), num_nodes, world if mode == "modulo" else -1
    table, cap, n_unique = _hash_loc_table(local_nids, num_nodes, device)
    return table, cap, n_unique, 0


def _hash_loc_table(local_nids, num_nodes, device):
    l = lib()
    world, rank = l.dgs_nccl_world(), l.dgs_nccl_rank()
    lists = ops._allgather_tensors(local_nids)
    if world > 1:
        # |union| exactly like the reference: bool mask + count (sampler.cc:117-125)
        mask = torch.zeros(num_nodes, dtype=torch.bool, device=device)
        for t in lists:
            mask[t.long()] = True
        n_unique = int(mask.sum().item())
        del mask
    else:
        n_unique = local_nids.numel()
    cap = l.dgs_loc_table_capacity(n_unique)
    table = torch.empty(cap * 2, dtype=torch.int64, device=device)
    check(l.dgs_loc_table_build(ptr(table), cap, itype(local_nids, "cache_nids"), world, rank,
                                _lib.vp_array([ptr(t) for t in lists]),
                                _lib.i64_array([t.numel() for t in lists]), stream()),
          "location table build")
    return table, cap, n_unique


def _unpack_loc_table(table, cap, dtype):
    key = torch.empty(cap, dtype=dtype, device=table.device)
    idx = torch.empty_like(key)
    dev = torch.empty_like(key)
    check(lib().dgs_loc_table_unpack(ptr(table), cap, itype(key), ptr(key), ptr(idx), ptr(dev),
                                     stream()), "location table unpack")
    return key, idx, dev


def _check_ids_in_range(indices, num_nodes, what="indices"):
    """The fused batch path addresses its relabel tables directly by node id (8 bytes per node), so
    a neighbour id outside [0, num_nodes) would be an out-of-bounds atomic: validate once at build
    time (one reduction over the edge array; the reference never checks)."""
    if indices is None or indices.numel() == 0:
        return
    lo, hi = torch.aminmax(indices)
    lo, hi = int(lo), int(hi)
    if lo < 0 or hi >= num_nodes:
        raise RuntimeError(f"{what} holds node id {lo if lo < 0 else hi} outside [0, {num_nodes}) "
                           f"(num_nodes = len(indptr) - 1)")


def _host_source(t, name, all_cached):
    """Pointer the kernels use for cache misses (the caller's pinned CPU tensor)."""
    if t is None or t.numel() == 0:
        return None
    if t.is_pinned():
        return t.data_ptr()
    if all_cached:
        return None  # never dereferenced
    raise RuntimeError(f"{name} must be pinned (torch pin_memory or _CAPI_tensor_pin_memory): "
                       "un-cached nodes are read from it by the GPU")


class _BlockPipeline:
    """Multi-hop sample + relabel over one dgs_graph_t: every hop is enqueued with device-side
    counts (dgs_sample_blocks, 3 kernels per hop) and the host reads the 2 L counts once.
    Workspaces (incl. the two alternating relabel tables) are cached per (batch, fan-out)."""

    def __init__(self, graph, device, id_dtype):
        self._graph = graph
        self._device = device
        self._id_dtype = id_dtype
        self._plans = {}

    def _plan(self, S, fan_out):
        key = (S, tuple(fan_out))
        pl = self._plans.get(key)
        if pl is None:
            l = lib()
            L = len(fan_out)
            ubs, nnz_ubs = [], []
            ub = S
            for li in range(L):
                k = fan_out[L - 1 - li]
                ubs.append(ub)
                nnz_ubs.append(ub * k)
                ub = ub + ub * k
            used = sum(u + 3 * n for u, n in zip(ubs, nnz_ubs))
            total = (used + 1) & ~1   # keeps the int64 counts at the arena tail 8-byte aligned
            pl = {"L": L, "ubs": ubs, "nnz_ubs": nnz_ubs, "total": total, "pad": total - used,
                  "ws": None}
            if total <= MAX_FUSED_ELEMS:
                fo = _lib.i64_array(fan_out)
                it = ID_DTYPES[self._id_dtype]
                nbytes = l.dgs_sample_blocks_ws_bytes(it, S, L, fo, self._graph.num_nodes)
                if nbytes < 0:
                    raise RuntimeError("sample_blocks: " + l.dgs_last_error().decode())
                ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self._device)
                check(l.dgs_sample_blocks_ws_init(ptr(ws), nbytes, it, S, L, fo,
                                                  self._graph.num_nodes, stream()),
                      "sample_blocks_ws_init")
                es = torch.empty(0, dtype=self._id_dtype).element_size()
                offs, off = [], 0
                for u, n in zip(ubs, nnz_ubs):
                    offs.append((off, off + u + n, off + u + 2 * n))
                    off += u + 3 * n
                counts_host = torch.empty(2 * L, dtype=torch.int64).pin_memory()
                pl.update(ws=ws, ws_bytes=int(nbytes), fo=fo, epoch=0, es=es, offs=offs,
                          hop_offs=_lib.i64_array([o for t in offs for o in t]),
                          cap_edges=_lib.i64_array(nnz_ubs),
                          cap_front=_lib.i64_array([u + n for u, n in zip(ubs, nnz_ubs)]),
                          a_fr=(C.c_void_p * L)(), a_row=(C.c_void_p * L)(),
                          a_col=(C.c_void_p * L)(), count_slots=2 * L * 8 // es,
                          counts_host=counts_host, counts_np=counts_host.numpy(),
                          counts_ptr=counts_host.data_ptr())
            # workspaces are cached per (batch, fan-out); with direct-addressed relabel tables one
            # holds 16 bytes per graph node, so keep fewer of them when they are large
            big = pl["ws"] is not None and pl["ws_bytes"] > (1 << 30)
            if len(self._plans) >= (2 if big else 8):
                self._plans.clear()
            self._plans[key] = pl
        return pl

    def sample(self, seeds, fan_out, replace=False, rng_seed=None):
        l = lib()
        check_cuda(seeds, "seeds")
        if seeds.dtype != self._id_dtype:
            raise RuntimeError("seeds must have the id type of indices")
        fan_out = [int(k) for k in fan_out]
        L = len(fan_out)
        if L == 0:
            return []
        seeds = seeds.contiguous()
        if rng_seed is None:
            rng_seed = l.dgs_randn_uint64()
        if any(k < 0 for k in fan_out) or seeds.numel() == 0:
            return self._sample_per_hop(seeds, fan_out, replace, rng_seed)
        with _on_device(self._device):
            S = seeds.numel()
            pl = self._plan(S, fan_out)
            if pl["ws"] is None:
                # worst-case buffers would be unreasonable (huge fan-out used as "all neighbours"):
                # size every hop exactly instead, at the price of a host sync per hop
                return self._sample_per_hop(seeds, fan_out, replace, rng_seed)
            # one arena per batch: [frontier | row | col] per hop, worst-case sized, + 2 L counts
            arena = torch.empty(pl["total"] + pl["count_slots"], dtype=seeds.dtype, device=self._device)
            base = arena.data_ptr()
            es = pl["es"]
            a_fr, a_row, a_col = pl["a_fr"], pl["a_row"], pl["a_col"]
            for li, (of, orow, ocol) in enumerate(pl["offs"]):
                a_fr[li] = base + of * es
                a_row[li] = base + orow * es
                a_col[li] = base + ocol * es
            counts_dev = base + pl["total"] * es
            check(l.dgs_sample_blocks(
                C.byref(self._graph), seeds.data_ptr(), S, L, pl["fo"], int(bool(replace)),
                C.c_uint64(rng_seed), a_fr, a_row, a_col, pl["cap_edges"], pl["cap_front"],
                counts_dev, pl["ws"].data_ptr(), pl["ws_bytes"], pl["epoch"], pl["counts_ptr"],
                stream()), "sample_blocks")          # syncs once, counts land in pinned memory
            pl["epoch"] += 1
            counts = pl["counts_np"].tolist()
            # exact-size views of the arena in one split
            sizes = []
            for li, (u, n) in enumerate(zip(pl["ubs"], pl["nnz_ubs"])):
                nnz, nf = counts[2 * li], counts[2 * li + 1]
                sizes += [nf, u + n - nf, nnz, n - nnz, nnz, n - nnz]
            sizes.append(pl["pad"] + pl["count_slots"])
            parts = arena.split_with_sizes(sizes)
        out = []
        cur = seeds
        for li in range(L):
            frontier = parts[6 * li]
            out.append((cur, frontier, parts[6 * li + 2], parts[6 * li + 4]))
            cur = frontier
        return out

    def enqueue_only(self, seeds, fan_out, replace=False, rng_seed=1, deliver_counts=False):
        """Extension (bench.py's roofline leg, BatchLoader): enqueue one batch WITHOUT the host round
        trip and return the raw arena (worst-case sized, counts in its last 2 L int64).  Lets K
        batches be issued back to back so the kernel's own duration can be timed with CUDA events.
        deliver_counts=True: the kernel also writes the hop sizes into the plan's pinned host
        buffer; collect them with wait_counts()."""
        l = lib()
        fan_out = [int(k) for k in fan_out]
        L = len(fan_out)
        S = seeds.numel()
        pl = self._plan(S, fan_out)
        if pl["ws"] is None:
            raise RuntimeError("enqueue_only needs the fused path")
        with torch.cuda.device(self._device):
            arena = torch.empty(pl["total"] + pl["count_slots"], dtype=seeds.dtype, device=self._device)
            base, es = arena.data_ptr(), pl["es"]
            for li, (of, orow, ocol) in enumerate(pl["offs"]):
                pl["a_fr"][li] = base + of * es
                pl["a_row"][li] = base + orow * es
                pl["a_col"][li] = base + ocol * es
            fn = l.dgs_sample_blocks_enqueue if deliver_counts else l.dgs_sample_blocks
            check(fn(
                C.byref(self._graph), seeds.data_ptr(), S, L, pl["fo"], int(bool(replace)),
                C.c_uint64(rng_seed), pl["a_fr"], pl["a_row"], pl["a_col"], pl["cap_edges"],
                pl["cap_front"], base + pl["total"] * es, pl["ws"].data_ptr(), pl["ws_bytes"],
                pl["epoch"], pl["counts_ptr"] if deliver_counts else None, stream()), "sample_blocks")
            pl["epoch"] += 1
        return arena

    # ---------------------------------------------------------------- B batches per launch
    def _plan_many(self, B, S, fan_out):
        """Workspace + arena layout for B mini-batches of S seeds in one cooperative launch
        (dgs_sample_blocks_multi); None when the configuration needs the single-batch path."""
        key = ("many", B, S, tuple(fan_out))
        pl = self._plans.get(key)
        if pl is None:
            l = lib()
            L = len(fan_out)
            ubs, nnz_ubs = [], []
            ub = S
            for li in range(L):
                k = fan_out[L - 1 - li]
                ubs.append(ub)
                nnz_ubs.append(ub * k)
                ub = ub + ub * k
            used = sum(u + 3 * n for u, n in zip(ubs, nnz_ubs))
            total = (used + 3) & ~3          # per-batch stride: 16-byte aligned for both id widths
            fo = _lib.i64_array(fan_out)
            it = ID_DTYPES[self._id_dtype]
            nbytes = -1
            if B * total <= MAX_MANY_ELEMS and all(k > 0 for k in fan_out):
                nbytes = l.dgs_sample_blocks_multi_ws_bytes(it, B, S, L, fo, self._graph.num_nodes)
            if nbytes < 0:
                pl = {"ws": None}
            else:
                ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self._device)
                check(l.dgs_sample_blocks_multi_ws_init(ptr(ws), nbytes, it, B, S, L, fo,
                                                        self._graph.num_nodes, stream()),
                      "sample_blocks_multi_ws_init")
                es = torch.empty(0, dtype=self._id_dtype).element_size()
                offs, off = [], 0
                for u, n in zip(ubs, nnz_ubs):
                    offs.append((off, off + u + n, off + u + 2 * n))
                    off += u + 3 * n
                counts_host = torch.empty(B * 2 * L, dtype=torch.int64).pin_memory()
                counts_host2 = torch.empty(B * 2 * L, dtype=torch.int64).pin_memory()   # iter_many: 2 in flight
                pl = {"L": L, "B": B, "counts2": (counts_host2, counts_host2.numpy(), counts_host2.data_ptr()), "ubs": ubs, "nnz_ubs": nnz_ubs, "total": total, "pad": total - used,
                      "ws": ws, "ws_bytes": int(nbytes), "fo": fo, "es": es, "offs": offs,
                      "cap_edges": _lib.i64_array(nnz_ubs),
                      "cap_front": _lib.i64_array([u + n for u, n in zip(ubs, nnz_ubs)]),
                      "a_fr": (C.c_void_p * L)(), "a_row": (C.c_void_p * L)(), "a_col": (C.c_void_p * L)(),
                      "count_slots": B * 2 * L * 8 // es, "counts_host": counts_host,
                      "counts_np": counts_host.numpy(), "counts_ptr": counts_host.data_ptr(),
                      "rng": (C.c_uint64 * B)()}
            big = pl["ws"] is not None and pl["ws_bytes"] > (1 << 30)
            if len(self._plans) >= (2 if big else 8):
                self._plans.clear()
            self._plans[key] = pl
        return pl

    def enqueue_many(self, seeds, fan_out, replace=False, rng_seeds=None, deliver_counts=True, slot=0):
        """Enqueue B = seeds.shape[0] mini-batches (seeds: [B, S] CUDA tensor) as ONE launch.
        Returns (plan, arena); arena = B regions of plan["total"] ids, then [B][2 L] int64 counts.
        slot: which of the plan's two pinned count buffers the kernel reports to (two launches can
        be in flight when the caller pipelines, BatchLoader.iter_many)."""
        l = lib()
        fan_out = [int(k) for k in fan_out]
        B, S = int(seeds.shape[0]), int(seeds.shape[1])
        pl = self._plan_many(B, S, fan_out)
        if pl["ws"] is None:
            raise RuntimeError("enqueue_many: this configuration needs the single-batch path")
        if rng_seeds is None:
            rng_seeds = [l.dgs_randn_uint64() for _ in range(B)]
        for b in range(B):
            pl["rng"][b] = int(rng_seeds[b]) & 0xFFFFFFFFFFFFFFFF
        with _on_device(self._device):
            arena = torch.empty(B * pl["total"] + pl["count_slots"], dtype=seeds.dtype, device=self._device)
            base, es = arena.data_ptr(), pl["es"]
            for li, (of, orow, ocol) in enumerate(pl["offs"]):
                pl["a_fr"][li] = base + of * es
                pl["a_row"][li] = base + orow * es
                pl["a_col"][li] = base + ocol * es
            check(l.dgs_sample_blocks_multi(
                C.byref(self._graph), B, seeds.data_ptr(), S * es, S, pl["L"], pl["fo"],
                int(bool(replace)), pl["rng"], pl["a_fr"], pl["a_row"], pl["a_col"], pl["total"] * es,
                pl["cap_edges"], pl["cap_front"], base + B * pl["total"] * es, pl["ws"].data_ptr(),
                pl["ws_bytes"], (pl["counts2"][2] if slot else pl["counts_ptr"]) if deliver_counts else None,
                0, stream()), "sample_blocks_multi")
        return pl, arena

    def wait_many(self, pl, arena, slot=0):
        B = pl["B"]
        with _on_device(self._device):
            check(lib().dgs_sample_blocks_wait(pl["counts2"][2] if slot else pl["counts_ptr"],
                                               arena.data_ptr() + B * pl["total"] * pl["es"],
                                               pl["L"] * B, stream()), "sample_blocks_wait")

    def _views_many(self, pl, arena, seeds, slot=0):
        """Exact-size views of every batch's blocks from the hop sizes in the pinned count buffer."""
        B, L = pl["B"], pl["L"]
        counts = (pl["counts2"][1] if slot else pl["counts_np"]).tolist()
        sizes = []
        for b in range(B):
            for li, (u, n) in enumerate(zip(pl["ubs"], pl["nnz_ubs"])):
                nnz, nf = counts[b * 2 * L + 2 * li], counts[b * 2 * L + 2 * li + 1]
                sizes += [nf, u + n - nf, nnz, n - nnz, nnz, n - nnz]
            sizes.append(pl["pad"])
        sizes.append(pl["count_slots"])
        parts = arena.split_with_sizes(sizes)
        out = []
        per = 6 * L + 1
        for b in range(B):
            cur = seeds[b]
            blocks = []
            for li in range(L):
                frontier = parts[b * per + 6 * li]
                blocks.append((cur, frontier, parts[b * per + 6 * li + 2], parts[b * per + 6 * li + 4]))
                cur = frontier
            out.append(blocks)
        return out

    def sample_many(self, seeds, fan_out, replace=False, rng_seeds=None):
        """B mini-batches in one launch: seeds [B, S] (CUDA) -> list of B block lists, each identical
        to sample(seeds[b], fan_out, replace, rng_seeds[b])."""
        check_cuda(seeds, "seeds")
        if seeds.dim() != 2 or seeds.dtype != self._id_dtype:
            raise RuntimeError("sample_many: seeds must be a [B, S] tensor of the id type of indices")
        seeds = seeds.contiguous()
        fan_out = [int(k) for k in fan_out]
        B = seeds.shape[0]
        if rng_seeds is None:
            rng_seeds = [lib().dgs_randn_uint64() for _ in range(B)]
        if len(fan_out) == 0 or seeds.shape[1] == 0 or self._plan_many(B, seeds.shape[1], fan_out)["ws"] is None:
            return [self.sample(seeds[b], fan_out, replace, rng_seeds[b]) for b in range(B)]
        pl, arena = self.enqueue_many(seeds, fan_out, replace, rng_seeds)
        self.wait_many(pl, arena)
        return self._views_many(pl, arena, seeds)

    def wait_counts(self, pl, arena):
        """Block until the hop sizes of the batch enqueued with deliver_counts=True are in
        pl["counts_np"] (polls the pinned buffer the kernel writes; no stream synchronisation)."""
        with _on_device(self._device):
            check(lib().dgs_sample_blocks_wait(pl["counts_ptr"], arena.data_ptr() + pl["total"] * pl["es"],
                                               pl["L"], stream()), "sample_blocks_wait")

    def _sample_per_hop(self, seeds, fan_out, replace, rng_seed):
        """Host-synchronised hop loop (used for fan-out -1 = all neighbours, an extension)."""
        out = []
        cur = seeds
        with torch.cuda.device(self._device):
            for li, k in enumerate(reversed(fan_out)):
                key = (rng_seed + 0x9E3779B97F4A7C15 * (li + 1)) & 0xFFFFFFFFFFFFFFFF
                row, col = ops._sample_one_hop(self._graph, cur, k, replace, key)
                frontier, (rrow, rcol) = ops._relabel([cur, col], [row, col])
                out.append((cur, frontier, rrow, rcol))
                cur = frontier
        return out


class CSRSampler:
    """Extension (no reference counterpart): the fused multi-hop pipeline over an UN-cached CSR that
    lives in device memory or pinned host memory - the "no cache" configuration, which in the
    reference is a Python loop over ops._CAPI_cuda_sample_neighbors + _CAPI_cuda_sampled_tensor_
    relabel with two host syncs per hop.  Same result tuples as P2PCacheSampler."""

    def __init__(self, indptr, indices, probs=None, device=None):
        ops._check_sampling_args(torch.empty(0, dtype=indices.dtype, device="cuda"), indptr, indices,
                                 probs if probs is not None and probs.numel() else None)
        self._keep = (indptr, indices, probs)
        if device is None:
            device = indices.device if indices.is_cuda else torch.device("cuda", torch.cuda.current_device())
        _check_ids_in_range(indices, indptr.numel() - 1)
        g = ops._make_graph(indptr, indices, probs if probs is not None and probs.numel() else None)
        g.num_nodes = indptr.numel() - 1    # known node count: direct-addressed relabel tables
        self._pipe = _BlockPipeline(g, device, indices.dtype)

    def _CAPI_sample_node_classifiction(self, seeds, fan_out, replace=False, rng_seed=None):
        return self._pipe.sample(seeds, fan_out, replace, rng_seed)

    def sample_many(self, seeds, fan_out, replace=False, rng_seeds=None):
        """Extension: B mini-batches (seeds [B, S]) in ONE cooperative launch; element b equals
        _CAPI_sample_node_classifiction(seeds[b], fan_out, replace, rng_seeds[b])."""
        return self._pipe.sample_many(seeds, fan_out, replace, rng_seeds)


class P2PCacheSampler:
    """sampling::P2PCacheSampler (src/sampling/sampler.h:39-73, sampler.cc:64-196)."""

    def __init__(self, indptr, indices, probs, cache_nids, device_id):
        l = lib()
        check_cpu(indptr, "indptr")
        check_cpu(indices, "indices")
        check_cpu(probs, "probs")
        if device_id != l.dgs_nccl_rank():
            raise RuntimeError(f"device_id ({device_id}) must equal the NCCL rank "
                               f"({l.dgs_nccl_rank()})")
        if cache_nids.numel() == 0:
            raise RuntimeError("cache_nids must not be empty (use dgs.ops._CAPI_cuda_sample_"
                               "neighbors for the un-cached path)")
        contiguous(indptr, "indptr")
        contiguous(indices, "indices")
        self.device_id_ = int(device_id)
        self._device = torch.device("cuda", self.device_id_)
        self.bias_ = probs.numel() > 0
        if self.bias_ and probs.dtype != torch.float32:
            raise RuntimeError("probs must be float32")
        self.cpu_indptr_, self.cpu_indices_ = indptr, indices
        self.cpu_probs_ = probs if self.bias_ else None
        num_nodes = indptr.numel() - 1
        for t in (indptr, indices) + ((probs,) if self.bias_ else ()):
            if not t.is_pinned():
                raise RuntimeError("indptr / indices / probs must be pinned CPU tensors: the "
                                   "sub-CSR of the cached nodes is extracted by the GPU from them")
        with torch.cuda.device(self._device):
            nids = cache_nids.to(self._device).contiguous()
            if nids.dtype != indices.dtype:
                nids = nids.to(indices.dtype)
            sub_indptr = ops._Test_ExtractIndptr(nids, indptr)
            sub_indices = ops._Test_ExtractEdgeData(nids, indptr, sub_indptr, indices)
            # (un-cached rows are read from the caller's pinned arrays: check those as a whole)
            _check_ids_in_range(indices, num_nodes)
            sub_probs = ops._Test_ExtractEdgeData(nids, indptr, sub_indptr, probs) if self.bias_ else None
            self._adopt_shards(sub_indptr, sub_indices, sub_probs, nids, num_nodes)

    def _adopt_shards(self, sub_indptr, sub_indices, sub_probs, nids, num_nodes):
        """Publish the local sub-CSR as p2p shards, build the location map, make the graph handle."""
        indptr, indices, probs = self.cpu_indptr_, self.cpu_indices_, self.cpu_probs_
        with torch.cuda.device(self._device):
            self.gpu_indptr_ = TensorP2PServer(sub_indptr)
            # a cached set without edges still needs a non-empty shard for IPC
            self.gpu_indices_ = TensorP2PServer(
                sub_indices if sub_indices.numel() else torch.zeros(1, dtype=sub_indices.dtype,
                                                                    device=self._device))
            self._num_cached_edges = sub_indices.numel()
            self.gpu_probs_ = None
            if self.bias_:
                self.gpu_probs_ = TensorP2PServer(
                    sub_probs if sub_probs.numel() else torch.zeros(1, dtype=torch.float32,
                                                                    device=self._device))
            self._nids = nids
            self._num_nodes = num_nodes
            self._table, self._cap, self._n_unique, self._mod_world = _build_loc_table(
                nids, num_nodes, self._device)
            torch.cuda.current_stream().synchronize()
            _barrier()
        all_cached = self._n_unique >= num_nodes
        if self._mod_world < 0:
            # full replica: the local shards ARE the graph - same handle as an un-cached CSR in HBM
            loc_pr = self.gpu_probs_._CAPI_get_local_device_tensor() if self.bias_ else None
            self._graph = ops._make_graph(self.gpu_indptr_._CAPI_get_local_device_tensor(),
                                          self.gpu_indices_._CAPI_get_local_device_tensor(), loc_pr)
            self._graph.num_nodes = num_nodes
            self._pipe = _BlockPipeline(self._graph, self._device, sub_indices.dtype)
            return
        g = _lib.Graph()
        g.itype = itype(sub_indices, "indices")
        g.etype = itype(sub_indptr, "indptr")
        g.indptr = _host_source(indptr, "indptr", all_cached)
        g.indices = _host_source(indices, "indices", all_cached)
        g.probs = _host_source(probs, "probs", all_cached) if self.bias_ else None
        g.p2p_indptr = self.gpu_indptr_._handle
        g.p2p_indices = self.gpu_indices_._handle
        g.p2p_probs = self.gpu_probs_._handle if self.bias_ else None
        g.loc_table = ptr(self._table) if self._table is not None else None
        g.loc_capacity = self._cap
        g.loc_mod_world = self._mod_world
        g.num_nodes = num_nodes
        self._graph = g
        self._pipe = _BlockPipeline(g, self._device, sub_indices.dtype)

    @classmethod
    def from_device_shards(cls, sub_indptr, sub_indices, sub_probs, cache_nids, num_nodes,
                           device_id, cpu_indptr=None, cpu_indices=None, cpu_probs=None):
        """Extension: build the sampler from a sub-CSR that already lives on this GPU
        (e.g. generated or loaded per shard) instead of extracting it from whole-graph pinned CPU
        tensors - at papers100M / friendster scale the reference's constructor would need the full
        graph in every process.  The cpu_* tensors are only needed when some nodes are not cached
        on any GPU.  Collective like the normal constructor."""
        self = cls.__new__(cls)
        l = lib()
        if device_id != l.dgs_nccl_rank():
            raise RuntimeError(f"device_id ({device_id}) must equal the NCCL rank ({l.dgs_nccl_rank()})")
        for t_, n_ in ((sub_indptr, "sub_indptr"), (sub_indices, "sub_indices"), (cache_nids, "cache_nids")):
            check_cuda(t_, n_)
        if cache_nids.numel() == 0 or sub_indptr.numel() != cache_nids.numel() + 1:
            raise RuntimeError("sub_indptr must have len(cache_nids) + 1 entries, cache_nids non-empty")
        self.device_id_ = int(device_id)
        self._device = torch.device("cuda", self.device_id_)
        self.bias_ = sub_probs is not None and sub_probs.numel() > 0
        self.cpu_indptr_, self.cpu_indices_ = cpu_indptr, cpu_indices
        self.cpu_probs_ = cpu_probs if self.bias_ else None
        _check_ids_in_range(sub_indices, int(num_nodes), "sub_indices")
        _check_ids_in_range(cpu_indices, int(num_nodes), "cpu_indices")
        self._adopt_shards(sub_indptr.contiguous(), sub_indices.contiguous(),
                           sub_probs.contiguous() if self.bias_ else None,
                           cache_nids.contiguous().to(sub_indices.dtype), int(num_nodes))
        return self

    def _CAPI_sample_node_classifiction(self, seeds, fan_out, replace=False, rng_seed=None):
        """NodeClassifictionSample (sampler.cc:146-166): list over hops, seed-side hop first, of
        (seeds, frontier, coo_row, coo_col) with coo_* relabelled to positions in frontier."""
        return self._pipe.sample(seeds, fan_out, replace, rng_seed)

    def sample_many(self, seeds, fan_out, replace=False, rng_seeds=None):
        """Extension: B mini-batches (seeds [B, S]) in ONE cooperative launch; element b equals
        _CAPI_sample_node_classifiction(seeds[b], fan_out, replace, rng_seeds[b])."""
        return self._pipe.sample_many(seeds, fan_out, replace, rng_seeds)

    def _CAPI_get_cpu_structure_tensors(self):
        """GetCPUStructureTensors (sampler.cc:168-178); probs is None for a uniform sampler (the
        reference returns an undefined tensor)."""
        return self.cpu_indptr_, self.cpu_indices_, self.cpu_probs_

    def _CAPI_get_local_cache_structure_tensors(self):
        """GetLocalCachedStructureTensors (sampler.cc:180-191)."""
        idx = self.gpu_indices_._CAPI_get_local_device_tensor()[:self._num_cached_edges]
        pr = None
        if self.bias_:
            pr = self.gpu_probs_._CAPI_get_local_device_tensor()[:self._num_cached_edges]
        return self.gpu_indptr_._CAPI_get_local_device_tensor(), idx, pr

    def _CAPI_get_local_cache_hashmap_tensors(self):
        """GetLocalCachedHashTensors (sampler.cc:193-196): (key, idx, devid).  Synthesised from the
        packed table; the slot layout is implementation-defined (it is race-dependent in the
        reference as well), only lookups are comparable."""
        with torch.cuda.device(self._device):
            if self._table is None:   # modulo-sharded fast path: build the reference-visible table on demand
                self._table, self._cap, _ = _hash_loc_table(self._nids, self._num_nodes, self._device)
            return _unpack_loc_table(self._table, self._cap, self._nids.dtype)

    def close(self, barrier=True):
        for s in (self.gpu_indptr_, self.gpu_indices_, self.gpu_probs_):
            if s is not None:
                s.close(barrier)


class P2PCacheFeatureServer:
    """feature::P2PCacheFeatureServer (src/feature/feature_sever.h:10-34, feature_server.cc:10-86)."""

    def __init__(self, data, cache_nids, device_id):
        l = lib()
        check_cpu(data, "data")
        if device_id != l.dgs_nccl_rank():
            raise RuntimeError(f"device_id ({device_id}) must equal the NCCL rank "
                               f"({l.dgs_nccl_rank()})")
        if cache_nids.numel() == 0:
            raise RuntimeError("cache_nids must not be empty (use dgs.ops._CAPI_cuda_index_select "
                               "for the un-cached path)")
        contiguous(data, "data")
        self.device_id_ = int(device_id)
        self._device = torch.device("cuda", self.device_id_)
        self.cpu_features_ = data
        with torch.cuda.device(self._device):
            nids = cache_nids.to(self._device).contiguous()
            shape = (nids.numel(),) + tuple(data.shape[1:])
            if data.is_pinned():
                # gather the cached rows straight from pinned host memory into the shard
                # (the reference index_selects on the CPU, then copies, feature_server.cc:33-35)
                server = TensorP2PServer._empty(shape, data.dtype, self._device)
                local = server._CAPI_get_local_device_tensor()
                stride = 1
                for d in data.shape[1:]:
                    stride *= d
                check(l.dgs_index_select(ptr(data), stride * data.element_size(),
                                         itype(nids, "cache_nids"), ptr(nids), nids.numel(),
                                         ptr(local), 1, stream()), "feature shard build")
                torch.cuda.current_stream().synchronize()
            else:
                sub = data.index_select(0, cache_nids.cpu().long()).to(self._device)
                server = TensorP2PServer(sub)
            self._adopt(server, nids, data.shape[0], tuple(data.shape[1:]), data.dtype)

    def _adopt(self, server, nids, num_items, tail_shape, dtype):
        self.gpu_features_ = server
        stride = 1
        for d in tail_shape:
            stride *= d
        self._stride = stride
        self._dtype = dtype
        self._row_bytes = stride * torch.empty(0, dtype=dtype).element_size()
        self._nids = nids
        self._num_items = num_items
        with torch.cuda.device(self._device):
            self._table, self._cap, self._n_unique, self._mod_world = _build_loc_table(
                nids, num_items, self._device)
            torch.cuda.current_stream().synchronize()
            _barrier()
        self._host_ptr = _host_source(self.cpu_features_, "data", self._n_unique >= num_items)

    @classmethod
    def from_device_shard(cls, local_rows, cache_nids, num_items, device_id, cpu_data=None):
        """Extension: adopt feature rows that already live on this GPU
        (`local_rows[i]` = row of node `cache_nids[i]`); `cpu_data` (pinned) is only needed when some
        rows are cached nowhere.  Collective like the normal constructor."""
        self = cls.__new__(cls)
        l = lib()
        if device_id != l.dgs_nccl_rank():
            raise RuntimeError(f"device_id ({device_id}) must equal the NCCL rank ({l.dgs_nccl_rank()})")
        check_cuda(local_rows, "local_rows")
        check_cuda(cache_nids, "cache_nids")
        if local_rows.shape[0] != cache_nids.numel() or cache_nids.numel() == 0:
            raise RuntimeError("local_rows must hold one row per cached id (and at least one)")
        self.device_id_ = int(device_id)
        self._device = torch.device("cuda", self.device_id_)
        self.cpu_features_ = cpu_data
        with torch.cuda.device(self._device):
            server = TensorP2PServer(local_rows)
        self._adopt(server, cache_nids.contiguous(), int(num_items), tuple(local_rows.shape[1:]),
                    local_rows.dtype)
        return self

    def _CAPI_get_cpu_feature(self):
        return self.cpu_features_

    def _CAPI_get_gpu_feature(self):
        return self.gpu_features_._CAPI_get_local_device_tensor()

    def _CAPI_get_feature(self, nids, algo=0):
        """GetFeatures (feature_server.cc:69-74) -> [len(nids), prod(data.shape[1:])]."""
        check_cuda(nids, "nids")
        nids = nids.contiguous()
        n = nids.numel()
        out = torch.empty((n, self._stride), dtype=self._dtype, device=nids.device)
        if n:
            if self._mod_world < 0:    # full replica: plain gather from the local shard
                check(lib().dgs_index_select(self.gpu_features_._local_ptr, self._row_bytes,
                                             itype(nids, "nids"), ptr(nids), n, ptr(out), int(algo),
                                             stream()), "_CAPI_get_feature")
            elif self._mod_world > 0:
                check(lib().dgs_extract_sharded(self.gpu_features_._handle, self._row_bytes,
                                                itype(nids, "nids"), ptr(nids), n, ptr(out),
                                                int(algo), stream()), "_CAPI_get_feature")
            else:
                check(lib().dgs_extract_p2p(self.gpu_features_._handle, self._host_ptr,
                                            self._row_bytes, ptr(self._table), self._cap,
                                            itype(nids, "nids"), ptr(nids), n, ptr(out), int(algo),
                                            stream()), "_CAPI_get_feature")
        return out

    def get_feature_exchange(self, nids, group=None):
        """Extension (north star (4)): the same rows as _CAPI_get_feature, but the ids
        travel to the owners over NCCL (all-to-all), the owners gather from their own shard and a
        second all-to-all returns the rows - instead of in-kernel NVLink peer loads.  Modulo layout
        only (every node cached, node n on rank n % world); collective over the torch.distributed
        `group` whose ranks are the NCCL ranks of this server.  DistGNN.dist.exchange_extract."""
        check_cuda(nids, "nids")
        if self._mod_world <= 0:
            raise RuntimeError("get_feature_exchange needs the modulo-sharded layout "
                               "(cache_nids = arange(rank, N, world) on every rank)")
        from DistGNN.dist import exchange_extract
        local = self._CAPI_get_gpu_feature().reshape(-1, self._stride)
        with torch.cuda.device(self._device):
            return exchange_extract(nids, self._mod_world, self.device_id_, local, group=group)

    def _CAPI_get_local_cache_hashmap_tensors(self):
        """Extension (tests): the (key, idx, devid) view of the location table."""
        with torch.cuda.device(self._device):
            if self._table is None:
                self._table, self._cap, _ = _hash_loc_table(self._nids, self._num_items, self._device)
            return _unpack_loc_table(self._table, self._cap, self._nids.dtype)

    def close(self, barrier=True):
        self.gpu_features_.close(barrier)


class BatchLoader:
    """Extension (the whole-batch pipeline; no reference counterpart): one call per
    mini-batch = seeds (host or device) -> multi-hop blocks + input features (+ labels), with every
    kernel enqueued back to back and ONE host round trip.  The reference's training loop
    (example/graphsage/node_classification.py:219-230) makes three plugin calls with a host sync
    after sampling; here the extract reads the frontier size from device memory, so it starts the
    moment the sampling kernel ends.

      loader = BatchLoader(sampler, features, labels)       # sampler: CSRSampler / P2PCacheSampler
      blocks, x, y = loader.load(seeds, [15, 10, 5])        # features: CUDA / pinned tensor or
                                                            #           P2PCacheFeatureServer
    Results are the same tensors the three separate calls return."""

    def __init__(self, sampler, features, labels=None):
        self._pipe = sampler._pipe
        self._device = self._pipe._device
        self._fs = features if isinstance(features, P2PCacheFeatureServer) else None
        if self._fs is None:
            check_device_readable(features, "features")
            contiguous(features, "features")
            self._table = features
            stride = 1
            for d in features.shape[1:]:
                stride *= d
            self._stride, self._dtype = stride, features.dtype
            self._row_bytes = stride * features.element_size()
            self._tail = tuple(features.shape[1:])
        else:
            self._stride, self._dtype, self._row_bytes = self._fs._stride, self._fs._dtype, self._fs._row_bytes
            self._tail = (self._stride,)
        self._labels = labels
        if labels is not None:
            check_device_readable(labels, "labels")
            contiguous(labels, "labels")
        # dgs_features_t of the one-call native loader (dgs_load_batch)
        f = _lib.Features()
        if self._fs is None:
            f.table = ptr(self._table)
        elif self._fs._mod_world < 0:     # full replica: the local shard is a plain table
            f.table = self._fs.gpu_features_._local_ptr
        else:
            fs = self._fs
            f.table = fs._host_ptr
            f.feat = fs.gpu_features_._handle
            f.loc_table = ptr(fs._table) if fs._table is not None else None
            f.loc_capacity = fs._cap
            f.mod_world = max(fs._mod_world, 0)
        f.row_bytes = self._row_bytes
        if labels is not None:
            lstride = 1
            for d in labels.shape[1:]:
                lstride *= d
            f.labels = ptr(labels)
            f.label_bytes = lstride * labels.element_size()
        self._feat_struct = f

    def _extract_dyn(self, it, front_ptr, n_ub, nf_dev, out_ptr, algo):
        """Enqueue the gather of a frontier whose size is still on the device (dgs_extract_dyn)."""
        l = lib()
        if self._fs is None or self._fs._mod_world < 0:
            table = ptr(self._table) if self._fs is None else self._fs.gpu_features_._local_ptr
            check(l.dgs_extract_dyn(table, None, None, 0, 0, self._row_bytes, it, front_ptr,
                                    n_ub, nf_dev, out_ptr, int(algo), stream()), "BatchLoader extract")
        else:
            fs = self._fs
            check(l.dgs_extract_dyn(fs._host_ptr, fs.gpu_features_._handle,
                                    ptr(fs._table) if fs._table is not None else None, fs._cap,
                                    fs._mod_world, self._row_bytes, it, front_ptr, n_ub, nf_dev,
                                    out_ptr, int(algo), stream()), "BatchLoader extract")

    def enqueue_many_only(self, seeds, fan_out, replace=False, rng_seeds=None):
        """bench.py's roofline leg: the B-batch sampling launch alone, no host round trip."""
        B = seeds.shape[0]
        return self._pipe.enqueue_many(seeds, fan_out, replace, rng_seeds or list(range(1, B + 1)),
                                       deliver_counts=False)[1]

    def iter_many(self, groups, fan_out, replace=False, rng_seeds=None, algo=0, gather_ctas_per_sm=None):
        """Software-pipelined load_many over an iterable of seed groups ([B, S] each, pinned host or
        CUDA): yields, per group, the list of B (blocks, features, labels) that load_many returns.
        The sampling launch of group g + 1 is enqueued (current stream) BEFORE the host waits for
        group g, and the extracts + label gather of a group run on a side stream behind an event -
        so the latency-bound sampling kernels of one group overlap the bandwidth-bound (HBM or
        NVLink) gathers of the previous one, and the host's view building overlaps both.  The
        tensors of a yielded group are safe to use on the current stream (it waits for the group's
        extract event).  rng_seeds: optional callable g -> list of B seeds."""
        if getattr(self, "_xstream", None) is None:
            self._xstream = torch.cuda.Stream(device=self._device)
        if gather_ctas_per_sm:    # leave room on every SM for the sampling kernels of the next group
            check(lib().dgs_set_gather_ctas_per_sm(int(gather_ctas_per_sm)), "gather_ctas_per_sm")
        try:
            prev = None
            for gi, seeds in enumerate(groups):
                cur = self._enqueue_group(seeds, fan_out, replace, rng_seeds(gi) if rng_seeds else None,
                                          algo, slot=gi & 1)
                if prev is not None:
                    yield self._finish_group(prev, algo)
                prev = cur
            if prev is not None:
                yield self._finish_group(prev, algo)
        finally:
            if gather_ctas_per_sm:
                lib().dgs_set_gather_ctas_per_sm(8)

    def _enqueue_group(self, seeds, fan_out, replace, rng_seeds, algo, slot):
        l = lib()
        fan_out = [int(k) for k in fan_out]
        L = len(fan_out)
        B, S = int(seeds.shape[0]), int(seeds.shape[1])
        if rng_seeds is None:
            rng_seeds = [l.dgs_randn_uint64() for _ in range(B)]
        with _on_device(self._device):
            if not seeds.is_cuda:
                seeds = seeds.to(self._device, non_blocking=True)
            seeds = seeds.contiguous()
            pl = self._pipe._plan_many(B, S, fan_out) if (L > 0 and S > 0 and all(k > 0 for k in fan_out)) \
                else {"ws": None}
            if pl["ws"] is None:
                return {"fallback": [self.load(seeds[b], fan_out, replace, rng_seeds[b], algo) for b in range(B)]}
            pl, arena = self._pipe.enqueue_many(seeds, fan_out, replace, rng_seeds, slot=slot)
            sampled = torch.cuda.Event()
            sampled.record()
            es, base = pl["es"], arena.data_ptr()
            counts_ptr = base + B * pl["total"] * es
            n_max = pl["ubs"][-1] + pl["nnz_ubs"][-1]
            seen = pl.get("front_seen", 0)
            n_ub = n_max if seen == 0 else min(n_max, seen + seen // 4 + 1024)
            it = ID_DTYPES[seeds.dtype]
            main = torch.cuda.current_stream()
            with torch.cuda.stream(self._xstream):
                self._xstream.wait_event(sampled)
                x_all = torch.empty((B, n_ub) + self._tail, dtype=self._dtype, device=self._device)
                x_stride = n_ub * self._row_bytes
                for b in range(B):
                    front_ptr = base + (b * pl["total"] + pl["offs"][-1][0]) * es
                    nf_dev = counts_ptr + 8 * (b * 2 * L + 2 * L - 1)
                    self._extract_dyn(it, front_ptr, n_ub, nf_dev, x_all.data_ptr() + b * x_stride, algo)
                y_all = None
                if self._labels is not None:
                    y_all = ops._CAPI_cuda_index_select(self._labels, seeds.reshape(-1)).reshape(
                        (B, S) + tuple(self._labels.shape[1:]))
                done = torch.cuda.Event()
                done.record()
            # memory handed across streams: keep the allocator from recycling it too early
            arena.record_stream(self._xstream)
            seeds.record_stream(self._xstream)
            x_all.record_stream(main)
            if y_all is not None:
                y_all.record_stream(main)
        return {"pl": pl, "arena": arena, "seeds": seeds, "x_all": x_all, "y_all": y_all, "done": done,
                "n_ub": n_ub, "seen": seen, "slot": slot, "B": B, "L": L}

    def _finish_group(self, h, algo):
        if "fallback" in h:
            return h["fallback"]
        pl, arena, B, L, n_ub = h["pl"], h["arena"], h["B"], h["L"], h["n_ub"]
        with _on_device(self._device):
            self._pipe.wait_many(pl, arena, slot=h["slot"])
            blocks_all = self._pipe._views_many(pl, arena, h["seeds"], slot=h["slot"])
            counts = pl["counts2"][1] if h["slot"] else pl["counts_np"]
            torch.cuda.current_stream().wait_event(h["done"])
            out, big = [], 0
            for b in range(B):
                nf = int(counts[b * 2 * L + 2 * L - 1])
                big = max(big, nf)
                if nf > n_ub:       # the adaptive bound was too small for this batch
                    fr = blocks_all[b][-1][1]
                    x = (ops._CAPI_cuda_index_select(self._table, fr, algo) if self._fs is None
                         else self._fs._CAPI_get_feature(fr, algo))
                else:
                    x = h["x_all"][b, :nf]
                out.append((blocks_all[b], x, h["y_all"][b] if h["y_all"] is not None else None))
            pl["front_seen"] = max(pl.get("front_seen", 0), big)
        return out

    def load_many(self, seeds, fan_out, replace=False, rng_seeds=None, algo=0):
        """B mini-batches per call: seeds [B, S] (pinned host or CUDA) -> list of B
        (blocks, features, labels), element b identical to load(seeds[b], ..., rng_seed=rng_seeds[b]).
        All B batches are sampled by ONE cooperative launch (their hops share every phase and grid
        barrier), then B extracts and one label gather are enqueued behind it; one host round trip.
        This is how a training loop that knows its next B seed batches (SeedGenerator) prefetches."""
        l = lib()
        fan_out = [int(k) for k in fan_out]
        L = len(fan_out)
        if seeds.dim() != 2:
            raise RuntimeError("load_many: seeds must be [B, S]")
        B, S = int(seeds.shape[0]), int(seeds.shape[1])
        if rng_seeds is None:
            rng_seeds = [l.dgs_randn_uint64() for _ in range(B)]
        with _on_device(self._device):
            if not seeds.is_cuda:
                seeds = seeds.to(self._device, non_blocking=True)     # pinned host -> device, one copy
            seeds = seeds.contiguous()
            pl = None
            if L > 0 and S > 0 and all(k > 0 for k in fan_out):
                pl = self._pipe._plan_many(B, S, fan_out)
            if pl is None or pl["ws"] is None:       # single-batch path, B times
                return [self.load(seeds[b], fan_out, replace, rng_seeds[b], algo) for b in range(B)]
            pl, arena = self._pipe.enqueue_many(seeds, fan_out, replace, rng_seeds)
            es, base = pl["es"], arena.data_ptr()
            counts_ptr = base + B * pl["total"] * es
            n_max = pl["ubs"][-1] + pl["nnz_ubs"][-1]
            seen = pl.get("front_seen", 0)
            n_ub = n_max if seen == 0 else min(n_max, seen + seen // 4 + 1024)
            x_all = torch.empty((B, n_ub) + self._tail, dtype=self._dtype, device=self._device)
            it = ID_DTYPES[seeds.dtype]
            x_stride = n_ub * self._row_bytes
            for b in range(B):
                front_ptr = base + (b * pl["total"] + pl["offs"][-1][0]) * es
                nf_dev = counts_ptr + 8 * (b * 2 * L + 2 * L - 1)
                self._extract_dyn(it, front_ptr, n_ub, nf_dev, x_all.data_ptr() + b * x_stride, algo)
            y_all = None
            if self._labels is not None:
                y_all = ops._CAPI_cuda_index_select(self._labels, seeds.reshape(-1)).reshape(
                    (B, S) + tuple(self._labels.shape[1:]))
            self._pipe.wait_many(pl, arena)
            blocks_all = self._pipe._views_many(pl, arena, seeds)
            counts = pl["counts_np"]
            out = []
            big = 0
            for b in range(B):
                nf = int(counts[b * 2 * L + 2 * L - 1])
                big = max(big, nf)
                if nf > n_ub:       # the adaptive bound was too small for this batch
                    fr = blocks_all[b][-1][1]
                    x = (ops._CAPI_cuda_index_select(self._table, fr, algo) if self._fs is None
                         else self._fs._CAPI_get_feature(fr, algo))
                else:
                    x = x_all[b, :nf]
                out.append((blocks_all[b], x, y_all[b] if y_all is not None else None))
            pl["front_seen"] = max(seen, big)
        return out

    def load(self, seeds, fan_out, replace=False, rng_seed=None, algo=0, labels_out=None):
        """-> (blocks, features of blocks[-1][1], labels of seeds or None).  `labels_out` (pinned host
        tensor) additionally receives the labels inside the same host round trip.  One native call
        (dgs_load_batch) enqueues seeds H2D -> labels (-> D2H) -> sample -> extract and waits for the
        hop sizes (and the labels' event).  The stream is not drained: blocks / features are
        stream-ordered CUDA tensors like the result of any torch op, `labels_out` is host-valid."""
        l = lib()
        fan_out = [int(k) for k in fan_out]
        L = len(fan_out)
        if L == 0 or any(k < 0 for k in fan_out):
            raise RuntimeError("BatchLoader needs fan-outs >= 0 (use the plugin calls for -1)")
        with _on_device(self._device):
            on_host = not seeds.is_cuda
            if on_host and not seeds.is_pinned():
                seeds = seeds.to(self._device)
                on_host = False
            seeds = seeds.contiguous()
            S = seeds.numel()
            if rng_seed is None:
                rng_seed = l.dgs_randn_uint64()
            pl = self._pipe._plan(S, fan_out)
            if pl["ws"] is None or S == 0:
                raise RuntimeError("BatchLoader: batch too large for the fused path")
            if seeds.dtype != self._pipe._id_dtype:
                raise RuntimeError("seeds must have the id type of indices")
            arena = torch.empty(pl["total"] + pl["count_slots"], dtype=seeds.dtype, device=self._device)
            seeds_dev = torch.empty(S, dtype=seeds.dtype, device=self._device) if on_host else seeds
            n_max = pl["ubs"][-1] + pl["nnz_ubs"][-1]
            # output rows: the worst case the first time, then 1.25x the largest frontier seen for
            # this (batch, fan-out) - a batch that overflows is re-extracted exactly (rare)
            seen = pl.get("front_seen", 0)
            n_ub = n_max if seen == 0 else min(n_max, seen + seen // 4 + 1024)
            x = torch.empty((n_ub,) + self._tail, dtype=self._dtype, device=self._device)
            y = None
            if self._labels is not None:
                y = torch.empty((S,) + tuple(self._labels.shape[1:]), dtype=self._labels.dtype,
                                device=self._device)
            check(l.dgs_load_batch(
                C.byref(self._pipe._graph), C.byref(self._feat_struct), seeds.data_ptr(), int(on_host),
                seeds_dev.data_ptr(), S, L, pl["fo"], int(bool(replace)), C.c_uint64(rng_seed),
                arena.data_ptr(), pl["hop_offs"], pl["cap_edges"], pl["cap_front"], pl["total"],
                pl["ws"].data_ptr(), pl["ws_bytes"], pl["epoch"], pl["counts_ptr"], x.data_ptr(), n_ub,
                ptr(y) if y is not None else None,
                labels_out.data_ptr() if (labels_out is not None and y is not None) else None,
                int(algo), stream()), "load_batch")
            pl["epoch"] += 1
            counts = pl["counts_np"].tolist()
            sizes = []
            for li, (u, n) in enumerate(zip(pl["ubs"], pl["nnz_ubs"])):
                nnz, nf = counts[2 * li], counts[2 * li + 1]
                sizes += [nf, u + n - nf, nnz, n - nnz, nnz, n - nnz]
            sizes.append(pl["pad"] + pl["count_slots"])
            parts = arena.split_with_sizes(sizes)
            nf = counts[2 * L - 1]
            pl["front_seen"] = max(seen, nf)
            if nf > n_ub:       # the adaptive bound was too small for this batch
                x = (ops._CAPI_cuda_index_select(self._table, parts[6 * (L - 1)], algo)
                     if self._fs is None else self._fs._CAPI_get_feature(parts[6 * (L - 1)], algo))
            else:
                x = x[:nf]
            blocks = []
            cur = seeds_dev
            for li in range(L):
                frontier = parts[6 * li]
                blocks.append((cur, frontier, parts[6 * li + 2], parts[6 * li + 4]))
                cur = frontier
        return blocks, x, y
