"""numpy front-end of oracle/dgs_oracle.c (the CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package dist-gnn_b200/.
Pinning status and reference file:line citations are in the header of dgs_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build():
    src = os.path.join(_HERE, "dgs_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_up_power.restype = C.c_int64
        _lib.orc_up_power.argtypes = [C.c_int64]
        _lib.orc_hashmap_capacity.restype = C.c_int64
        _lib.orc_hashmap_capacity.argtypes = [C.c_int64]
        _lib.orc_sample_copy_path.restype = C.c_int64
        _lib.orc_relabel.restype = C.c_int64
        _lib.orc_cpu_sample_neighbors.restype = C.c_int64
        _lib.orc_cpu_relabel.restype = C.c_int64
        _lib.orc_cpu_batch.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n=0):
    """OpenMP threads of the CPU baseline (n <= 0: every host core this process may use).  torchrun
    sets OMP_NUM_THREADS=1 for its workers, so bench.py's CPU arms call this first."""
    if n <= 0:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def index_select(table, nids):
    """feature_ops.cu:140-210 - out[i] = table[nids[i]] (bytes are copied, any dtype)."""
    table = np.ascontiguousarray(table)
    nids = _i64(nids)
    out = np.empty((len(nids),) + table.shape[1:], dtype=table.dtype)
    rb = table.dtype.itemsize * int(np.prod(table.shape[1:], dtype=np.int64))
    lib().orc_index_select(_p(table), C.c_int64(rb), _p(nids), C.c_int64(len(nids)), _p(out))
    return out


def hashmap_capacity(n_unique):
    return int(lib().orc_hashmap_capacity(int(n_unique)))


def hashmap_build(dev_nids, rank, n_unique=None):
    """hashmap.cu:15-77 - returns (key, idx, devid) as the reference lays them out when the inserts
    of one device run sequentially."""
    lists = [_i64(x) for x in dev_nids]
    world = len(lists)
    if n_unique is None:
        n_unique = len(np.unique(np.concatenate(lists)))
    cap = hashmap_capacity(n_unique)
    key = np.empty(cap, np.int64)
    idx = np.empty(cap, np.int64)
    dev = np.empty(cap, np.int64)
    ptrs = (C.c_void_p * world)(*[x.ctypes.data for x in lists])
    counts = (C.c_int64 * world)(*[len(x) for x in lists])
    lib().orc_hashmap_build(_p(key), _p(idx), _p(dev), C.c_int64(cap), C.c_int(world), C.c_int(rank),
                            ptrs, counts)
    return key, idx, dev


def hashmap_lookup(key, idx, dev, nids):
    nids = _i64(nids)
    od = np.empty(len(nids), np.int64)
    oi = np.empty(len(nids), np.int64)
    lib().orc_hashmap_lookup(_p(key), _p(idx), _p(dev), C.c_int64(len(key)), _p(nids),
                             C.c_int64(len(nids)), _p(od), _p(oi))
    return od, oi


def extract_p2p(cpu_table, shards, key, idx, dev, nids):
    """feature_ops.cu:38-138 - cached gather from `shards[devid][idx]` or the host table."""
    cpu_table = np.ascontiguousarray(cpu_table)
    shards = [np.ascontiguousarray(s) for s in shards]
    nids = _i64(nids)
    rb = cpu_table.dtype.itemsize * int(np.prod(cpu_table.shape[1:], dtype=np.int64))
    out = np.empty((len(nids), rb // cpu_table.dtype.itemsize), dtype=cpu_table.dtype)
    ptrs = (C.c_void_p * len(shards))(*[s.ctypes.data for s in shards])
    lib().orc_extract_p2p(_p(cpu_table), ptrs, C.c_int64(rb), _p(key), _p(idx), _p(dev),
                          C.c_int64(len(key)), _p(nids), C.c_int64(len(nids)), _p(out))
    return out


def extract_indptr(nids, indptr):
    nids, indptr = _i64(nids), _i64(indptr)
    sub = np.empty(len(nids) + 1, np.int64)
    lib().orc_extract_indptr(_p(nids), C.c_int64(len(nids)), _p(indptr), _p(sub))
    return sub


def extract_edge_data(nids, indptr, sub_indptr, edge_data):
    nids, indptr, sub_indptr = _i64(nids), _i64(indptr), _i64(sub_indptr)
    edge_data = np.ascontiguousarray(edge_data)
    out = np.empty(int(sub_indptr[-1]), dtype=edge_data.dtype)
    lib().orc_extract_edge_data(_p(nids), C.c_int64(len(nids)), _p(indptr), _p(sub_indptr),
                                _p(edge_data), C.c_int64(edge_data.dtype.itemsize), _p(out))
    return out


def sample_all_neighbors(seeds, indptr, indices):
    """rowwise_sampling.cu:71-77 (+ offsets of :16-45): every neighbour, CSR order, seed-major."""
    seeds, indptr, indices = _i64(seeds), _i64(indptr), _i64(indices)
    nnz = int(lib().orc_sample_copy_path(_p(seeds), C.c_int64(len(seeds)), _p(indptr), _p(indices),
                                         None, None))
    row = np.empty(nnz, np.int64)
    col = np.empty(nnz, np.int64)
    lib().orc_sample_copy_path(_p(seeds), C.c_int64(len(seeds)), _p(indptr), _p(indices), _p(row),
                               _p(col))
    return row, col


def relabel(mapping_tensors, to_relabel_tensors):
    """tensor_relabel.cu:182-205 - (unique, [relabelled...])."""
    mapping = _i64(np.concatenate([np.asarray(m).reshape(-1) for m in mapping_tensors]))
    sizes = [int(np.asarray(t).size) for t in to_relabel_tensors]
    rel = _i64(np.concatenate([np.asarray(t).reshape(-1) for t in to_relabel_tensors])
               if sizes else np.empty(0, np.int64))
    unique = np.empty(len(mapping), np.int64)
    out = np.empty(len(rel), np.int64)
    u = int(lib().orc_relabel(_p(mapping), C.c_int64(len(mapping)), _p(rel), C.c_int64(len(rel)),
                              _p(unique), _p(out)))
    outs, o = [], 0
    for s in sizes:
        outs.append(out[o:o + s].copy())
        o += s
    return unique[:u].copy(), outs


def sample_blocks_all_neighbors(seeds, indptr, indices, num_layers):
    """sampler.cc:14-36 with every hop on the copy path: list of (seeds, frontier, row, col)."""
    out = []
    cur = _i64(seeds)
    for _ in range(num_layers):
        row, col = sample_all_neighbors(cur, indptr, indices)
        frontier, (rrow, rcol) = relabel([cur, col], [row, col])
        out.append((cur, frontier, rrow, rcol))
        cur = frontier
    return out


def frontier_heat(seeds, indptr, indices, probs, seeds_heat, num_picks, indptr_diff=0):
    seeds, indptr, indices = _i64(seeds), _i64(indptr), _i64(indices)
    seeds_heat = np.ascontiguousarray(seeds_heat, np.float32)
    pr = np.ascontiguousarray(probs, np.float32) if probs is not None else None
    out = np.zeros_like(seeds_heat)
    lib().orc_frontier_heat(_p(seeds), C.c_int64(len(seeds)), _p(indptr), _p(indices), _p(pr),
                            _p(seeds_heat), _p(out), C.c_int64(num_picks), C.c_int64(indptr_diff))
    return out


# ------------------------------------------------------------------ CPU baseline (DGL semantics)
def cpu_sample_neighbors(seeds, indptr, indices, probs, k, rng_seed=0):
    seeds, indptr, indices = _i64(seeds), _i64(indptr), _i64(indices)
    pr = np.ascontiguousarray(probs, np.float32) if probs is not None else None
    n = len(seeds)
    offsets = np.empty(n + 1, np.int64)
    nnz = int(lib().orc_cpu_sample_neighbors(_p(seeds), C.c_int64(n), _p(indptr), _p(indices), _p(pr),
                                             C.c_int64(k), C.c_uint64(rng_seed), _p(offsets), None,
                                             None))
    row = np.empty(nnz, np.int64)
    col = np.empty(nnz, np.int64)
    lib().orc_cpu_sample_neighbors(_p(seeds), C.c_int64(n), _p(indptr), _p(indices), _p(pr),
                                   C.c_int64(k), C.c_uint64(rng_seed), _p(offsets), _p(row), _p(col))
    return row, col


class CpuBatchRunner:
    """Pre-allocated buffers for orc_cpu_batch (the timed CPU baseline step)."""

    def __init__(self, indptr, indices, probs, feat, batch, fan_out):
        self.indptr, self.indices = _i64(indptr), _i64(indices)
        self.probs = np.ascontiguousarray(probs, np.float32) if probs is not None else None
        self.feat = np.ascontiguousarray(feat) if feat is not None else None
        self.fan_out = _i64(fan_out)
        ub = batch
        ubs = []
        for k in reversed(list(fan_out)):
            ubs.append(ub)
            ub = ub * (1 + int(k))
        self.ub_last = ub
        nnz_max = max(u * int(k) for u, k in zip(ubs, reversed(list(fan_out))))
        self.row = np.empty(nnz_max, np.int64)
        self.col = np.empty(nnz_max, np.int64)
        self.fa = np.empty(ub, np.int64)
        self.fb = np.empty(ub, np.int64)
        self.offsets = np.empty(max(ubs) + 1, np.int64)
        self.scratch = np.full(len(self.indptr) - 1, -1, np.int64)
        self.row_bytes = 0
        self.feat_out = None
        if self.feat is not None:
            self.row_bytes = self.feat.dtype.itemsize * int(np.prod(self.feat.shape[1:]))
            self.feat_out = np.empty(ub * self.row_bytes, np.uint8)

    def run(self, seeds, rng_seed=0):
        seeds = _i64(seeds)
        rows = C.c_int64(0)
        edges = lib().orc_cpu_batch(
            _p(seeds), C.c_int64(len(seeds)), _p(self.indptr), _p(self.indices), _p(self.probs),
            _p(self.fan_out), C.c_int(len(self.fan_out)), C.c_uint64(rng_seed), _p(self.feat),
            C.c_int64(self.row_bytes), _p(self.row), _p(self.col), _p(self.fa), _p(self.fb),
            _p(self.offsets), _p(self.scratch), _p(self.feat_out), C.byref(rows))
        return int(edges), int(rows.value)
