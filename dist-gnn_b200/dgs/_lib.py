"""ctypes binding of the C-ABI library ``libdgs_b200.so`` (declared in ``include/dgs_b200.h``).

This is the only place that touches the shared library.  There is NO fallback: if the library is
missing or an entry fails, a RuntimeError is raised (the reference exit()s / abort()s instead,
src/common/dgs_headers.h:11-34).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGS_B200_LIB", os.path.join(_HERE, "..", "lib", "libdgs_b200.so"))

c_i64 = C.c_int64
c_vp = C.c_void_p
c_i64p = C.POINTER(C.c_int64)
c_vpp = C.POINTER(C.c_void_p)


class Graph(C.Structure):
    """dgs_graph_t"""
    _fields_ = [
        ("itype", C.c_int),
        ("etype", C.c_int),
        ("indptr", c_vp),
        ("indices", c_vp),
        ("probs", c_vp),
        ("p2p_indptr", c_vp),
        ("p2p_indices", c_vp),
        ("p2p_probs", c_vp),
        ("loc_table", c_vp),
        ("loc_capacity", c_i64),
        ("loc_mod_world", C.c_int32),
        ("num_nodes", c_i64),
    ]


class Features(C.Structure):
    """dgs_features_t"""
    _fields_ = [
        ("table", c_vp),
        ("feat", c_vp),
        ("loc_table", c_vp),
        ("loc_capacity", c_i64),
        ("mod_world", C.c_int32),
        ("row_bytes", c_i64),
        ("labels", c_vp),
        ("label_bytes", c_i64),
    ]


# name -> (restype, argtypes); must list every symbol include/dgs_b200.h declares
SIGNATURES = {
    "dgs_abi_version": (C.c_int, []),
    "dgs_last_error": (C.c_char_p, []),
    "dgs_launch_count": (c_i64, []),
    "dgs_sm_count": (C.c_int, []),
    "dgs_randn_uint64": (C.c_uint64, []),
    "dgs_seed": (None, [C.c_uint64]),
    "dgs_host_register": (C.c_int, [c_vp, C.c_size_t]),
    "dgs_host_unregister": (C.c_int, [c_vp]),
    "dgs_enable_peer_access": (C.c_int, [C.c_int]),
    "dgs_set_l2_fetch_granularity": (C.c_int, [C.c_int]),
    "dgs_nccl_get_unique_id": (C.c_int, [c_i64p]),
    "dgs_nccl_set": (C.c_int, [C.c_int, c_i64p, C.c_int]),
    "dgs_nccl_rank": (C.c_int, []),
    "dgs_nccl_world": (C.c_int, []),
    "dgs_nccl_barrier": (C.c_int, []),
    "dgs_nccl_allgather_i64": (C.c_int, [c_i64, c_i64p]),
    "dgs_nccl_allgatherv": (C.c_int, [c_vp, c_i64, c_vpp, c_i64p]),
    "dgs_p2p_server_create": (C.c_int, [c_vp, c_i64, c_vpp]),
    "dgs_p2p_server_create_virtual": (C.c_int, [C.c_int, C.c_int, c_vpp, c_i64p, c_vpp]),
    "dgs_p2p_server_ptr": (c_vp, [c_vp, C.c_int]),
    "dgs_p2p_server_nbytes": (c_i64, [c_vp, C.c_int]),
    "dgs_p2p_server_world": (C.c_int, [c_vp]),
    "dgs_p2p_server_rank": (C.c_int, [c_vp]),
    "dgs_p2p_server_destroy": (C.c_int, [c_vp, C.c_int]),
    "dgs_loc_table_capacity": (c_i64, [c_i64]),
    "dgs_loc_table_build": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_vpp, c_i64p, c_vp]),
    "dgs_loc_table_lookup": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dgs_loc_table_unpack": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "dgs_index_select": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, C.c_int, c_vp]),
    "dgs_set_gather_ctas_per_sm": (C.c_int, [C.c_int]),
    "dgs_set_gather_tile_rows": (C.c_int, [C.c_int]),
    "dgs_extract_p2p": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp,
                                  C.c_int, c_vp]),
    "dgs_extract_sharded": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, C.c_int, c_vp]),
    "dgs_route_ws_bytes": (c_i64, [c_i64, C.c_int]),
    "dgs_route_ids": (C.c_int, [C.c_int, c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "dgs_extract_dyn": (C.c_int, [c_vp, c_vp, c_vp, c_i64, C.c_int, c_i64, C.c_int, c_vp, c_i64, c_vp,
                                  c_vp, C.c_int, c_vp]),
    "dgs_coo_rows_to_indptr": (C.c_int, [C.c_int, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "dgs_extract_indptr_ws_bytes": (c_i64, [c_i64]),
    "dgs_extract_indptr": (C.c_int, [C.c_int, C.c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "dgs_extract_edge_data": (C.c_int, [C.c_int, C.c_int, C.c_int, c_vp, c_i64, c_vp, c_vp, c_vp,
                                        c_vp, c_vp]),
    "dgs_sample_ws_bytes": (c_i64, [c_i64]),
    "dgs_sample_neighbors": (C.c_int, [C.POINTER(Graph), c_vp, c_i64, c_vp, c_i64, C.c_int,
                                       C.c_uint64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dgs_debug_ares_keys": (C.c_int, [c_vp, c_i64, C.c_uint64, C.c_uint64, c_vp, c_vp]),
    "dgs_sample_blocks_ws_bytes": (c_i64, [C.c_int, c_i64, C.c_int, c_i64p, c_i64]),
    "dgs_sample_blocks_ws_init": (C.c_int, [c_vp, c_i64, C.c_int, c_i64, C.c_int, c_i64p, c_i64, c_vp]),
    "dgs_sample_blocks": (C.c_int, [C.POINTER(Graph), c_vp, c_i64, C.c_int, c_i64p, C.c_int,
                                    C.c_uint64, c_vpp, c_vpp, c_vpp, c_i64p, c_i64p, c_vp, c_vp,
                                    c_i64, c_i64, c_vp, c_vp]),
    "dgs_sample_blocks_enqueue": (C.c_int, [C.POINTER(Graph), c_vp, c_i64, C.c_int, c_i64p, C.c_int,
                                            C.c_uint64, c_vpp, c_vpp, c_vpp, c_i64p, c_i64p, c_vp, c_vp,
                                            c_i64, c_i64, c_vp, c_vp]),
    "dgs_sample_blocks_wait": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "dgs_sample_blocks_multi_ws_bytes": (c_i64, [C.c_int, C.c_int, c_i64, C.c_int, c_i64p, c_i64]),
    "dgs_sample_blocks_multi_ws_init": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_i64, C.c_int, c_i64p,
                                                  c_i64, c_vp]),
    "dgs_sample_blocks_multi": (C.c_int, [C.POINTER(Graph), C.c_int, c_vp, c_i64, c_i64, C.c_int, c_i64p,
                                          C.c_int, C.POINTER(C.c_uint64), c_vpp, c_vpp, c_vpp, c_i64,
                                          c_i64p, c_i64p, c_vp, c_vp, c_i64, c_vp, C.c_int, c_vp]),
    "dgs_load_batch": (C.c_int, [C.POINTER(Graph), C.POINTER(Features), c_vp, C.c_int, c_vp, c_i64, C.c_int,
                                 c_i64p, C.c_int, C.c_uint64, c_vp, c_i64p, c_i64p, c_i64p, c_i64, c_vp,
                                 c_i64, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, C.c_int, c_vp]),
    "dgs_relabel_table_capacity": (c_i64, [c_i64]),
    "dgs_relabel_table_bytes": (c_i64, [c_i64]),
    "dgs_relabel_ws_bytes": (c_i64, [c_i64]),
    "dgs_relabel": (C.c_int, [C.c_int, C.c_int, c_vpp, c_i64p, c_vpp, C.c_int, c_vpp, c_i64p,
                              c_vpp, c_vpp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "dgs_frontier_heat": (C.c_int, [C.c_int, C.c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_i64, c_i64, c_vp]),
}

_lib = None


def lib():
    """Load (once) and return the C-ABI library.  Fails loudly when it is not built."""
    global _lib
    if _lib is None:
        path = os.path.abspath(LIB_PATH)
        if not os.path.exists(path):
            raise RuntimeError(
                f"dgs_b200: native library not found at {path}; build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C dist-gnn_b200` "
                "(there is no CPU / PyTorch fallback)")
        l = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if l.dgs_abi_version() != 2:
            raise RuntimeError("dgs_b200: ABI version mismatch")
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dgs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"dgs_b200 {what} failed (code {rc}): {msg}")


def vp_array(ptrs):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def i64_array(vals):
    arr = (C.c_int64 * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr
