"""CPU: the C-ABI library loads without a GPU, exports every symbol include/dgs_b200.h declares,
and its host-only entries / argument checks behave (no kernel is launched here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dgs_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgs_[a-z0-9_]+)\s*\(", src)))


def test_header_is_plain_c_and_cxx():
    """The drop-in boundary must be consumable from C (cgo / JNI style FFI) and C++ (the reference's
    pybind module): include/dgs_b200.h compiles stand-alone in both languages."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None or shutil.which("g++") is None:
        pytest.skip("no host compiler")
    for cc, lang, std in (("gcc", "c", "-std=c11"), ("g++", "c++", "-std=c++17")):
        p = subprocess.run([cc, std, "-Wall", "-Werror", "-fsyntax-only", "-x", lang, HEADER],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr


def test_header_symbols_exported():
    from dgs import _lib
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dgs_b200.h but not exported"
    # and the binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_channel():
    from dgs import _lib
    lib = _lib.lib()
    assert lib.dgs_abi_version() == 2
    assert lib.dgs_launch_count() == 0 or lib.dgs_launch_count() > 0
    # argument error: reported through the return code + dgs_last_error, nothing launched
    before = lib.dgs_launch_count()
    rc = lib.dgs_index_select(None, 0, 1, None, 5, None, 0, None)
    assert rc != 0 and b"bad sizes" in lib.dgs_last_error()
    rc = lib.dgs_relabel(1, 9, None, None, None, 0, None, None, None, None, None, None, None, 64, None, None)
    assert rc != 0 and b"mapping parts" in lib.dgs_last_error()
    rc = lib.dgs_loc_table_build(None, 8, 1, 1, 0, None, None, None)
    assert rc != 0
    with pytest.raises(RuntimeError, match="dgs_b200 x failed"):
        _lib.check(rc, "x")
    assert lib.dgs_launch_count() == before


def test_random_engine_seedable():
    import dgs
    vals = {dgs.ops._Test_Randn() for _ in range(10)}
    assert len(vals) == 10  # tests/test_random_engine.py:5-8 prints 10 values
    dgs.ops.seed(42)
    a = [dgs.ops._Test_Randn() for _ in range(4)]
    dgs.ops.seed(42)
    assert a == [dgs.ops._Test_Randn() for _ in range(4)]
    assert all(0 <= v < 2 ** 64 for v in a)


def test_capacity_and_workspace_sizes():
    import oracle
    from dgs import _lib
    lib = _lib.lib()
    for n in [1, 2, 3, 4, 5, 7, 8, 9, 1000, 1024, 1025, 2449029, 111059956]:
        assert lib.dgs_loc_table_capacity(n) == oracle.hashmap_capacity(n)
    for n in [1, 31, 32, 33, 1000, 1 << 20]:
        cap = lib.dgs_relabel_table_capacity(n)
        assert cap >= 2 * n and cap & (cap - 1) == 0
        assert lib.dgs_relabel_table_bytes(n) == 16 * cap
        assert lib.dgs_relabel_ws_bytes(n) >= 8 * n
        assert lib.dgs_sample_ws_bytes(n) >= 24 * n
        assert lib.dgs_extract_indptr_ws_bytes(n) >= 256


def test_batch_workspace_planning_without_gpu():
    """Host-only logic of the whole-batch sampler: which kernel path a workspace is laid out for,
    how it scales with the batches per launch, and the registry check that refuses a workspace the
    library has never initialised - no kernel is launched."""
    from dgs import _lib
    lib = _lib.lib()
    fo = _lib.i64_array([15, 10, 5])
    N = 2449029
    one = lib.dgs_sample_blocks_multi_ws_bytes(1, 1, 1024, 3, fo, N)
    assert one > 4 * N                                   # the 4-byte-per-node table + per-slot arrays
    assert one == lib.dgs_sample_blocks_ws_bytes(1, 1024, 3, fo, N)     # B = 1 is the same layout
    assert lib.dgs_sample_blocks_multi_ws_bytes(1, 8, 1024, 3, fo, N) == 8 * one
    assert lib.dgs_sample_blocks_multi_ws_bytes(0, 1, 1024, 3, fo, N) < one      # int32 ids: smaller slots
    # no multi-batch path: unknown node count, fan-out 0, too many batches, items of a hop >= 2^24
    assert lib.dgs_sample_blocks_multi_ws_bytes(1, 1, 1024, 3, fo, 0) == -1
    assert b"single-batch path" in lib.dgs_last_error()
    assert lib.dgs_sample_blocks_multi_ws_bytes(1, 1, 1024, 2, _lib.i64_array([5, 0]), N) == -1
    assert lib.dgs_sample_blocks_multi_ws_bytes(1, 17, 1024, 3, fo, N) == -1
    assert lib.dgs_sample_blocks_multi_ws_bytes(1, 1, 8192, 3, _lib.i64_array([20, 15, 10]), N) == -1
    assert lib.dgs_sample_blocks_multi_ws_bytes(1, 1, 4096, 3, _lib.i64_array([20, 15, 10]), N) > 0   # friendster
    # ... while the single-batch entry still sizes a (hashed / wiped-table) workspace for them
    assert lib.dgs_sample_blocks_ws_bytes(1, 1024, 3, fo, 0) > 0
    assert lib.dgs_sample_blocks_ws_bytes(1, 1024, 2, _lib.i64_array([5, 0]), N) > 0
    # a workspace pointer the library never initialised is refused before anything is launched
    before = lib.dgs_launch_count()
    g = _lib.Graph()
    g.itype = g.etype = 1
    g.num_nodes = N
    buf = (C.c_char * 64)()
    arr = (C.c_void_p * 3)(1, 1, 1)
    caps = _lib.i64_array([1 << 30] * 3)
    rc = lib.dgs_sample_blocks(C.byref(g), C.addressof(buf), 1024, 3, fo, 0, C.c_uint64(1), arr, arr, arr, caps,
                               caps, C.addressof(buf), C.addressof(buf), one, 0, None, None)
    assert rc != 0 and b"never initialised" in lib.dgs_last_error()
    rng = (C.c_uint64 * 2)(1, 2)
    rc = lib.dgs_sample_blocks_multi(C.byref(g), 2, C.addressof(buf), 8192, 1024, 3, fo, 0, rng, arr, arr, arr, 0,
                                     caps, caps, C.addressof(buf), C.addressof(buf), one, None, 0, None)
    assert rc != 0 and b"never initialised" in lib.dgs_last_error()
    rc = lib.dgs_load_batch(None, None, None, 0, None, 0, 0, None, 0, C.c_uint64(0), None, None, None, None, 0,
                            None, 0, 0, None, None, 0, None, None, 0, None)
    assert rc != 0 and b"null argument" in lib.dgs_last_error()
    assert lib.dgs_set_gather_ctas_per_sm(0) != 0 and lib.dgs_set_gather_ctas_per_sm(8) == 0
    assert lib.dgs_set_gather_tile_rows(65) != 0 and lib.dgs_set_gather_tile_rows(-1) != 0
    assert lib.dgs_set_gather_tile_rows(32) == 0 and lib.dgs_set_gather_tile_rows(0) == 0
    # request routing of the id-exchange extract: one 4-byte counter per (owner, 2048-id chunk);
    # bad arguments are refused before anything is launched
    assert lib.dgs_route_ws_bytes(0, 8) == 8 * 4 and lib.dgs_route_ws_bytes(2048, 8) == 8 * 4
    assert lib.dgs_route_ws_bytes(2049, 3) == 2 * 3 * 4
    assert lib.dgs_route_ws_bytes(100, 0) == -1 and lib.dgs_route_ws_bytes(100, 17) == -1
    assert lib.dgs_route_ws_bytes(-1, 2) == -1
    a = C.addressof(buf)
    assert lib.dgs_route_ids(1, a, 10, 0, a, a, a, a, 64, None) != 0 and b"world" in lib.dgs_last_error()
    assert lib.dgs_route_ids(1, a, 10, 2, a, a, None, a, 64, None) != 0 and b"null" in lib.dgs_last_error()
    assert lib.dgs_route_ids(1, a, 100000, 2, a, a, a, a, 64, None) != 0 and b"workspace" in lib.dgs_last_error()
    assert lib.dgs_route_ids(0, a, 1 << 31, 2, a, a, a, a, 1 << 40, None) != 0 and b"int32" in lib.dgs_last_error()
    assert lib.dgs_launch_count() == before


def test_nccl_context_defaults():
    import dgs
    assert dgs.ops._Test_GetWorldSize() == 1
    assert dgs.ops._Test_GetLocalRank() == 0


def test_python_argument_checks_without_gpu():
    import torch
    import dgs
    cpu = torch.arange(4)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        dgs.ops._CAPI_cuda_index_select(torch.zeros(4, 2), cpu)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        dgs.ops._CAPI_cuda_sample_neighbors(cpu, cpu, cpu, 2, False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dgs.classes.TensorP2PServer(cpu)
    with pytest.raises(RuntimeError, match="must not be empty"):
        dgs.classes.P2PCacheSampler(cpu, cpu, torch.Tensor(), torch.empty(0, dtype=torch.int64), 0)
    with pytest.raises(RuntimeError, match="must equal the NCCL rank"):
        dgs.classes.P2PCacheFeatureServer(torch.zeros(4, 2), cpu, 3)
    with pytest.raises(RuntimeError, match="16 int64"):
        dgs.ops._CAPI_set_nccl(1, [0, 1], 0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dgs import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "dist-gnn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "dgs_oracle" not in txt, f
