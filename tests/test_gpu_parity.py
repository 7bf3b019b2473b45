"""GPU: the sm_100a path (through the C-ABI, via the dgs plugin API) against the CPU oracle.

Bit-exact: extract / index_select, location-table lookups, sub-CSR extraction, full-neighbour
sampling, relabel, multi-hop blocks on the copy path.  Random sampling: validity, no duplicates
without replacement, exact counts, reproducibility, and chi-square tests (tolerance stated in
each test)."""
import numpy as np
import pytest
import torch

import dgs_synth
import oracle

pytestmark = pytest.mark.gpu


def t2n(t):
    return t.detach().cpu().numpy()


def small_graph(N=3000, E=60000, seed=0, weights=False, dtype=torch.int64):
    indptr, indices, probs = dgs_synth.make_csr(N, E, seed=seed, weights=weights, id_dtype=dtype)
    return indptr, indices, probs


# ------------------------------------------------------------------ extract
@pytest.mark.parametrize("algo", [1, 2, 3, 7])
@pytest.mark.parametrize("dim,dtype", [(100, torch.float32), (128, torch.float32),
                                       (256, torch.bfloat16), (4, torch.float32),
                                       (36, torch.int64), (260, torch.float32)])
def test_index_select_bit_exact(dgs, cuda, algo, dim, dtype):
    N, R = 20000, 70001
    ids = torch.arange(N)
    feat = dgs_synth.feature_rows(ids, dim, dtype if dtype != torch.int64 else torch.int64)
    g = torch.Generator().manual_seed(dim)
    nids = torch.randint(0, N, (R,), generator=g)
    out = dgs.ops._CAPI_cuda_index_select(feat.to(cuda), nids.to(cuda), algo)
    assert out.shape == (R, dim) and out.dtype == feat.dtype
    exp = oracle.index_select(t2n(feat.view(torch.int16) if dtype == torch.bfloat16 else feat), t2n(nids))
    got = t2n(out.view(torch.int16) if dtype == torch.bfloat16 else out)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("dim,dtype", [(1, torch.int64), (1, torch.float32), (10, torch.float32),
                                       (3, torch.int32), (7, torch.float16)])
def test_index_select_unaligned_rows(dgs, cuda, dim, dtype):
    """row sizes that are not multiples of 16 bytes, 1-D tables (labels), int ids of both widths.
    (The reference's 1-D path truncates floats through an integer temp, feature_ops.cu:26-34 -
    a bug; we return the values unchanged.)"""
    N, R = 999, 5003
    g = torch.Generator().manual_seed(3)
    if dtype.is_floating_point:
        table = torch.randn(N, dim, generator=g).to(dtype)
    else:
        table = torch.randint(-1000, 1000, (N, dim), generator=g, dtype=dtype)
    if dim == 1:
        table = table.reshape(N)
    for idt in (torch.int64, torch.int32):
        nids = torch.randint(0, N, (R,), generator=g).to(idt)
        out = dgs.ops._CAPI_cuda_index_select(table.to(cuda), nids.to(cuda))
        assert out.shape == (R,) + tuple(table.shape[1:])
        assert torch.equal(out.cpu(), table[nids.long()])


def test_index_select_edge_cases(dgs, cuda):
    feat = torch.arange(100).float().reshape(10, 10)
    empty = dgs.ops._CAPI_cuda_index_select(feat.to(cuda), torch.empty(0, dtype=torch.int64, device=cuda))
    assert empty.shape == (0, 10)
    one = dgs.ops._CAPI_cuda_index_select(feat.to(cuda), torch.tensor([9], device=cuda))
    assert torch.equal(one.cpu(), feat[9:10])
    # repeated ids, pinned-host table (labels path of the training loop, node_classification.py:228)
    pinned = feat.pin_memory()
    nids = torch.tensor([7, 7, 0, 9, 7] * 50, device=cuda)
    out = dgs.ops._CAPI_cuda_index_select(pinned, nids)
    assert torch.equal(out.cpu(), feat[nids.cpu()])
    with pytest.raises(RuntimeError, match="neither pinned"):
        dgs.ops._CAPI_cuda_index_select(feat, nids)
    with pytest.raises(RuntimeError, match="int32 or int64"):
        dgs.ops._CAPI_cuda_index_select(feat.to(cuda), nids.to(torch.int16))


def test_index_select_3d_table(dgs, cuda):
    table = torch.arange(6 * 4 * 5, dtype=torch.float32).reshape(6, 4, 5)
    nids = torch.tensor([5, 0, 3], device=cuda)
    out = dgs.ops._CAPI_cuda_index_select(table.to(cuda), nids)
    assert out.shape == (3, 4, 5) and torch.equal(out.cpu(), table[nids.cpu()])


# ------------------------------------------------------------------ location table + cached extract
def _virtual_feature_setup(dgs, cuda, N, D, P, per, seed, dtype=torch.float32):
    rng = np.random.default_rng(seed)
    feat = dgs_synth.feature_rows(torch.arange(N), D, dtype)
    lists = [np.sort(rng.choice(N, per, replace=False)) for _ in range(P)]
    shards = [feat[torch.from_numpy(l)].to(cuda) for l in lists]
    return feat, lists, shards


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_loc_table_lookup_matches_oracle(dgs, cuda, P):
    from dgs import _lib
    from dgs._util import ptr, stream
    lib = _lib.lib()
    N, per = 50000, 6000
    rng = np.random.default_rng(P)
    lists = [rng.choice(N, per, replace=False) for _ in range(P)]
    n_unique = len(np.unique(np.concatenate(lists)))
    cap = lib.dgs_loc_table_capacity(n_unique)
    assert cap == oracle.hashmap_capacity(n_unique)
    q = rng.integers(0, N, 30000)
    dl = [torch.from_numpy(l).to(cuda) for l in lists]
    qd = torch.from_numpy(q).to(cuda)
    for rank in range(P):
        table = torch.empty(2 * cap, dtype=torch.int64, device=cuda)
        _lib.check(lib.dgs_loc_table_build(ptr(table), cap, 1, P, rank,
                                           _lib.vp_array([ptr(t) for t in dl]),
                                           _lib.i64_array([per] * P), stream()))
        od = torch.empty_like(qd)
        oi = torch.empty_like(qd)
        _lib.check(lib.dgs_loc_table_lookup(ptr(table), cap, 1, ptr(qd), len(q), ptr(od), ptr(oi), stream()))
        key, idx, dev = oracle.hashmap_build(lists, rank, n_unique)
        ed, ei = oracle.hashmap_lookup(key, idx, dev, q)
        assert np.array_equal(t2n(od), ed) and np.array_equal(t2n(oi), ei)
        # unpacked (key, idx, devid) tensors hold the same set of entries as the reference layout
        k2 = torch.empty(cap, dtype=torch.int64, device=cuda)
        i2 = torch.empty_like(k2)
        d2 = torch.empty_like(k2)
        _lib.check(lib.dgs_loc_table_unpack(ptr(table), cap, 1, ptr(k2), ptr(i2), ptr(d2), stream()))
        ours = sorted(zip(t2n(k2)[t2n(k2) >= 0].tolist(), t2n(d2)[t2n(k2) >= 0].tolist(), t2n(i2)[t2n(k2) >= 0].tolist()))
        refs = sorted(zip(key[key >= 0].tolist(), dev[key >= 0].tolist(), idx[key >= 0].tolist()))
        assert ours == refs
        assert (t2n(i2)[t2n(k2) < 0] == -1).all() and (t2n(d2)[t2n(k2) < 0] == -1).all()


@pytest.mark.parametrize("P", [4, 8])
def test_loc_table_identical_cache_lists(dgs, cuda, P):
    """Every rank caches the SAME hot nodes (what get_cache_nids_selfish / the "selfish" policy
    produce): sum(counts) = P * n but only n distinct keys, which is what the reference sizes the
    table for (hashmap.cu:20; Update() overwrites, hashmap.h:18-32).  The build must accept it and
    resolve every id to the local rank."""
    from dgs import _lib
    from dgs._util import ptr, stream
    lib = _lib.lib()
    N, n = 100000, 20000
    rng = np.random.default_rng(P)
    hot = rng.choice(N, n, replace=False)
    cap = lib.dgs_loc_table_capacity(n)
    assert P * n > cap                     # 80 000 / 160 000 listed ids > 65 536 slots, 20 000 distinct
    dl = [torch.from_numpy(hot).to(cuda) for _ in range(P)]
    q = rng.integers(0, N, 50000)
    qd = torch.from_numpy(q).to(cuda)
    for rank in (0, P - 1):
        table = torch.empty(2 * cap, dtype=torch.int64, device=cuda)
        _lib.check(lib.dgs_loc_table_build(ptr(table), cap, 1, P, rank, _lib.vp_array([ptr(t) for t in dl]),
                                           _lib.i64_array([n] * P), stream()))
        od, oi = torch.empty_like(qd), torch.empty_like(qd)
        _lib.check(lib.dgs_loc_table_lookup(ptr(table), cap, 1, ptr(qd), len(q), ptr(od), ptr(oi), stream()))
        key, idx, dev = oracle.hashmap_build([hot] * P, rank, n)
        ed, ei = oracle.hashmap_lookup(key, idx, dev, q)
        assert np.array_equal(t2n(od), ed) and np.array_equal(t2n(oi), ei)
        assert set(t2n(od).tolist()) <= {-1, rank}
    # too many DISTINCT ids for the capacity is still an error (reported, not a hang)
    many = [torch.arange(i * 3000, (i + 1) * 3000, device=cuda) for i in range(4)]
    small = lib.dgs_loc_table_capacity(4000)      # 8192 slots < 12 000 distinct ids
    table = torch.empty(2 * small, dtype=torch.int64, device=cuda)
    rc = lib.dgs_loc_table_build(ptr(table), small, 1, 4, 0, _lib.vp_array([ptr(t) for t in many]),
                                 _lib.i64_array([3000] * 4), stream())
    assert rc != 0 and b"distinct" in lib.dgs_last_error()


@pytest.mark.parametrize("algo", [0, 1, 2, 3, 7])
@pytest.mark.parametrize("P,rank", [(1, 0), (2, 1), (4, 2), (8, 7)])
def test_extract_p2p_virtual_ranks_bit_exact(dgs, cuda, P, rank, algo):
    """Multi-rank cached extract emulated on one GPU: every emulated rank's shard is a tensor of this
    process, the location table is built from rank `rank`'s point of view (local wins)."""
    from dgs import _lib
    from dgs._util import ptr, stream
    lib = _lib.lib()
    N, D, per = 30000, 100, 3000
    feat, lists, shards = _virtual_feature_setup(dgs, cuda, N, D, P, per, seed=10 * P + rank)
    server = dgs.classes.TensorP2PServer._virtual(shards, rank)
    n_unique = len(np.unique(np.concatenate(lists)))
    cap = lib.dgs_loc_table_capacity(n_unique)
    table = torch.empty(2 * cap, dtype=torch.int64, device=cuda)
    dl = [torch.from_numpy(l).to(cuda) for l in lists]
    _lib.check(lib.dgs_loc_table_build(ptr(table), cap, 1, P, rank, _lib.vp_array([ptr(t) for t in dl]),
                                       _lib.i64_array([per] * P), stream()))
    q = torch.randint(0, N, (40000,), generator=torch.Generator().manual_seed(1))
    host = feat.pin_memory()
    out = torch.empty((len(q), D), dtype=feat.dtype, device=cuda)
    qd = q.to(cuda)
    if algo == 2:
        # TMA variant is only selected when every source is device memory: restrict to cached ids
        cached = torch.from_numpy(np.unique(np.concatenate(lists)))
        q = cached[torch.randint(0, len(cached), (40000,), generator=torch.Generator().manual_seed(2))]
        qd = q.to(cuda)
    _lib.check(lib.dgs_extract_p2p(server._handle, ptr(host), D * 4, ptr(table), cap, 1, ptr(qd), len(q),
                                   ptr(out), algo, stream()))
    key, idx, dev = oracle.hashmap_build(lists, rank, n_unique)
    exp = oracle.extract_p2p(t2n(feat), [t2n(s) for s in shards], key, idx, dev, t2n(q))
    assert np.array_equal(t2n(out), exp)
    assert np.array_equal(t2n(out), t2n(feat)[t2n(q)])


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_route_ids_and_virtual_exchange(dgs, cuda, idt, P):
    """dgs_route_ids (request routing of the id-exchange extract, north star (4)): counts = requests
    per owner, send_idx = slot numbers grouped by owner, inv = the exact inverse of the grouping -
    and, with P virtual ranks on one GPU (shard r = rows r::P, the all-to-all done by slicing), the
    routed requests reproduce the closed-form feature rows in request order.  Request counts
    around the 2048-id CTA chunk."""
    N, D = 50000, 6
    feat = dgs_synth.make_features(N, D, device=cuda)
    shards = [feat[r::P].contiguous() for r in range(P)]
    g = torch.Generator().manual_seed(3 + P)
    for n in (0, 1, 31, 2048, 2049, 100003):
        q = torch.randint(0, N, (n,), generator=g).to(idt).to(cuda)
        send_idx, inv, counts = dgs.ops.route_ids(q, P)
        assert send_idx.dtype == idt and inv.dtype == idt and counts.dtype == torch.int64
        c = counts.cpu()
        assert torch.equal(c, torch.bincount((q % P).cpu().long(), minlength=P))
        assert torch.equal(torch.sort(inv).values.long(), torch.arange(n, device=cuda))
        assert torch.equal(send_idx[inv.long()], q // P)
        owner_of_pos = torch.repeat_interleave(torch.arange(P), c).to(cuda)
        assert torch.equal(owner_of_pos[inv.long()], (q % P).long())
        # the owners' side: rows of their own shard, concatenated in send order; then request order
        offs = [0] + torch.cumsum(c, 0).tolist()
        rows = torch.cat([dgs.ops._CAPI_cuda_index_select(shards[d], send_idx[offs[d]:offs[d + 1]])
                          for d in range(P)])
        out = dgs.ops._CAPI_cuda_index_select(rows, inv)
        assert torch.equal(out, dgs_synth.feature_rows(q.long(), D, torch.float32))
    with pytest.raises(RuntimeError):
        dgs.ops.route_ids(q, 0)


def test_feature_server_reference_test_input(dgs, cuda):
    """tests/test_feature_server.py:20-52 with one rank (cache [0, 3], misses from pinned host)."""
    feature = torch.arange(0, 100, 1).float().pin_memory().reshape(10, 10)
    fs = dgs.classes.P2PCacheFeatureServer(feature, torch.tensor([0, 3]).to(cuda), 0)
    assert torch.equal(fs._CAPI_get_cpu_feature(), feature)
    assert torch.equal(fs._CAPI_get_gpu_feature().cpu(), feature[[0, 3]])
    out = fs._CAPI_get_feature(torch.tensor([0, 3, 5, 7]).to(cuda))
    assert out.shape == (4, 10)
    assert torch.equal(out.cpu(), feature[[0, 3, 5, 7]])
    key, idx, dev = fs._CAPI_get_local_cache_hashmap_tensors()
    assert key.numel() == 8 and sorted(key[key >= 0].tolist()) == [0, 3]
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        fs._CAPI_get_feature(torch.tensor([0]))


@pytest.mark.parametrize("ratio", [0.01, 0.25, 1.0])
def test_feature_server_cache_ratio(dgs, cuda, ratio):
    N, D = 20000, 100
    feat = dgs_synth.feature_rows(torch.arange(N), D).pin_memory()
    g = torch.Generator().manual_seed(7)
    cache = torch.randperm(N, generator=g)[:max(1, int(N * ratio))]
    fs = dgs.classes.P2PCacheFeatureServer(feat, cache, 0)
    q = torch.randint(0, N, (50000,), generator=g)
    out = fs._CAPI_get_feature(q.to(cuda))
    assert torch.equal(out.cpu(), feat[q])
    key, idx, dev = oracle.hashmap_build([t2n(cache)], 0)
    exp = oracle.extract_p2p(t2n(feat), [t2n(feat[cache])], key, idx, dev, t2n(q))
    assert np.array_equal(t2n(out), exp)


# ------------------------------------------------------------------ sub-CSR extraction
def test_extract_indptr_edge_data_kat(dgs, cuda):
    """tests/test_extract.py:5-18"""
    indptr = torch.tensor([0, 4, 5, 5, 5, 5, 10, 10, 10, 10, 10, 10]).to(cuda)
    indices = torch.tensor([1, 2, 3, 4, 5, 6, 7, 8, 9, 10]).to(cuda)
    probs = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.1, 0.2, 0.3, 0.4, 0.5]).to(cuda)
    nids = torch.tensor([0, 1, 5]).to(cuda)
    sub = dgs.ops._Test_ExtractIndptr(nids, indptr)
    assert sub.tolist() == [0, 4, 5, 10]
    assert dgs.ops._Test_ExtractEdgeData(nids, indptr, sub, indices).tolist() == list(range(1, 11))
    assert torch.equal(dgs.ops._Test_ExtractEdgeData(nids, indptr, sub, probs), probs)


@pytest.mark.parametrize("src", ["cuda", "pinned"])
def test_extract_subcsr_random(dgs, cuda, src):
    indptr, indices, probs = small_graph(30000, 700000, seed=2, weights=True)
    g = torch.Generator().manual_seed(0)
    nids = torch.randperm(30000, generator=g)[:9000]
    put = (lambda t: t.to(cuda)) if src == "cuda" else (lambda t: t.pin_memory())
    ip, ix, pr = put(indptr), put(indices), put(probs)
    sub = dgs.ops._Test_ExtractIndptr(nids.to(cuda), ip)
    exp_sub = oracle.extract_indptr(t2n(nids), t2n(indptr))
    assert np.array_equal(t2n(sub), exp_sub)
    got_i = dgs.ops._Test_ExtractEdgeData(nids.to(cuda), ip, sub, ix)
    got_p = dgs.ops._Test_ExtractEdgeData(nids.to(cuda), ip, sub, pr)
    assert np.array_equal(t2n(got_i), oracle.extract_edge_data(t2n(nids), t2n(indptr), exp_sub, t2n(indices)))
    assert np.array_equal(t2n(got_p), oracle.extract_edge_data(t2n(nids), t2n(indptr), exp_sub, t2n(probs)))
    # int32 everything
    sub32 = dgs.ops._Test_ExtractIndptr(nids.int().to(cuda), put(indptr.int()))
    assert sub32.dtype == torch.int32 and np.array_equal(t2n(sub32), exp_sub)


# ------------------------------------------------------------------ full-neighbour sampling (bit-exact)
@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
@pytest.mark.parametrize("mode", ["minus1", "k>=maxdeg"])
def test_sample_all_neighbors_bit_exact(dgs, cuda, idt, mode):
    indptr, indices, _ = small_graph(5000, 120000, seed=1, dtype=idt)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    g = torch.Generator().manual_seed(0)
    seeds = torch.randint(0, 5000, (3000,), generator=g).to(idt)
    k = -1 if mode == "minus1" else maxdeg + 3
    ip = indptr.to(idt).to(cuda)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(seeds.to(cuda), ip, indices.to(cuda), k, False)
    er, ec = oracle.sample_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices))
    assert row.dtype == idt and col.dtype == idt
    assert np.array_equal(t2n(row), er) and np.array_equal(t2n(col), ec)
    # biased op, copy path: identical
    w = torch.rand(indices.numel()) + 0.1
    row2, col2 = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds.to(cuda), ip, indices.to(cuda), w.to(cuda), k, False)
    assert torch.equal(row2, row) and torch.equal(col2, col)


def test_sample_pinned_host_graph(dgs, cuda):
    """indptr / indices in pinned host memory, read through UVA (scripts/ncu_sampling.py:9-26)."""
    indptr, indices, _ = small_graph(2000, 30000, seed=4)
    seeds = torch.arange(0, 2000, 3)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(seeds.to(cuda), indptr.pin_memory(), indices.pin_memory(), -1, False)
    er, ec = oracle.sample_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices))
    assert np.array_equal(t2n(row), er) and np.array_equal(t2n(col), ec)


def test_sample_empty_and_isolated(dgs, cuda):
    indptr = torch.tensor([0, 0, 0, 3, 3]).to(cuda)
    indices = torch.tensor([1, 0, 3]).to(cuda)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(torch.empty(0, dtype=torch.int64, device=cuda), indptr, indices, 2, False)
    assert row.numel() == 0 and col.numel() == 0
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(torch.tensor([0, 1, 3]).to(cuda), indptr, indices, 2, False)
    assert row.numel() == 0
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(torch.tensor([0, 2, 3]).to(cuda), indptr, indices, 5, True)
    assert row.tolist() == [2] * 5 and set(col.tolist()) <= {0, 1, 3}


# ------------------------------------------------------------------ random sampling properties
def _check_sample(seeds, indptr, indices, row, col, k, replace):
    seeds, indptr, indices, row, col = map(t2n, (seeds, indptr, indices, row, col))
    o = 0
    for s in seeds:
        nb = indices[indptr[s]:indptr[s + 1]]
        d = len(nb)
        c = (k if d > 0 else 0) if replace else min(d, k)
        assert (row[o:o + c] == s).all()
        got = col[o:o + c]
        if not replace and d <= k:
            assert got.tolist() == nb.tolist()          # copy path, CSR order
        else:
            pool = {}
            for v in nb.tolist():
                pool[v] = pool.get(v, 0) + 1
            cnt = {}
            for v in got.tolist():
                assert v in pool                           # valid neighbour
                cnt[v] = cnt.get(v, 0) + 1
            if not replace:                                # no duplicate *edge* picked
                assert all(cnt[v] <= pool[v] for v in cnt)
        o += c
    assert o == len(row) == len(col)


@pytest.mark.parametrize("k", [1, 5, 15, 25, 32, 33, 70])
@pytest.mark.parametrize("replace", [False, True])
def test_uniform_sampling_properties(dgs, cuda, k, replace):
    indptr, indices, _ = small_graph(4000, 100000, seed=3)
    seeds = torch.randperm(4000, generator=torch.Generator().manual_seed(k))[:1500]
    args = (seeds.to(cuda), indptr.to(cuda), indices.to(cuda), k, replace)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(*args, rng_seed=123)
    _check_sample(seeds, indptr, indices, row, col, k, replace)
    row2, col2 = dgs.ops._CAPI_cuda_sample_neighbors(*args, rng_seed=123)
    assert torch.equal(col, col2)                          # same seed -> same sample
    row3, col3 = dgs.ops._CAPI_cuda_sample_neighbors(*args, rng_seed=124)
    assert not torch.equal(col, col3)


@pytest.mark.parametrize("k", [1, 10, 25, 32, 40])
@pytest.mark.parametrize("replace", [False, True])
def test_biased_sampling_properties(dgs, cuda, k, replace):
    indptr, indices, probs = small_graph(4000, 100000, seed=5, weights=True)
    seeds = torch.randperm(4000, generator=torch.Generator().manual_seed(k))[:1500]
    args = (seeds.to(cuda), indptr.to(cuda), indices.to(cuda), probs.to(cuda), k, replace)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(*args, rng_seed=9)
    _check_sample(seeds, indptr, indices, row, col, k, replace)
    row2, col2 = dgs.ops._CAPI_cuda_sample_neighbors_bias(*args, rng_seed=9)
    assert torch.equal(col, col2)


def _star_graph(deg, copies, cuda, weights=None):
    """`copies` seeds, each with the same `deg` distinct neighbours 0..deg-1 -> per-position counts."""
    indptr = torch.arange(0, (copies + 1) * deg, deg)
    indices = torch.arange(deg).repeat(copies)
    w = None if weights is None else weights.repeat(copies)
    return indptr.to(cuda), indices.to(cuda), None if w is None else w.to(cuda)


def _chi2(counts, expected):
    return float(((counts - expected) ** 2 / expected).sum())


@pytest.mark.parametrize("deg,k", [(10, 3), (40, 15), (200, 25), (33, 32)])
def test_uniform_without_replacement_is_uniform(dgs, cuda, deg, k):
    """Inclusion frequency of every neighbour position must be k/deg.  Chi-square over `deg` cells
    against the exact expectation; threshold = df + 6*sqrt(2*df) (about 6 sigma of the chi-square
    distribution, false-alarm rate < 1e-6)."""
    copies = 20000
    indptr, indices, _ = _star_graph(deg, copies, cuda)
    seeds = torch.arange(copies, device=cuda)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(seeds, indptr, indices, k, False, rng_seed=77)
    assert col.numel() == copies * k
    counts = torch.bincount(col, minlength=deg).double().cpu().numpy()
    exp = copies * k / deg
    # inclusion indicators are Bernoulli(p): variance exp*(1-p); normalise so that chi2 ~ chi2(df)
    p = k / deg
    chi2 = _chi2(counts, exp) / (1 - p) * (deg - 1) / deg
    df = deg - 1
    assert chi2 < df + 6 * np.sqrt(2 * df), (chi2, df)
    # and no duplicates inside a seed
    per_seed = col.reshape(copies, k).sort(dim=1).values
    assert (per_seed[:, 1:] != per_seed[:, :-1]).all()


@pytest.mark.parametrize("deg,k", [(10, 4), (100, 20)])
def test_uniform_with_replacement_is_uniform(dgs, cuda, deg, k):
    copies = 20000
    indptr, indices, _ = _star_graph(deg, copies, cuda)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors(torch.arange(copies, device=cuda), indptr, indices, k, True, rng_seed=5)
    counts = torch.bincount(col, minlength=deg).double().cpu().numpy()
    df = deg - 1
    assert _chi2(counts, copies * k / deg) < df + 6 * np.sqrt(2 * df)


@pytest.mark.parametrize("deg", [8, 50, 300])
def test_biased_k1_and_replace_follow_weights(dgs, cuda, deg):
    """k = 1 without replacement and any k with replacement must pick edge i with probability
    w_i / sum(w) exactly.  Chi-square against the weights, 6-sigma bound."""
    copies = 30000
    g = torch.Generator().manual_seed(deg)
    w = (torch.rand(deg, generator=g) * 3 + 0.05).float()
    indptr, indices, probs = _star_graph(deg, copies, cuda, w)
    seeds = torch.arange(copies, device=cuda)
    pw = (w / w.sum()).double().numpy()
    df = deg - 1
    bound = df + 6 * np.sqrt(2 * df)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, 1, False, rng_seed=3)
    counts = torch.bincount(col, minlength=deg).double().cpu().numpy()
    assert _chi2(counts, copies * pw) < bound
    for k in (1, 7):
        row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, k, True, rng_seed=4 + k)
        counts = torch.bincount(col, minlength=deg).double().cpu().numpy()
        assert _chi2(counts, copies * k * pw) < bound


def test_biased_without_replacement_matches_ares_reference(dgs, cuda):
    """k > 1: A-Res inclusion probabilities are not proportional to the weights (only the first draw
    is); compare with the CPU A-Res restatement of the oracle (two-sample chi-square, 6 sigma)."""
    deg, k, copies = 12, 4, 40000
    w = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.1, 0.2, 0.3, 0.4, 0.5, 2.0, 3.0])
    indptr, indices, probs = _star_graph(deg, copies, cuda, w)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(torch.arange(copies, device=cuda), indptr, indices, probs, k, False, rng_seed=1)
    a = torch.bincount(col, minlength=deg).double().cpu().numpy()
    r2, c2 = oracle.cpu_sample_neighbors(np.arange(copies), t2n(indptr), t2n(indices), t2n(probs), k, 99)
    b = np.bincount(c2, minlength=deg).astype(np.float64)
    chi2 = float(((a - b) ** 2 / (a + b)).sum())
    df = deg - 1
    assert chi2 < df + 6 * np.sqrt(2 * df), chi2
    per_seed = col.reshape(copies, k).sort(dim=1).values
    assert (per_seed[:, 1:] != per_seed[:, :-1]).all()


# ------------------------------------------------------------------ biased sampling: EXACT check
GOLDEN = 0x9E3779B97F4A7C15
M64 = (1 << 64) - 1


def _ares_graph(N=30000, seed=0):
    """Node n has deg_n DISTINCT neighbours (n * 1009 + t) % N, t < deg_n, so a returned id tells the
    position t that was picked.  Degrees cover the copy path (<= k), one-pass rows (<= 512 weights),
    multi-pass rows (513, 2 000) and hub rows shared by the whole CTA in the fused kernel (20 000)."""
    g = torch.Generator().manual_seed(seed)
    deg = torch.randint(5, 60, (N,), generator=g)
    special = {11: 40, 12: 513, 13: 2000, 14: 20000, 15: 512, 16: 1024, 17: 33, 18: 4097}
    for n, d in special.items():
        deg[n] = d
    indptr = torch.zeros(N + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(deg, 0)
    owner = torch.repeat_interleave(torch.arange(N), deg)
    t = torch.arange(int(indptr[-1])) - indptr[owner]
    indices = (owner * 1009 + t) % N
    w = (torch.randn(indices.numel(), generator=g).abs() + 1e-3).float()
    return indptr, indices, w, sorted(special)


def _expected_ares(dgs, w_dev, indptr, nid, k, key, item):
    b, e = int(indptr[nid]), int(indptr[nid + 1])
    keys = dgs.ops._Test_AresKeys(w_dev[b:e], key, item)
    # k largest keys, ties broken by the smaller position: a stable descending sort
    order = torch.sort(keys, descending=True, stable=True).indices
    return order[:k].cpu()


def _check_ares_hop(dgs, w_dev, indptr, N, seed_ids, picked_ids_per_seed, k, key, ordered):
    for i, (nid, got_ids) in enumerate(zip(seed_ids, picked_ids_per_seed)):
        deg = int(indptr[nid + 1] - indptr[nid])
        pos = (got_ids - nid * 1009) % N
        if deg <= k:
            assert pos.tolist() == list(range(deg))                        # copy path, CSR order
            continue
        exp = _expected_ares(dgs, w_dev, indptr, nid, k, key, i)
        if ordered:
            assert pos.tolist() == exp.tolist(), (nid, deg, k)
        else:
            assert sorted(pos.tolist()) == sorted(exp.tolist()), (nid, deg, k)


@pytest.mark.parametrize("k", [10, 25, 32, 40, 70])
def test_biased_without_replacement_exact_topk(dgs, cuda, k):
    """Weighted sampling without replacement IS the top-k of the A-Res keys log2(u_t) / w_t
    (rowwise_sampling_bias.cu:111-132), and the keys are a pure function of (launch key, seed index,
    edge position).  So every selection path can be checked EXACTLY against a top-k of the keys
    (dgs_debug_ares_keys): the per-hop op (warp per seed: threshold -> collect -> tighten passes for
    rows > 512 weights; replace-the-minimum reservoir for k > 32) and the fused batch kernel
    (CTA-shared hub rows, cross-warp merge).  k <= 32: picks come out in descending key order (set
    AND order compared); k > 32: reservoir order is unspecified (set compared)."""
    N = 30000
    indptr, indices, w, special = _ares_graph(N)
    ip, ix, wd = indptr.to(cuda), indices.to(cuda), w.to(cuda)
    g = torch.Generator().manual_seed(k)
    others = torch.randperm(N - 100, generator=g)[:700] + 100
    seeds = torch.cat([torch.tensor(special), others])
    seeds = seeds[torch.randperm(seeds.numel(), generator=g)]
    R = 0x1234567 + k
    # ---- per-hop op: launch key = rng_seed, item = seed position
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds.to(cuda), ip, ix, wd, k, False, rng_seed=R)
    cnt = torch.clamp(indptr[seeds + 1] - indptr[seeds], max=k)
    per_seed = torch.split(col.cpu(), cnt.tolist())
    assert torch.equal(row.cpu(), torch.repeat_interleave(seeds, cnt))
    _check_ares_hop(dgs, wd, indptr, N, seeds.tolist(), per_seed, k, R, ordered=k <= 32)
    # ---- fused batch kernel, two hops: hop l uses key rng_seed + GOLDEN * (l + 1), item = position
    # of the seed in that hop's seed list (hop 1's seeds = hop 0's frontier)
    smp = dgs.classes.CSRSampler(ip, ix, wd)
    k2 = 10 if k > 10 else 25
    out = smp._pipe.sample(seeds.to(cuda), [k2, k], False, R)     # fan-out walked from the back
    cur = seeds
    for l, ((s_, f_, r_, c_), kk) in enumerate(zip(out, (k, k2))):
        assert torch.equal(s_.cpu(), cur)
        f = f_.cpu()
        cnt = torch.clamp(indptr[cur + 1] - indptr[cur], max=kk)
        assert torch.equal(r_.cpu(), torch.repeat_interleave(torch.arange(cur.numel()), cnt))
        per_seed = torch.split(f[c_.cpu()], cnt.tolist())
        key = (R + GOLDEN * (l + 1)) & M64
        # hop 1 has ~10^4 seeds: check the special rows wherever they appear + a sample of the rest
        idx = list(range(cur.numel())) if l == 0 else \
            [i for i, n in enumerate(cur.tolist()) if n in special] + list(range(0, cur.numel(), 37))
        for i in idx:
            nid = int(cur[i])
            deg = int(indptr[nid + 1] - indptr[nid])
            pos = (per_seed[i] - nid * 1009) % N
            if deg <= kk:
                assert pos.tolist() == list(range(deg))
                continue
            exp = _expected_ares(dgs, wd, indptr, nid, kk, key, i)
            if kk <= 32:
                assert pos.tolist() == exp.tolist(), (l, nid, deg, kk)
            else:
                assert sorted(pos.tolist()) == sorted(exp.tolist()), (l, nid, deg, kk)
        cur = f


@pytest.mark.parametrize("k", [10, 25, 40])
def test_biased_zero_weight_edges_exact(dgs, cuda, k):
    """Zero-weight edges (key u^(1/0) = 0 in rowwise_sampling_bias.cu:112-113 - an edge
    of weight 0 is selectable only when fewer than k positive-weight edges exist).  Here the key of
    such an edge is -inf and ties go to the smaller position, so the sample is still EXACTLY the
    stable top-k of the keys for k <= 32: rows with fewer positive weights than k (every positive
    edge first, then the lowest zero-weight positions), with more than k (no zero-weight edge ever),
    all-zero rows, positives only at the very end of a multi-chunk row - through the one-pass,
    multi-pass and hub paths of the per-hop op and of the batch kernel; the k = 40 reservoir (beyond
    the reference's k <= 32) must pick exactly the top positive keys and fill up with zero-weight
    edges of its choice.  With replacement a zero-weight edge is never drawn."""
    N = 30000
    indptr, indices, w, special = _ares_graph(N, seed=3)
    keep = {11: [2, 17, 39], 12: [0, 100, 255, 256, 512], 13: list(range(0, 2000, 10)), 14: [5, 4095, 4096, 8191, 12000, 16384, 19999],
            15: list(range(500, 512)), 16: [], 17: [32], 18: list(range(4087, 4097))}
    for n, pos in keep.items():
        b, e = int(indptr[n]), int(indptr[n + 1])
        ww = torch.zeros(e - b)
        ww[pos] = w[b:e][pos]
        w[b:e] = ww
    # ordinary rows: a third of all weights are zero
    g = torch.Generator().manual_seed(40 + k)
    mask = torch.rand(w.numel(), generator=g) < 0.33
    lo = int(indptr[100])
    w[lo:][mask[lo:]] = 0.0
    ip, ix, wd = indptr.to(cuda), indices.to(cuda), w.to(cuda)
    others = torch.randperm(N - 100, generator=g)[:500] + 100
    seeds = torch.cat([torch.tensor(special), others])
    seeds = seeds[torch.randperm(seeds.numel(), generator=g)]
    R = 0x7654321 + k
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds.to(cuda), ip, ix, wd, k, False, rng_seed=R)
    cnt = torch.clamp(indptr[seeds + 1] - indptr[seeds], max=k)
    per_seed = torch.split(col.cpu(), cnt.tolist())
    assert torch.equal(row.cpu(), torch.repeat_interleave(seeds, cnt))

    def check(per_seed, key):
        if k <= 32:      # descending key order, ties (the -inf keys) by position: exact
            _check_ares_hop(dgs, wd, indptr, N, seeds.tolist(), per_seed, k, key, ordered=True)
        for i, (nid, got) in enumerate(zip(seeds.tolist(), per_seed)):
            b, e = int(indptr[nid]), int(indptr[nid + 1])
            if e - b <= k:
                assert ((got - nid * 1009) % N).tolist() == list(range(e - b))
                continue
            # the property the reference states, independent of tie-breaking (the k > 32 reservoir
            # may keep ANY zero-weight edges): positives first, and exactly the top positive keys
            pos = ((got - nid * 1009) % N).tolist()
            assert len(set(pos)) == k and all(0 <= t < e - b for t in pos)
            wrow = w[b:e]
            npos = int((wrow > 0).sum())
            picked = sorted(t for t in pos if wrow[t] > 0)
            assert len(picked) == min(npos, k), (nid, npos, len(picked))
            exp = _expected_ares(dgs, wd, indptr, nid, k, key, i).tolist()
            assert picked == sorted(t for t in exp if wrow[t] > 0), (nid, e - b, k)

    check(per_seed, R)
    # batch kernel (hub rows deferred to the chunked hub phase), first hop checked exactly
    smp = dgs.classes.CSRSampler(ip, ix, wd)
    out = smp._pipe.sample(seeds.to(cuda), [5, k], False, R)
    s_, f_, r_, c_ = out[0]
    per_seed = torch.split(f_.cpu()[c_.cpu()], cnt.tolist())
    check(per_seed, (R + GOLDEN) & M64)
    # with replacement: never a zero-weight edge (rows with at least one positive weight)
    has_pos = torch.tensor([bool((w[int(indptr[n]):int(indptr[n + 1])] > 0).any()) for n in seeds.tolist()])
    sd = seeds[has_pos]
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(sd.to(cuda), ip, ix, wd, k, True, rng_seed=R)
    rowc, colc = row.cpu(), col.cpu()
    t = (colc - rowc * 1009) % N
    assert bool((w[indptr[rowc] + t] > 0).all())


@pytest.mark.parametrize("case", ["many_hubs", "huge_row"])
def test_biased_hub_phase_fallbacks_exact(dgs, cuda, case):
    """The chunked hub phase of the batch kernel has static limits (1024 deferred rows, 32 chunks of
    4096 weights per row); beyond them the rows are scanned one CTA per row.  Both fallbacks must
    return exactly the top-k of the A-Res keys: (a) 2 000 seeds that are all > 512-weight rows,
    (b) one row of 140 000 weights (35 chunks) next to ordinary ones."""
    g = torch.Generator().manual_seed(5)
    if case == "many_hubs":
        N = 3000
        deg = torch.full((N,), 600, dtype=torch.int64)
        seeds = torch.randperm(N, generator=g)[:2000]
    else:
        N = 150000
        deg = torch.randint(3, 9, (N,), generator=g)
        deg[7] = 140000
        deg[9] = 70000
        seeds = torch.cat([torch.tensor([7, 9]), torch.randperm(N - 10, generator=g)[:300] + 10])
        seeds = seeds[torch.randperm(seeds.numel(), generator=g)]
    indptr = torch.zeros(N + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(deg, 0)
    owner = torch.repeat_interleave(torch.arange(N), deg)
    t = torch.arange(int(indptr[-1])) - indptr[owner]
    indices = (owner * 1009 + t) % N
    w = (torch.randn(indices.numel(), generator=g).abs() + 1e-3).float()
    ip, ix, wd = indptr.to(cuda), indices.to(cuda), w.to(cuda)
    smp = dgs.classes.CSRSampler(ip, ix, wd)
    k, R = 25, 4242
    (s_, f_, r_, c_), = smp._pipe.sample(seeds.to(cuda), [k], False, R)
    cnt = torch.clamp(indptr[seeds + 1] - indptr[seeds], max=k)
    assert torch.equal(r_.cpu(), torch.repeat_interleave(torch.arange(seeds.numel()), cnt))
    per_seed = torch.split(f_.cpu()[c_.cpu()], cnt.tolist())
    key = (R + GOLDEN) & M64
    check = range(seeds.numel()) if case == "huge_row" else range(0, seeds.numel(), 41)
    for i in check:
        nid = int(seeds[i])
        d = int(indptr[nid + 1] - indptr[nid])
        pos = (per_seed[i] - nid * 1009) % N
        if d <= k:
            assert pos.tolist() == list(range(d))
        else:
            assert pos.tolist() == _expected_ares(dgs, wd, indptr, nid, k, key, i).tolist(), (case, nid, d)


def test_biased_k1_and_replace_follow_weights_long_row(dgs, cuda):
    """deg 2 000 (four 512-weight passes per row): k = 1 w/o replacement and k = 7 with replacement
    must follow w_i / sum(w); chi-square over 2 000 cells, 6-sigma bound.  Through the per-hop op and
    through the fused batch kernel."""
    deg, copies = 2000, 30000
    g = torch.Generator().manual_seed(deg)
    w = (torch.rand(deg, generator=g) * 3 + 0.05).float()
    indptr, indices, probs = _star_graph(deg, copies, cuda, w)
    seeds = torch.arange(copies, device=cuda)
    pw = (w / w.sum()).double().numpy()
    df = deg - 1
    bound = df + 6 * np.sqrt(2 * df)
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, 1, False, rng_seed=3)
    counts = torch.bincount(col, minlength=deg).double().cpu().numpy()
    assert _chi2(counts, copies * pw) < bound
    row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, indptr, indices, probs, 7, True, rng_seed=11)
    counts = torch.bincount(col, minlength=deg).double().cpu().numpy()
    assert _chi2(counts, copies * 7 * pw) < bound
    # fused kernel (num_nodes = copies; neighbour ids < deg <= copies)
    smp = dgs.classes.CSRSampler(indptr, indices, probs)
    for k, rep, rs in ((1, False, 5), (7, True, 6)):
        (s_, f_, r_, c_), = smp._CAPI_sample_node_classifiction(seeds, [k], rep, rng_seed=rs)
        counts = torch.bincount(f_[c_], minlength=deg).double().cpu().numpy()[:deg]
        assert _chi2(counts, copies * k * pw) < bound


# ------------------------------------------------------------------ relabel (bit-exact)
@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
def test_relabel_bit_exact(dgs, cuda, idt):
    g = torch.Generator().manual_seed(0)
    for n_seed, n_col, space in [(3, 9, 11), (1024, 5000, 3000), (5000, 200000, 50000), (10, 300000, 2449029)]:
        seeds = torch.randperm(space, generator=g)[:n_seed].to(idt)
        col = torch.randint(0, space, (n_col,), generator=g).to(idt)
        row = seeds[torch.randint(0, n_seed, (n_col,), generator=g)]
        uniq, (rrow, rcol) = dgs.ops._CAPI_cuda_sampled_tensor_relabel([seeds.to(cuda), col.to(cuda)], [row.to(cuda), col.to(cuda)])
        eu, (er, ec) = oracle.relabel([t2n(seeds), t2n(col)], [t2n(row), t2n(col)])
        assert uniq.dtype == idt
        assert np.array_equal(t2n(uniq), eu) and np.array_equal(t2n(rrow), er) and np.array_equal(t2n(rcol), ec)


def test_relabel_absent_keys_and_many_parts(dgs, cuda):
    a = torch.tensor([5, 7, 5, 9]).to(cuda)
    parts = [a, torch.tensor([1, 7]).to(cuda), torch.tensor([2]).to(cuda), torch.tensor([9, 3]).to(cuda), torch.tensor([4, 4]).to(cuda)]
    q = torch.tensor([[9, 100], [4, 5]]).to(cuda)
    uniq, (rq,) = dgs.ops._CAPI_cuda_sampled_tensor_relabel(parts, [q])
    assert uniq.tolist() == [5, 7, 9, 1, 2, 3, 4]
    assert rq.tolist() == [[2, -1], [6, 0]]
    # the reference test graph
    uniq, (rr, rc) = dgs.ops._CAPI_cuda_sampled_tensor_relabel(
        [torch.tensor([0, 3, 5]).to(cuda), torch.tensor([1, 2, 3, 4, 6, 7, 8, 9, 10]).to(cuda)],
        [torch.tensor([0, 0, 0, 0, 5, 5, 5, 5, 5]).to(cuda), torch.tensor([1, 2, 3, 4, 6, 7, 8, 9, 10]).to(cuda)])
    assert uniq.tolist() == [0, 3, 5, 1, 2, 4, 6, 7, 8, 9, 10]
    assert rr.tolist() == [0, 0, 0, 0, 2, 2, 2, 2, 2] and rc.tolist() == [3, 4, 1, 5, 6, 7, 8, 9, 10]


# ------------------------------------------------------------------ sampler class
def _make_sampler(dgs, cuda, indptr, indices, probs, cache):
    ip, ix = indptr.pin_memory(), indices.pin_memory()
    pr = probs.pin_memory() if probs is not None else torch.Tensor()
    return dgs.classes.P2PCacheSampler(ip, ix, pr, cache, 0), ip, ix


def test_sampler_build_reference_test_input(dgs, cuda):
    """tests/test_build_sampler.py:17-44 from rank 1's cache set, one rank."""
    indptr = torch.tensor([0, 4, 5, 5, 5, 5, 10, 10, 10, 10, 10, 10])
    indices = torch.tensor([1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
    probs = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.1, 0.2, 0.3, 0.4, 0.5])
    s, ip, ix = _make_sampler(dgs, cuda, indptr, indices, probs, torch.tensor([3, 5]))
    a, b, c = s._CAPI_get_cpu_structure_tensors()
    assert a is ip and b is ix and torch.equal(c, probs)
    li, lx, lp = s._CAPI_get_local_cache_structure_tensors()
    assert li.tolist() == [0, 0, 5] and lx.tolist() == [6, 7, 8, 9, 10]
    assert torch.equal(lp.cpu(), probs[5:])
    key, idx, dev = s._CAPI_get_local_cache_hashmap_tensors()
    assert key.numel() == 8                      # 2 * _UpPower(2)
    ent = sorted(zip(key[key >= 0].tolist(), idx[key >= 0].tolist(), dev[key >= 0].tolist()))
    assert ent == [(3, 0, 0), (5, 1, 0)]


def test_sampler_uniform_reference_test_input(dgs, cuda):
    """tests/test_sampler_uniform.py:14-38, one rank."""
    indptr = torch.tensor([0, 4, 5, 5, 5, 5, 10, 10, 10, 10, 10, 10])
    indices = torch.tensor([1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
    s, _, _ = _make_sampler(dgs, cuda, indptr, indices, None, torch.tensor([0, 3]))
    out = s._CAPI_sample_node_classifiction(torch.tensor([0, 3, 5]).to(cuda), [2, 2], False)
    assert len(out) == 2
    seeds, frontier, row, col = out[0]
    assert seeds.tolist() == [0, 3, 5] and row.tolist() == [0, 0, 2, 2]
    assert frontier[:3].tolist() == [0, 3, 5]
    nb = frontier[col].tolist()
    assert set(nb[:2]) <= {1, 2, 3, 4} and len(set(nb[:2])) == 2
    assert set(nb[2:]) <= {6, 7, 8, 9, 10} and len(set(nb[2:])) == 2
    seeds2, frontier2, row2, col2 = out[1]
    assert torch.equal(seeds2, frontier)
    # second hop: node 0 and 5 give 2 each, node 1 (if reached) gives [5], everything else nothing
    exp = 4 + (1 if 1 in frontier.tolist() else 0)
    assert row2.numel() == exp


@pytest.mark.parametrize("bias", [False, True])
@pytest.mark.parametrize("cache_frac", [0.02, 0.5, 1.0])
def test_sampler_blocks_copy_path_bit_exact(dgs, cuda, bias, cache_frac):
    """Multi-hop blocks with every fan-out >= max degree (the reference's way to ask for all
    neighbours) through the fused whole-batch path, and with fan-out -1 through the per-hop path:
    both must equal the oracle's layer loop bit for bit (sampler.cc:14-62)."""
    N = 1500
    indptr, indices, probs = dgs_synth.make_csr(N, 9000, seed=11, weights=bias, classes=5)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    assert maxdeg < 400
    g = torch.Generator().manual_seed(1)
    cache = torch.randperm(N, generator=g)[:max(1, int(N * cache_frac))]
    s, _, _ = _make_sampler(dgs, cuda, indptr, indices, probs, cache)
    seeds = torch.randperm(N, generator=g)[:40]
    exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices), 2)
    for fan in ([maxdeg, maxdeg], [-1, -1]):
        out = s._CAPI_sample_node_classifiction(seeds.to(cuda), fan, False)
        assert len(out) == 2
        for (s_, f_, r_, c_), (es, ef, er, ec) in zip(out, exp):
            assert np.array_equal(t2n(s_), es) and np.array_equal(t2n(f_), ef)
            assert np.array_equal(t2n(r_), er) and np.array_equal(t2n(c_), ec)
    # repeated calls reuse the persistent relabel table: must stay clean
    out2 = s._CAPI_sample_node_classifiction(seeds.to(cuda), [maxdeg, maxdeg], False)
    assert np.array_equal(t2n(out2[1][1]), exp[1][1])


@pytest.mark.parametrize("bias", [False, True])
def test_sampler_blocks_random_properties(dgs, cuda, bias):
    N = 20000
    indptr, indices, probs = small_graph(N, 500000, seed=12, weights=bias)
    g = torch.Generator().manual_seed(2)
    cache = torch.randperm(N, generator=g)[:5000]
    s, _, _ = _make_sampler(dgs, cuda, indptr, indices, probs, cache)
    seeds = torch.randperm(N, generator=g)[:1024]
    fan = [15, 10, 5]
    out = s._CAPI_sample_node_classifiction(seeds.to(cuda), fan, False, rng_seed=5)
    assert len(out) == 3
    cur = seeds
    for li, (s_, f_, r_, c_) in enumerate(out):
        k = fan[len(fan) - 1 - li]
        assert torch.equal(s_.cpu(), cur)
        f = f_.cpu()
        assert torch.equal(f[:len(cur)], cur)             # seeds come first (they are unique)
        assert len(torch.unique(f)) == len(f)
        row_g, col_g = cur[r_.cpu()], f[c_.cpu()]
        # relabelled COO maps back to a valid sample of the hop
        _check_sample(cur, indptr, indices, row_g, col_g, k, False)
        # frontier = first-occurrence unique of cat(seeds, cols)
        eu, _ = oracle.relabel([t2n(cur), t2n(col_g)], [])
        assert np.array_equal(t2n(f), eu)
        cur = f
    again = s._CAPI_sample_node_classifiction(seeds.to(cuda), fan, False, rng_seed=5)
    assert all(torch.equal(a[3], b[3]) and torch.equal(a[1], b[1]) for a, b in zip(out, again))


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
def test_fused_blocks_edge_cases(dgs, cuda, idt):
    """Fused whole-batch path: duplicate seeds, fan-out 0, 1 and 4 hops, int32 ids, replace=True,
    tiny and non-multiple-of-tile seed counts - against the oracle's layer loop (copy path) or the
    hop invariants (random path)."""
    N = 2500
    indptr, indices, _ = dgs_synth.make_csr(N, 30000, seed=17, classes=6, id_dtype=idt)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    smp = dgs.classes.CSRSampler(indptr.to(idt).to(cuda), indices.to(cuda))
    g = torch.Generator().manual_seed(5)
    for n_seeds in (1, 7, 129, 1000):
        seeds = torch.randint(0, N, (n_seeds,), generator=g).to(idt)      # duplicates allowed
        for L in (1, 4) if n_seeds <= 7 else (2,):
            exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices), L)
            out = smp._CAPI_sample_node_classifiction(seeds.to(cuda), [maxdeg] * L, False)
            assert len(out) == L
            for a, e in zip(out, exp):
                assert a[0].dtype == idt
                for x, z in zip(a, e):
                    assert np.array_equal(t2n(x), z)
    seeds = torch.randperm(N, generator=g)[:300].to(idt)
    # fan-out 0: no edges, frontier = seeds
    out = smp._CAPI_sample_node_classifiction(seeds.to(cuda), [0], False)
    assert out[0][2].numel() == 0 and torch.equal(out[0][1].cpu(), seeds)
    out = smp._CAPI_sample_node_classifiction(seeds.to(cuda), [3, 0], False)   # hop 1 k=0, hop 2 k=3
    assert out[0][2].numel() == 0 and torch.equal(out[1][0].cpu(), seeds)
    # with replacement through the fused path
    out = smp._CAPI_sample_node_classifiction(seeds.to(cuda), [4, 3], True, rng_seed=2)
    cur = seeds
    for (s_, f_, r_, c_), k in zip(out, (3, 4)):
        f = f_.cpu()
        _check_sample(cur, indptr, indices, cur[r_.cpu().long()], f[c_.cpu().long()], k, True)
        eu, _ = oracle.relabel([t2n(cur), t2n(f[c_.cpu().long()])], [])
        assert np.array_equal(t2n(f), eu)
        cur = f


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
def test_fused_blocks_small_first_hop(dgs, cuda, idt):
    """Small first hops (<= 1024 seeds, fan-out <= 16, two to four hops) through the whole-batch
    kernel.  Copy path against the oracle's layer loop on a graph whose rows are all <= 12 long
    (isolated nodes, duplicate seeds, seed counts around the 8-seed pick tiles, the 128-seed rank
    tiles and 1024), random path against the per-hop ops with the same Philox counters.  (Written
    for the one-CTA first hop that round 2 measured and dropped - DESIGN.md section 5 - and kept:
    it pins the tile phases on the same boundaries.)"""
    N = 3000
    g = torch.Generator().manual_seed(23)
    deg = torch.randint(0, 13, (N,), generator=g)
    deg[::7] = 0
    indptr = torch.zeros(N + 1, dtype=torch.int64)
    torch.cumsum(deg, 0, out=indptr[1:])
    indices = torch.randint(0, N, (int(indptr[-1]),), generator=g).to(idt)
    smp = dgs.classes.CSRSampler(indptr.to(idt).to(cuda), indices.to(cuda))
    for n_seeds in (1, 3, 4, 5, 255, 257, 1023, 1024, 1025):
        seeds = torch.randint(0, N, (n_seeds,), generator=g).to(idt)           # duplicates allowed
        for fan in ([12, 12], [16, 12, 12]):
            exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices), len(fan))
            out = smp._CAPI_sample_node_classifiction(seeds.to(cuda), fan, False)
            for a, e in zip(out, exp):
                for x, z in zip(a, e):
                    assert x.dtype == idt and np.array_equal(t2n(x), z)
    # random path: degrees on both sides of the fan-out, hubs, duplicates
    indptr, indices, _ = dgs_synth.make_csr(9000, 400000, seed=29, id_dtype=idt)
    smp = dgs.classes.CSRSampler(indptr.to(idt).to(cuda), indices.to(cuda))
    for it, n_seeds in enumerate((2, 130, 777, 1024)):
        seeds = torch.randint(0, 9000, (n_seeds,), generator=g).to(idt).to(cuda)
        for fan in ([10, 5], [15, 10, 5], [4, 16], [3, 2, 1, 7]):
            a = smp._pipe.sample(seeds, fan, False, 900 + it)
            b = smp._pipe._sample_per_hop(seeds, fan, False, 900 + it)
            assert len(a) == len(b) == len(fan)
            assert all(torch.equal(u, v) for x, y in zip(a, b) for u, v in zip(x, y)), (n_seeds, fan)


def test_fused_blocks_cooperative_equals_multi_kernel(dgs, cuda, monkeypatch):
    """The cooperative single-launch kernel and the 3-kernels-per-hop path share their phase code
    and RNG counters: identical outputs for the same seed (DGS_BLOCKS_MODE is read once per
    process, so the comparison runs in a child process)."""
    import subprocess, sys, os, json
    code = r"""
import sys, os, hashlib
sys.path.insert(0, os.path.join(os.getcwd(), 'dist-gnn_b200'))
import torch, dgs, dgs_synth
ip, ix, pr = dgs_synth.make_csr(20000, 500000, seed=12, weights=True)
seeds = torch.randperm(20000, generator=torch.Generator().manual_seed(2))[:1024].cuda()
h = hashlib.sha256()
for probs in (None, pr.cuda()):
    s = dgs.classes.CSRSampler(ip.cuda(), ix.cuda(), probs)
    for rep in (False, True):
        for _ in range(2):
            for blk in s._CAPI_sample_node_classifiction(seeds, [15, 10, 5], rep, rng_seed=11):
                for t in blk:
                    h.update(t.cpu().numpy().tobytes())
print(h.hexdigest())
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    # coop: one cooperative launch per batch; split: one kernel per phase (what a big batch or B >= 2
    # gets; forced here for B = 1); multi: the round-1 kernels, 3 launches per hop, 8-byte tables + wipes
    for env_extra in ({"DGS_MB_SPLIT": "0"}, {"DGS_MB_SPLIT": "1"}, {"DGS_BLOCKS_MODE": "multi"}):
        env = dict(os.environ, **env_extra)
        p = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        digests.append(p.stdout.strip().splitlines()[-1])
    assert digests[0] == digests[1] == digests[2]


@pytest.mark.parametrize("replace", [False, True])
@pytest.mark.parametrize("bias", [False, True])
def test_fused_blocks_equal_per_hop_ops(dgs, cuda, bias, replace):
    """The fused whole-batch kernel (tile phases: thread-per-seed Floyd in registers, CTA-shared hub
    rows for weighted sampling) and the per-hop ops (plan + warp-per-seed pick + relabel op) draw
    from the same Philox counters: with the same seed every hop must be identical, tensor by
    tensor - on a graph whose degrees reach past the 512-weight hub threshold and the fan-outs
    cover k <= 16, 16 < k <= 32 and k > 32."""
    N = 12000
    indptr, indices, probs = dgs_synth.make_csr(N, 1500000, seed=61, weights=bias)
    deg = indptr[1:] - indptr[:-1]
    assert int(deg.max()) > 4096 and int((deg > 512).sum()) > 20
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda), probs.to(cuda) if bias else None)
    hubs = torch.argsort(deg, descending=True)[:40]
    seeds = torch.unique(torch.cat([hubs, torch.randperm(N, generator=torch.Generator().manual_seed(3))[:400]]))
    seeds = seeds[torch.randperm(seeds.numel(), generator=torch.Generator().manual_seed(4))].to(cuda)
    for fan in ([12, 5], [25, 10], [40]):
        a = smp._pipe.sample(seeds, fan, replace, 777)
        b = smp._pipe._sample_per_hop(seeds, fan, replace, 777)
        assert len(a) == len(b) == len(fan)
        for x, y in zip(a, b):
            for u, v in zip(x, y):
                assert torch.equal(u, v)


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
@pytest.mark.parametrize("bias", [False, True])
def test_fused_blocks_direct_equals_hashed_tables(dgs, cuda, idt, bias):
    """The relabel tables of the fused path are direct-addressed when the node count is known
    (dgs_graph_t.num_nodes) and hashed otherwise: both must give the same blocks, batch after batch
    (the tables are wiped and reused), including duplicate seeds and a node count that is odd."""
    N = 30001
    indptr, indices, probs = dgs_synth.make_csr(N, 600000, seed=51, weights=bias, classes=6, id_dtype=idt)
    args = (indptr.to(idt).to(cuda), indices.to(cuda), probs.to(cuda) if bias else None)
    direct, hashed = dgs.classes.CSRSampler(*args), dgs.classes.CSRSampler(*args)
    assert direct._pipe._graph.num_nodes == N
    hashed._pipe._graph.num_nodes = 0
    g = torch.Generator().manual_seed(9)
    for it in range(6):
        seeds = torch.randint(0, N, (777,), generator=g).to(idt).to(cuda)
        seeds[-1] = N - 1
        for fan, rep in (([15, 10, 5], False), ([4, 3], True)):
            a = direct._CAPI_sample_node_classifiction(seeds, fan, rep, rng_seed=it)
            b = hashed._CAPI_sample_node_classifiction(seeds, fan, rep, rng_seed=it)
            assert all(torch.equal(x, y) for u, v in zip(a, b) for x, y in zip(u, v))


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
@pytest.mark.parametrize("bias,replace", [(False, False), (True, False), (False, True), (True, True)])
def test_sample_many_equals_single_calls(dgs, cuda, idt, bias, replace):
    """B mini-batches in ONE cooperative launch (dgs_sample_blocks_multi) must be bit-identical to B
    single-batch calls with the same RNG seeds - every hop, every tensor - for B = 1, 3, 8, on a graph
    with hub rows, duplicate seeds in hop 0, batch sizes that are not tile multiples, and repeated
    launches on the same workspace (its relabel tables must come back clean)."""
    N = 20000
    indptr, indices, probs = dgs_synth.make_csr(N, 900000, seed=71, weights=bias, id_dtype=idt)
    smp = dgs.classes.CSRSampler(indptr.to(idt).to(cuda), indices.to(cuda), probs.to(cuda) if bias else None)
    g = torch.Generator().manual_seed(7)
    fan = [15, 10, 5]
    for B, S in ((1, 1024), (3, 333), (8, 512), (8, 512)):
        seeds = torch.randint(0, N, (B, S), generator=g).to(idt).to(cuda)     # duplicates inside a batch
        rng = [1000 + 17 * b for b in range(B)]
        many = smp.sample_many(seeds, fan, replace, rng)
        assert len(many) == B
        for b in range(B):
            one = smp._CAPI_sample_node_classifiction(seeds[b].clone(), fan, replace, rng_seed=rng[b])
            assert len(one) == len(many[b]) == 3
            for x, y in zip(many[b], one):
                for u, v in zip(x, y):
                    assert torch.equal(u, v)
    # a second fan-out shape (two hops, k > 16: Floyd in shared memory) and the deterministic path
    maxdeg_fan = [40, 20]
    seeds = torch.randperm(N, generator=g)[:4 * 200].reshape(4, 200).to(idt).to(cuda)
    many = smp.sample_many(seeds, maxdeg_fan, replace, [5, 6, 7, 8])
    for b in range(4):
        one = smp._CAPI_sample_node_classifiction(seeds[b].clone(), maxdeg_fan, replace, rng_seed=5 + b)
        assert all(torch.equal(u, v) for x, y in zip(many[b], one) for u, v in zip(x, y))


def test_relabel_table_epoch_tags_wrap(dgs, cuda):
    """The direct relabel tables are never wiped: entries carry an 8-bit epoch tag that counts down
    per hop and the host clears the table when the tags run out (every 255 hops).  300 consecutive
    calls on ONE workspace (3 hops each -> three wraps), alternating two seed sets so that stale
    entries of the previous call are always in the table: every call must reproduce its reference."""
    N = 6000
    indptr, indices, _ = dgs_synth.make_csr(N, 150000, seed=81)
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda))
    g = torch.Generator().manual_seed(1)
    sets = [torch.randint(0, N, (256,), generator=g).to(cuda) for _ in range(2)]
    fan = [6, 5, 4]
    ref = [smp._pipe._sample_per_hop(sd, fan, False, 40 + i) for i, sd in enumerate(sets)]
    for it in range(300):
        i = it & 1
        out = smp._CAPI_sample_node_classifiction(sets[i], fan, False, rng_seed=40 + i)
        assert all(torch.equal(u, v) for x, y in zip(out, ref[i]) for u, v in zip(x, y)), it
    many = torch.stack(sets)
    for it in range(100):
        out = smp.sample_many(many, fan, False, [40, 41])
        for i in range(2):
            assert all(torch.equal(u, v) for x, y in zip(out[i], ref[i]) for u, v in zip(x, y)), it


def test_sample_many_copy_path_matches_oracle(dgs, cuda):
    """All-neighbour fan-outs through the multi-batch kernel against the oracle's layer loop."""
    N = 1500
    indptr, indices, _ = dgs_synth.make_csr(N, 9000, seed=11, classes=5)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda))
    g = torch.Generator().manual_seed(3)
    seeds = torch.randint(0, N, (5, 40), generator=g)
    many = smp.sample_many(seeds.to(cuda), [maxdeg, maxdeg], False, [1, 2, 3, 4, 5])
    for b in range(5):
        exp = oracle.sample_blocks_all_neighbors(t2n(seeds[b]), t2n(indptr), t2n(indices), 2)
        for a, e in zip(many[b], exp):
            for x, z in zip(a, e):
                assert np.array_equal(t2n(x), z)


def test_sample_blocks_workspace_is_checked(dgs, cuda):
    """A workspace remembers what it was initialised for: fewer seeds than at init are fine with
    direct tables (same result as a fresh call), anything else is refused instead of silently
    reading the relabel tables at the wrong offsets; so is a pointer that was never initialised."""
    import ctypes as C
    from dgs import _lib
    from dgs._util import stream
    l = _lib.lib()
    N = 8000
    indptr, indices, _ = dgs_synth.make_csr(N, 200000, seed=5)
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda))
    pipe = smp._pipe
    fan = [10, 5]
    big = pipe._plan(512, fan)
    seeds = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:300].to(cuda)

    def call(pl, ws, S, fo):
        arena = torch.empty(pl["total"] + pl["count_slots"], dtype=torch.int64, device=cuda)
        base = arena.data_ptr()
        for li, (of, orow, ocol) in enumerate(pl["offs"]):
            pl["a_fr"][li], pl["a_row"][li], pl["a_col"][li] = base + of * 8, base + orow * 8, base + ocol * 8
        rc = l.dgs_sample_blocks(C.byref(pipe._graph), seeds.data_ptr(), S, len(fo), _lib.i64_array(fo), 0,
                                 C.c_uint64(9), pl["a_fr"], pl["a_row"], pl["a_col"], pl["cap_edges"],
                                 pl["cap_front"], base + pl["total"] * 8, ws.data_ptr(), ws.numel(), 0,
                                 pl["counts_ptr"], stream())
        return rc, arena
    rc, arena = call(big, big["ws"], 300, fan)
    assert rc == 0
    ref = smp._CAPI_sample_node_classifiction(seeds, fan, False, rng_seed=9)
    counts = big["counts_np"].tolist()
    assert counts == [ref[0][2].numel(), ref[0][1].numel(), ref[1][2].numel(), ref[1][1].numel()]
    of, orow, ocol = big["offs"][1]
    assert torch.equal(arena[of:of + counts[3]], ref[1][1]) and torch.equal(arena[ocol:ocol + counts[2]], ref[1][3])
    rc, _ = call(big, big["ws"], 300, [10, 6])                    # another fan-out
    assert rc != 0 and b"another configuration" in l.dgs_last_error()
    bigger = pipe._plan(600, fan)
    rc, _ = call(bigger, big["ws"], 600, fan)                     # more seeds than at init
    assert rc != 0 and b"another configuration" in l.dgs_last_error()
    stray = torch.empty(big["ws"].numel(), dtype=torch.uint8, device=cuda)
    rc, _ = call(big, stray, 300, fan)
    assert rc != 0 and b"never initialised" in l.dgs_last_error()


def test_batch_loader_load_many(dgs, cuda):
    """BatchLoader.load_many (B batches, one sampling launch) == B x load() with the same seeds."""
    N, D = 9000, 100
    indptr, indices, _ = dgs_synth.make_csr(N, 200000, seed=45, classes=8)
    feat = dgs_synth.feature_rows(torch.arange(N), D)
    labels = (torch.arange(N) % 47).to(cuda)
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda))
    for src in (feat.to(cuda), dgs.classes.P2PCacheFeatureServer(feat.pin_memory(), torch.arange(N), 0)):
        loader = dgs.classes.BatchLoader(smp, src, labels)
        g = torch.Generator().manual_seed(8)
        for it in range(3):
            seeds = torch.randperm(N, generator=g)[:6 * 128].reshape(6, 128)
            rng = [it * 10 + b for b in range(6)]
            res = loader.load_many(seeds.pin_memory() if it % 2 else seeds.to(cuda), [10, 5], False, rng)
            assert len(res) == 6
            for b, (blocks, x, y) in enumerate(res):
                rb, rx, ry = loader.load(seeds[b].to(cuda), [10, 5], False, rng_seed=rng[b])
                assert all(torch.equal(u, v) for p_, q_ in zip(blocks, rb) for u, v in zip(p_, q_))
                assert torch.equal(x, rx) and torch.equal(y, ry)
                assert torch.equal(x.cpu(), feat[blocks[-1][1].cpu()])


@pytest.mark.parametrize("bias", [False, True])
def test_huge_num_picks_uses_output_scratch(dgs, cuda, bias):
    """num_picks far beyond what shared memory holds (the reference asserts num_picks <= 32 for the
    biased kernel, rowwise_sampling_bias.cu:73): rows above k are still sampled without
    duplicates, rows below are copied; and the class API falls back to exactly-sized hops."""
    N = 3000
    indptr, indices, probs = dgs_synth.make_csr(N, 400000, seed=13, weights=True, classes=9)
    deg = indptr[1:] - indptr[:-1]
    k = 4000   # 8 warps * k * 4 (or 8) bytes > the 96 KiB shared-memory budget of the pick kernel
    assert int((deg > k).sum()) >= 3
    seeds = torch.argsort(deg, descending=True)[:64]
    # distinct neighbour ids per row so "no duplicates" is checkable on ids
    for s_ in seeds.tolist():
        b, e = int(indptr[s_]), int(indptr[s_ + 1])
        indices[b:e] = torch.arange(e - b) % N if e - b <= N else indices[b:e]
    args = [seeds.to(cuda), indptr.to(cuda), indices.to(cuda)]
    if bias:
        row, col = dgs.ops._CAPI_cuda_sample_neighbors_bias(*args, probs.to(cuda), k, False, rng_seed=3)
    else:
        row, col = dgs.ops._CAPI_cuda_sample_neighbors(*args, k, False, rng_seed=3)
    _check_sample(seeds, indptr, indices, row, col, k, False)


@pytest.mark.parametrize("bias", [False, True])
def test_modulo_sharded_fast_path_single_rank(dgs, cuda, bias):
    """cache_nids == arange(N): every node cached in id order -> the owner is arithmetic and no
    location table is built; results must equal the hash path and the oracle, through both the
    CPU-tensor constructors and the from_device_shard(s) extensions."""
    N, D = 4000, 100
    indptr, indices, probs = dgs_synth.make_csr(N, 90000, seed=41, weights=bias, classes=6)
    feat = dgs_synth.feature_rows(torch.arange(N), D)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    ip, ix = indptr.pin_memory(), indices.pin_memory()
    pr = probs.pin_memory() if bias else torch.Tensor()
    allnodes = torch.arange(N)
    s1 = dgs.classes.P2PCacheSampler(ip, ix, pr, allnodes, 0)
    assert s1._mod_world == 1 and s1._table is None
    nids, sp, si, spr = dgs_synth.make_shard(N, 90000, 0, 1, seed=41, device=cuda, weights=bias, classes=6)
    s2 = dgs.classes.P2PCacheSampler.from_device_shards(sp, si, spr, nids, N, 0)
    assert s2._mod_world == 1
    seeds = torch.randperm(N, generator=torch.Generator().manual_seed(3))[:200].to(cuda)
    exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices), 2)
    for smp in (s1, s2):
        for fan in ([maxdeg, maxdeg], [-1, -1]):
            out = smp._CAPI_sample_node_classifiction(seeds, fan, False)
            for a, e in zip(out, exp):
                for x, z in zip(a, e):
                    assert np.array_equal(t2n(x), z)
    r1 = s1._CAPI_sample_node_classifiction(seeds, [10, 5], False, rng_seed=9)
    r2 = s2._CAPI_sample_node_classifiction(seeds, [10, 5], False, rng_seed=9)
    assert all(torch.equal(a, b) for x, y in zip(r1, r2) for a, b in zip(x, y))
    key, idx, dev = s1._CAPI_get_local_cache_hashmap_tensors()   # built on demand
    assert sorted(key[key >= 0].tolist()) == list(range(N))
    f1 = dgs.classes.P2PCacheFeatureServer(feat.pin_memory(), allnodes, 0)
    f2 = dgs.classes.P2PCacheFeatureServer.from_device_shard(feat.to(cuda), allnodes.to(cuda), N, 0)
    assert f1._mod_world == 1 and f2._mod_world == 1
    q = torch.randint(0, N, (30000,), generator=torch.Generator().manual_seed(4)).to(cuda)
    for fs in (f1, f2):
        for algo in (1, 2):
            assert torch.equal(fs._CAPI_get_feature(q, algo).cpu(), feat[q.cpu()])


def test_from_device_shard_with_hash_table(dgs, cuda):
    """from_device_shard(s) with a partial cache (hash path) and a pinned host fallback."""
    N, D = 3000, 36
    indptr, indices, _ = dgs_synth.make_csr(N, 50000, seed=42, classes=6)
    feat = dgs_synth.feature_rows(torch.arange(N), D)
    cache = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:1000]
    sub = oracle.extract_indptr(t2n(cache), t2n(indptr))
    sub_idx = oracle.extract_edge_data(t2n(cache), t2n(indptr), sub, t2n(indices))
    smp = dgs.classes.P2PCacheSampler.from_device_shards(
        torch.from_numpy(sub).to(cuda), torch.from_numpy(sub_idx).to(cuda), None, cache.to(cuda), N, 0,
        cpu_indptr=indptr.pin_memory(), cpu_indices=indices.pin_memory())
    assert smp._mod_world == 0
    seeds = torch.arange(0, N, 7).to(cuda)
    out = smp._CAPI_sample_node_classifiction(seeds, [-1], False)[0]
    e = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices), 1)[0]
    assert all(np.array_equal(t2n(x), z) for x, z in zip(out, e))
    fs = dgs.classes.P2PCacheFeatureServer.from_device_shard(feat[cache].to(cuda), cache.to(cuda), N, 0,
                                                             cpu_data=feat.pin_memory())
    q = torch.randint(0, N, (9000,), generator=torch.Generator().manual_seed(2)).to(cuda)
    assert torch.equal(fs._CAPI_get_feature(q).cpu(), feat[q.cpu()])


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
@pytest.mark.parametrize("source", ["cuda", "pinned", "server_mod", "server_hash"])
def test_batch_loader_equals_separate_calls(dgs, cuda, idt, source):
    """BatchLoader (one enqueue, device-side frontier count) == sample + extract + label gather
    made as three plugin calls with the same RNG seed; covers the adaptive output bound (incl. a
    batch that overflows it) and pinned-host seeds."""
    N, D = 6000, 100
    indptr, indices, _ = dgs_synth.make_csr(N, 120000, seed=44, classes=6, id_dtype=idt)
    feat = dgs_synth.feature_rows(torch.arange(N), D)
    labels = (torch.arange(N) % 47).to(cuda)
    smp = dgs.classes.CSRSampler(indptr.to(idt).to(cuda), indices.to(cuda))
    if source == "cuda":
        src = feat.to(cuda)
        ext = lambda q: dgs.ops._CAPI_cuda_index_select(src, q)
    elif source == "pinned":
        src = feat.pin_memory()
        ext = lambda q: dgs.ops._CAPI_cuda_index_select(src, q)
    elif source == "server_mod":
        src = dgs.classes.P2PCacheFeatureServer(feat.pin_memory(), torch.arange(N), 0)
        assert src._mod_world == 1
        ext = src._CAPI_get_feature
    else:
        cache = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:2000]
        src = dgs.classes.P2PCacheFeatureServer(feat.pin_memory(), cache, 0)
        assert src._mod_world == 0
        ext = src._CAPI_get_feature
    loader = dgs.classes.BatchLoader(smp, src, labels)
    g = torch.Generator().manual_seed(8)
    lab_host = torch.empty(64, dtype=torch.int64).pin_memory()
    # batch sizes: small first (tight adaptive bound), then seeds of high degree to overflow it
    deg = indptr[1:] - indptr[:-1]
    order = torch.argsort(deg)
    batches = [order[:64], order[-64:], torch.randperm(N, generator=g)[:64], order[-64:]]
    for bi, sd in enumerate(batches):
        sd = sd.to(idt)
        seeds = sd.pin_memory() if bi % 2 else sd.to(cuda)
        blocks, x, y = loader.load(seeds, [10, 5], False, rng_seed=100 + bi, labels_out=lab_host)
        ref = smp._CAPI_sample_node_classifiction(sd.to(cuda), [10, 5], False, rng_seed=100 + bi)
        assert len(blocks) == len(ref)
        for a, b in zip(blocks, ref):
            for u, v in zip(a, b):
                assert torch.equal(u, v)
        fr = ref[-1][1]
        assert x.shape == (fr.numel(), D)
        assert torch.equal(x, ext(fr)) and torch.equal(x.cpu(), feat[fr.cpu().long()])
        assert torch.equal(y, labels[sd.to(cuda).long()]) and torch.equal(lab_host, y.cpu())
    with pytest.raises(RuntimeError):
        loader.load(batches[0].to(idt).to(cuda), [-1, 5])


@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
def test_build_blocks_csc(dgs, cuda, idt):
    """DistGNN.dataloading.build_blocks: coo_row of every hop is ascending, the CSC row pointer equals
    searchsorted, blocks come input-side first with the caller's srcdata / dstdata ids."""
    from DistGNN.dataloading import NID, build_blocks
    N = 5000
    indptr, indices, _ = dgs_synth.make_csr(N, 80000, seed=23, classes=6, id_dtype=idt)
    smp = dgs.classes.CSRSampler(indptr.to(idt).to(cuda), indices.to(cuda))
    seeds = torch.randperm(N, generator=torch.Generator().manual_seed(2))[:300].to(idt).to(cuda)
    batch = smp._CAPI_sample_node_classifiction(seeds, [10, 5, 3], False, rng_seed=4)
    blocks = build_blocks(batch)
    assert len(blocks) == 3 and torch.equal(blocks[-1].dstdata[NID], seeds)
    for blk, (sd, fr, row, col) in zip(blocks, reversed(batch)):
        ip, src, eid = blk.csc()
        r = t2n(row)
        assert np.all(np.diff(r) >= 0)
        assert ip.dtype == idt and np.array_equal(t2n(ip), np.searchsorted(r, np.arange(sd.numel() + 1)))
        assert torch.equal(src, col) and torch.equal(eid.cpu(), torch.arange(row.numel(), dtype=idt))
        assert blk.num_src_nodes() == fr.numel() and blk.num_dst_nodes() == sd.numel()
        assert blk.num_edges() == row.numel() and torch.equal(blk.srcdata[NID], fr)
        assert int(blk.in_degrees().sum()) == row.numel()
    for a, b in zip(blocks[:-1], blocks[1:]):       # chained: dst of the outer hop = src of the next
        assert torch.equal(a.dstdata[NID], b.srcdata[NID])
    # rows with no edges at both ends, empty input, and the sortedness check
    row = torch.tensor([2, 2, 5], dtype=idt, device=cuda)
    assert dgs.ops.coo_rows_to_indptr(row, 8, check_sorted=True).tolist() == [0, 0, 0, 2, 2, 2, 3, 3, 3]
    assert dgs.ops.coo_rows_to_indptr(row[:0], 3).tolist() == [0, 0, 0, 0]
    assert dgs.ops.coo_rows_to_indptr(row[:0], 0).tolist() == [0]
    with pytest.raises(RuntimeError, match="ascending"):
        dgs.ops.coo_rows_to_indptr(torch.tensor([3, 1], dtype=idt, device=cuda), 4, check_sorted=True)
    with pytest.raises(RuntimeError, match="ascending"):
        dgs.ops.coo_rows_to_indptr(torch.tensor([1, 4], dtype=idt, device=cuda), 4, check_sorted=True)


def test_batch_loader_iter_many_equals_load_many(dgs, cuda):
    """BatchLoader.iter_many (two streams: the sampling of group g + 1 overlaps the extracts of
    group g; double-buffered hop sizes) yields exactly what load_many returns group by group."""
    N, D = 9000, 100
    indptr, indices, _ = dgs_synth.make_csr(N, 200000, seed=46, classes=8)
    feat = dgs_synth.feature_rows(torch.arange(N), D)
    labels = (torch.arange(N) % 47).to(cuda)
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda))
    loader = dgs.classes.BatchLoader(smp, feat.to(cuda), labels)
    g = torch.Generator().manual_seed(9)
    groups = [torch.randperm(N, generator=g)[:4 * 128].reshape(4, 128) for _ in range(7)]
    src = [grp.pin_memory() if i % 2 else grp.to(cuda) for i, grp in enumerate(groups)]
    rng = lambda gi: [1000 * gi + b for b in range(4)]
    got = list(loader.iter_many(src, [10, 5], False, rng))
    assert len(got) == 7
    for gi, res in enumerate(got):
        ref = loader.load_many(groups[gi].to(cuda), [10, 5], False, rng(gi))
        assert len(res) == len(ref) == 4
        for (blocks, x, y), (rb, rx, ry) in zip(res, ref):
            assert all(torch.equal(u, v) for p_, q_ in zip(blocks, rb) for u, v in zip(p_, q_))
            assert torch.equal(x, rx) and torch.equal(y, ry)
            assert torch.equal(x.cpu(), feat[blocks[-1][1].cpu()])


def test_guard_bands_stay_untouched(dgs, cuda):
    """compute-sanitizer is closed on the GPU pool, so out-of-bounds WRITES are hunted with guard
    bands: every output / workspace handed to the C-ABI sits inside a larger buffer filled with a
    sentinel, and the bands on both sides must be intact afterwards (extract: every algo and odd row
    counts; whole-batch sampling: arena, counts and workspace)."""
    import ctypes as C
    from dgs import _lib
    from dgs._util import stream
    l = _lib.lib()
    G = 4096                                                    # guard bytes on each side
    N, D = 5000, 100
    feat = dgs_synth.feature_rows(torch.arange(N), D).to(cuda)
    g = torch.Generator().manual_seed(3)
    for algo in (1, 2, 3, 7):
        for R in (1, 31, 33, 4097):
            nids = torch.randint(0, N, (R,), generator=g).to(cuda)
            buf = torch.full((2 * G + R * D * 4,), 0xA5, dtype=torch.uint8, device=cuda)
            _lib.check(l.dgs_index_select(feat.data_ptr(), D * 4, 1, nids.data_ptr(), R,
                                          buf.data_ptr() + G, algo, stream()))
            assert bool((buf[:G] == 0xA5).all()) and bool((buf[G + R * D * 4:] == 0xA5).all()), (algo, R)
            assert torch.equal(buf[G:G + R * D * 4].view(torch.float32).reshape(R, D), feat[nids])
    indptr, indices, _ = dgs_synth.make_csr(N, 120000, seed=9)
    smp = dgs.classes.CSRSampler(indptr.to(cuda), indices.to(cuda))
    pipe = smp._pipe
    fan = [10, 5]
    for B, S in ((1, 300), (3, 200)):
        pl = pipe._plan_many(B, S, fan)
        L, es = 2, 8
        nbytes = B * pl["total"] * es + B * 2 * L * 8
        arena = torch.full((2 * G + nbytes,), 0xA5, dtype=torch.uint8, device=cuda)
        ws = torch.full((2 * G + pl["ws_bytes"],), 0xA5, dtype=torch.uint8, device=cuda)
        fo = _lib.i64_array(fan)
        _lib.check(l.dgs_sample_blocks_multi_ws_init(ws.data_ptr() + G, pl["ws_bytes"], 1, B, S, L, fo, N, stream()))
        seeds = torch.randint(0, N, (B, S), generator=g).to(cuda)
        base = arena.data_ptr() + G
        for li, (of, orow, ocol) in enumerate(pl["offs"]):
            pl["a_fr"][li], pl["a_row"][li], pl["a_col"][li] = base + of * es, base + orow * es, base + ocol * es
        rng = (C.c_uint64 * B)(*range(1, B + 1))
        for _ in range(3):
            _lib.check(l.dgs_sample_blocks_multi(C.byref(pipe._graph), B, seeds.data_ptr(), S * es, S, L, fo, 0, rng,
                                                 pl["a_fr"], pl["a_row"], pl["a_col"], pl["total"] * es,
                                                 pl["cap_edges"], pl["cap_front"], base + B * pl["total"] * es,
                                                 ws.data_ptr() + G, pl["ws_bytes"], pl["counts_ptr"], 1, stream()))
        torch.cuda.synchronize()
        for t, n in ((arena, nbytes), (ws, pl["ws_bytes"])):
            assert bool((t[:G] == 0xA5).all()) and bool((t[G + n:] == 0xA5).all()), (B, S)
        ref = smp._CAPI_sample_node_classifiction(seeds[B - 1], fan, False, rng_seed=B)
        cnt = pl["counts_np"].tolist()[(B - 1) * 4:B * 4]
        assert cnt == [ref[0][2].numel(), ref[0][1].numel(), ref[1][2].numel(), ref[1][1].numel()]


def test_p2p_server_single_rank(dgs, cuda):
    t = torch.arange(24, dtype=torch.float32, device=cuda).reshape(6, 4)
    srv = dgs.classes.TensorP2PServer(t)
    loc = srv._CAPI_get_local_device_tensor()
    assert loc.shape == (6, 4) and torch.equal(loc, t) and loc.data_ptr() != t.data_ptr()
    flat = srv._CAPI_get_device_tensor(0)
    assert flat.shape == (24,) and torch.equal(flat, t.reshape(-1))
    with pytest.raises(RuntimeError):
        srv._CAPI_get_device_tensor(1)
    with pytest.raises(RuntimeError, match="> 0"):
        dgs.classes.TensorP2PServer(torch.empty(0, 4, device=cuda))
    srv.close()


def test_frontier_heat_matches_oracle(dgs, cuda):
    indptr, indices, probs = small_graph(3000, 60000, seed=6, weights=True)
    seeds = torch.randperm(3000, generator=torch.Generator().manual_seed(0))[:500]
    heat = torch.zeros(3000)
    heat[seeds] = torch.rand(500) + 0.1
    for pr in (None, probs):
        if pr is None:
            got = dgs.ops._CAPI_compute_frontier_heat(seeds.to(cuda), indptr.to(cuda), indices.to(cuda), heat.to(cuda), 10, 0)
        else:
            got = dgs.ops._CAPI_compute_frontier_heat_with_bias(seeds.to(cuda), indptr.to(cuda), indices.to(cuda), pr.to(cuda), heat.to(cuda), 10, 0)
        exp = oracle.frontier_heat(t2n(seeds), t2n(indptr), t2n(indices), None if pr is None else t2n(pr), t2n(heat), 10, 0)
        # float atomics: summation order differs -> tolerance 1e-4 relative
        assert np.allclose(t2n(got), exp, rtol=1e-4, atol=1e-5)


def test_get_node_heat_reference_semantics(dgs, cuda):
    """DistGNN.cache.get_node_heat restated with the oracle's heat kernel
    (python/DistGNN/cache/cache_value.py:6-53)."""
    from DistGNN.cache import get_node_heat
    N = 2000
    indptr, indices, probs = dgs_synth.make_csr(N, 40000, seed=8, weights=True, classes=6)
    nodes = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:300]
    for pr in (None, probs):
        for mode in ("cuda", "uva"):
            sh, fh = get_node_heat(indptr.clone(), indices.clone(), nodes, [5, 10], None if pr is None else pr.clone(), mode)
            samp = np.zeros(N, np.float32)
            seeds_heat = np.zeros(N, np.float32)
            seeds_heat[t2n(nodes)] = 1
            seeds = t2n(nodes)
            for k in (10, 5):
                fr = oracle.frontier_heat(seeds, t2n(indptr), t2n(indices), None if pr is None else t2n(pr), seeds_heat, k, 0)
                samp += seeds_heat
                seeds_heat = seeds_heat + fr
                seeds = np.nonzero(seeds_heat > 0)[0]
            assert np.allclose(t2n(sh), samp, rtol=1e-4, atol=1e-5)
            assert np.allclose(t2n(fh), samp + fr, rtol=1e-4, atol=1e-5)


def test_pin_memory_roundtrip(dgs, cuda):
    t = torch.arange(1000, dtype=torch.float32).reshape(100, 10).clone()
    assert not t.is_pinned()
    dgs.ops._CAPI_tensor_pin_memory(t)
    out = dgs.ops._CAPI_cuda_index_select(t, torch.tensor([3, 99], device=cuda))
    assert torch.equal(out.cpu(), t[[3, 99]])
    dgs.ops._CAPI_tensor_unpin_memory(t)


# ------------------------------------------------------------------ full-size properties
# BASELINE configs at (or scaled to) their shape: (nodes, edges, feature dim, dtype, batch, fan-out).
# products is configs[1] at full size; the papers100M / friendster cases keep the config's row
# format (128 x fp32 / 256 x bf16 = 512-byte rows), mean degree, batch and fan-out - so they take
# the same kernels (friendster: Floyd in 32 registers for k = 20, one kernel per phase for the 14 M
# padded slots of the last hop, the warp-autonomous gather) - on 1/32 resp. 1/16 of the nodes, which
# is what fits next to torch's checking code on one GPU in seconds.
_SHAPE_CASES = {
    "products": dgs_synth.SHAPES["products"] + (1024, [15, 10, 5]),
    "papers100M/32": (111_059_956 // 32, 1_615_685_872 // 32, 128, torch.float32, 1024, [15, 10, 5]),
    "friendster/16": (65_608_366 // 16, 1_806_067_135 // 16, 256, torch.bfloat16, 4096, [20, 15, 10]),
}


@pytest.mark.parametrize("case", list(_SHAPE_CASES))
def test_full_size_products_properties(dgs, cuda, case):
    """BASELINE configs[1] at full size (2.45 M nodes, ~61 M edges, 100-dim fp32, batch 1024,
    [15,10,5]) and configs 4 / 5 in their row format, batch and fan-out (_SHAPE_CASES) - too large
    for the CPU oracle, so size-independent properties are checked on the GPU
    with torch ops: every sampled (row, col) is an edge of the graph, per-seed counts are
    min(deg, k), no neighbour position is taken twice, the frontier is the first-occurrence unique
    of cat(seeds, cols), relabelled ids invert through the frontier, extract composes
    (gather(gather(T, p), q) == gather(T, p[q])) and reproduces the closed-form feature rows."""
    N, E, D, dt, batch, fan = _SHAPE_CASES[case]
    indptr, indices, _ = dgs_synth.make_csr(N, E, device=cuda)
    feat = dgs_synth.make_features(N, D, dt, device=cuda)
    smp = dgs.classes.CSRSampler(indptr, indices)
    seeds = dgs_synth.seed_batches(N, batch, 1, seed=7, device=cuda)[0]
    out = smp._CAPI_sample_node_classifiction(seeds, fan, False, rng_seed=99)
    cur = seeds
    # edge keys of the whole graph, sorted once: (src << 32) | dst  (N < 2^31)
    deg_all = indptr[1:] - indptr[:-1]
    for (s_, f_, r_, c_), k in zip(out, fan[::-1]):
        assert torch.equal(s_, cur)
        deg = deg_all[cur]
        cnt = torch.clamp(deg, max=k)
        assert r_.numel() == int(cnt.sum())
        # rows are seed-major and each seed owns exactly cnt edges
        assert torch.equal(torch.bincount(r_, minlength=cur.numel()), cnt)
        assert bool((r_[1:] >= r_[:-1]).all())
        src, dst = cur[r_], f_[c_]
        # membership: every (src, dst) appears in src's adjacency.  Search dst inside the row with a
        # per-edge scan over a sorted copy of the row segment (rows are short after clamping the
        # search to the edge's own row): use a global sorted key array
        seg_start = indptr[src]
        # position of each sampled edge inside its seed's block
        first = torch.cumsum(cnt, 0) - cnt
        j = torch.arange(r_.numel(), device=cuda) - first[r_]
        full = deg[r_] <= k
        # copy path: exactly the CSR row, in order
        assert torch.equal(dst[full], indices[(seg_start + j)[full]])
        # sampled path: dst must be a neighbour; verify with sorted (src, dst) keys of the touched rows
        rows_u = torch.unique(src[~full])
        if rows_u.numel():
            rdeg = deg_all[rows_u]
            owner = torch.repeat_interleave(rows_u, rdeg)
            pos = torch.arange(int(rdeg.sum()), device=cuda) - torch.repeat_interleave(torch.cumsum(rdeg, 0) - rdeg, rdeg)
            nb = indices[indptr[owner] + pos]
            keys = torch.sort(owner * (1 << 32) + nb).values
            q = src[~full] * (1 << 32) + dst[~full]
            at = torch.searchsorted(keys, q)
            assert bool((keys[at.clamp(max=keys.numel() - 1)] == q).all())
            # without replacement: a seed never returns more copies of a neighbour than its row holds
            qs, qc = torch.unique(q, return_counts=True)
            lo = torch.searchsorted(keys, qs)
            hi = torch.searchsorted(keys, qs, right=True)
            assert bool((qc <= hi - lo).all())
        # frontier = first-occurrence unique of cat(seeds, cols)
        allids = torch.cat([cur, dst])
        uniq, inv = torch.unique(allids, return_inverse=True)
        firstpos = torch.full((uniq.numel(),), allids.numel(), device=cuda, dtype=torch.int64)
        firstpos.scatter_reduce_(0, inv, torch.arange(allids.numel(), device=cuda), reduce="amin")
        assert torch.equal(f_, uniq[torch.argsort(firstpos)])
        assert torch.equal(f_[: cur.numel()], cur)
        cur = f_
    x = dgs.ops._CAPI_cuda_index_select(feat, cur)
    assert torch.equal(x, dgs_synth.feature_rows(cur, D, dt))
    q = torch.randint(0, cur.numel(), (100000,), device=cuda)
    assert torch.equal(dgs.ops._CAPI_cuda_index_select(x, q), dgs.ops._CAPI_cuda_index_select(feat, cur[q]))
    for algo in (1, 2):
        assert torch.equal(dgs.ops._CAPI_cuda_index_select(feat, cur, algo), x)
