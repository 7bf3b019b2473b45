"""GPU: differential tests against the UNMODIFIED reference compiled for sm_100a
(oracle/_ref/dgs.cpython-*.so, built by oracle/build_ref.sh; skipped when it is absent).

This is what pins the oracle and the CUDA path: the reference's own tests hold no expected
outputs, so its kernels are run here on the same inputs.  Deterministic ops must agree bit for
bit between reference, oracle and this repo.  When DGS_WRITE_GOLDEN is set the reference outputs
are also written to gpurun_out/golden/ (committed afterwards as tests/golden/ref_*.npz and
checked on the CPU by tests/test_golden_fixtures.py)."""
import os

import numpy as np
import pytest
import torch

import dgs_synth
import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def t2n(t):
    return t.detach().cpu().numpy()


def _dump(name, **arrays):
    if os.environ.get("DGS_WRITE_GOLDEN"):
        d = os.path.join(ROOT, "gpurun_out", "golden")
        os.makedirs(d, exist_ok=True)
        np.savez_compressed(os.path.join(d, name + ".npz"), **arrays)


def _graph(cuda, N=4000, E=90000, seed=21, weights=False):
    return dgs_synth.make_csr(N, E, seed=seed, weights=weights)


def test_ref_index_select(ref, dgs, cuda):
    feat = dgs_synth.feature_rows(torch.arange(5000), 100)
    nids = torch.randint(0, 5000, (20000,), generator=torch.Generator().manual_seed(0))
    r = ref.ops._CAPI_cuda_index_select(feat.to(cuda), nids.to(cuda))
    for algo in (1, 2):
        o = dgs.ops._CAPI_cuda_index_select(feat.to(cuda), nids.to(cuda), algo)
        assert torch.equal(r, o)
    assert np.array_equal(t2n(r), oracle.index_select(t2n(feat), t2n(nids)))
    _dump("ref_index_select", nids=t2n(nids)[:2000], out=t2n(r)[:2000])
    # pinned-host table + int64 labels (1-D)
    labels = torch.randint(0, 47, (5000,)).pin_memory()
    assert torch.equal(ref.ops._CAPI_cuda_index_select(labels, nids.to(cuda)),
                       dgs.ops._CAPI_cuda_index_select(labels, nids.to(cuda)))


def test_ref_extract_subcsr(ref, dgs, cuda):
    indptr, indices, probs = _graph(cuda, weights=True)
    nids = torch.randperm(4000, generator=torch.Generator().manual_seed(1))[:1500].to(cuda)
    ip, ix, pr = indptr.to(cuda), indices.to(cuda), probs.to(cuda)
    rs = ref.ops._Test_ExtractIndptr(nids, ip)
    os_ = dgs.ops._Test_ExtractIndptr(nids, ip)
    assert torch.equal(rs, os_)
    assert np.array_equal(t2n(rs), oracle.extract_indptr(t2n(nids), t2n(indptr)))
    for data in (ix, pr):
        rd = ref.ops._Test_ExtractEdgeData(nids, ip, rs, data)
        od = dgs.ops._Test_ExtractEdgeData(nids, ip, os_, data)
        assert torch.equal(rd, od)
        assert np.array_equal(t2n(rd), oracle.extract_edge_data(t2n(nids), t2n(indptr), t2n(rs), t2n(data)))
    _dump("ref_extract_subcsr", nids=t2n(nids), sub_indptr=t2n(rs))


def test_ref_relabel(ref, dgs, cuda):
    g = torch.Generator().manual_seed(2)
    seeds = torch.randperm(50000, generator=g)[:1024]
    col = torch.randint(0, 50000, (60000,), generator=g)
    row = seeds[torch.randint(0, 1024, (60000,), generator=g)]
    ru, rr = ref.ops._CAPI_cuda_sampled_tensor_relabel([seeds.to(cuda), col.to(cuda)], [row.to(cuda), col.to(cuda)])
    ou, orr = dgs.ops._CAPI_cuda_sampled_tensor_relabel([seeds.to(cuda), col.to(cuda)], [row.to(cuda), col.to(cuda)])
    assert torch.equal(ru, ou) and torch.equal(rr[0], orr[0]) and torch.equal(rr[1], orr[1])
    eu, (er, ec) = oracle.relabel([t2n(seeds), t2n(col)], [t2n(row), t2n(col)])
    assert np.array_equal(t2n(ru), eu) and np.array_equal(t2n(rr[0]), er) and np.array_equal(t2n(rr[1]), ec)
    _dump("ref_relabel", seeds=t2n(seeds), col=t2n(col)[:5000], row=t2n(row)[:5000],
          unique=t2n(ref.ops._CAPI_cuda_sampled_tensor_relabel([seeds.to(cuda), col[:5000].to(cuda)], [row[:5000].to(cuda)])[0]))


def test_ref_full_neighbor_sampling(ref, dgs, cuda):
    """num_picks >= max degree on the reference's uniform op == our num_picks = -1 == oracle."""
    indptr, indices, _ = _graph(cuda)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    seeds = torch.randperm(4000, generator=torch.Generator().manual_seed(3))[:800].to(cuda)
    rr, rc = ref.ops._CAPI_cuda_sample_neighbors(seeds, indptr.to(cuda), indices.to(cuda), maxdeg, False)
    orow, ocol = dgs.ops._CAPI_cuda_sample_neighbors(seeds, indptr.to(cuda), indices.to(cuda), -1, False)
    assert torch.equal(rr, orow) and torch.equal(rc, ocol)
    o2r, o2c = dgs.ops._CAPI_cuda_sample_neighbors(seeds, indptr.to(cuda), indices.to(cuda), maxdeg, False)
    assert torch.equal(rr, o2r) and torch.equal(rc, o2c)
    er, ec = oracle.sample_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices))
    assert np.array_equal(t2n(rr), er) and np.array_equal(t2n(rc), ec)
    _dump("ref_full_neighbor", seeds=t2n(seeds), row=t2n(rr)[:4000], col=t2n(rc)[:4000], nnz=np.array([rr.numel()]))


def test_ref_feature_server_and_hash_lookups(ref, dgs, cuda):
    """One rank, 30 % cached, misses from pinned host: extract is bit-exact and every key resolves
    to the same (device, slot) - the table layout itself is implementation-defined."""
    N, D = 6000, 100
    feat = dgs_synth.feature_rows(torch.arange(N), D).pin_memory()
    g = torch.Generator().manual_seed(4)
    cache = torch.randperm(N, generator=g)[:1800]
    q = torch.randint(0, N, (30000,), generator=g).to(cuda)
    rfs = ref.classes.P2PCacheFeatureServer(feat, cache.to(cuda), 0)
    ofs = dgs.classes.P2PCacheFeatureServer(feat, cache.to(cuda), 0)
    r = rfs._CAPI_get_feature(q)
    o = ofs._CAPI_get_feature(q)
    assert torch.equal(r, o) and torch.equal(r.cpu(), feat[q.cpu()])
    assert torch.equal(rfs._CAPI_get_gpu_feature(), ofs._CAPI_get_gpu_feature())
    ok, oi, od = ofs._CAPI_get_local_cache_hashmap_tensors()
    key, idx, dev = oracle.hashmap_build([t2n(cache)], 0)
    assert ok.numel() == len(key)
    ours = sorted(zip(t2n(ok)[t2n(ok) >= 0].tolist(), t2n(od)[t2n(ok) >= 0].tolist(), t2n(oi)[t2n(ok) >= 0].tolist()))
    exp = sorted(zip(key[key >= 0].tolist(), dev[key >= 0].tolist(), idx[key >= 0].tolist()))
    assert ours == exp
    _dump("ref_feature_server", cache=t2n(cache), q=t2n(q)[:1000], out=t2n(r)[:1000])
    del rfs


def test_ref_sampler_blocks_copy_path(ref, dgs, cuda):
    """Reference P2PCacheSampler vs ours, 2 hops, fan-out >= max degree: bit-exact blocks; and the
    reference's hash tensors resolve every cached id like ours do."""
    N = 1200
    indptr, indices, _ = dgs_synth.make_csr(N, 6000, seed=23)
    maxdeg = int((indptr[1:] - indptr[:-1]).max())
    ip, ix = indptr.pin_memory(), indices.pin_memory()
    g = torch.Generator().manual_seed(5)
    cache = torch.randperm(N, generator=g)[:400]
    seeds = torch.randperm(N, generator=g)[:32].to(cuda)
    rs = ref.classes.P2PCacheSampler(ip, ix, torch.Tensor(), cache, 0)
    os_ = dgs.classes.P2PCacheSampler(ip, ix, torch.Tensor(), cache, 0)
    rout = rs._CAPI_sample_node_classifiction(seeds, [maxdeg, maxdeg], False)
    oout = os_._CAPI_sample_node_classifiction(seeds, [maxdeg, maxdeg], False)
    eout = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(indptr), t2n(indices), 2)
    assert len(rout) == len(oout) == 2
    for a, b, e in zip(rout, oout, eout):
        for x, y, z in zip(a, b, e):
            assert torch.equal(x, y) and np.array_equal(t2n(x), z)
    rl = rs._CAPI_get_local_cache_structure_tensors()
    ol = os_._CAPI_get_local_cache_structure_tensors()
    assert torch.equal(rl[0], ol[0]) and torch.equal(rl[1], ol[1])
    rk, ri, rd = rs._CAPI_get_local_cache_hashmap_tensors()
    ok, oi, od = os_._CAPI_get_local_cache_hashmap_tensors()
    assert rk.numel() == ok.numel()
    f = lambda k, i, d: sorted(zip(t2n(k)[t2n(k) >= 0].tolist(), t2n(d)[t2n(k) >= 0].tolist(), t2n(i)[t2n(k) >= 0].tolist()))
    assert f(rk, ri, rd) == f(ok, oi, od)
    _dump("ref_sampler_blocks", seeds=t2n(seeds), cache=t2n(cache),
          frontier0=t2n(rout[0][1]), row0=t2n(rout[0][2]), col0=t2n(rout[0][3]),
          frontier1=t2n(rout[1][1])[:5000], nnz1=np.array([rout[1][2].numel()]))
    del rs


def test_ref_random_sampling_same_support(ref, dgs, cuda):
    """Random paths cannot match bit for bit (the reference seeds from std::random_device): both must
    return the same number of edges per seed and only valid neighbours."""
    indptr, indices, probs = _graph(cuda, weights=True)
    seeds = torch.randperm(4000, generator=torch.Generator().manual_seed(6))[:1000].to(cuda)
    ip, ix, pr = indptr.to(cuda), indices.to(cuda), probs.to(cuda)
    for k in (5, 15):
        rr, rc = ref.ops._CAPI_cuda_sample_neighbors(seeds, ip, ix, k, False)
        orow, ocol = dgs.ops._CAPI_cuda_sample_neighbors(seeds, ip, ix, k, False)
        assert torch.equal(rr, orow)
        br, bc = ref.ops._CAPI_cuda_sample_neighbors_bias(seeds, ip, ix, pr, k, False)
        obr, obc = dgs.ops._CAPI_cuda_sample_neighbors_bias(seeds, ip, ix, pr, k, False)
        assert torch.equal(br, obr)


def test_ref_gpu_timing_report(ref, dgs, cuda):
    """Not a pass/fail check of speed: times the reference's own kernels (compiled for sm_100a) and
    this repo on the SAME B200, same products-shaped workload, same plugin calls
    (sampler._CAPI_sample_node_classifiction + feature_server._CAPI_get_feature, everything cached
    on the GPU), and writes gpurun_out/ref_gpu_timing.json (BASELINE.md section 2, row
    "reference oracle on B200").  Asserts only that both return the same per-hop edge counts."""
    import json
    import time
    N, E, D, dt = dgs_synth.SHAPES["products"]
    ip, ix, _ = dgs_synth.make_csr(N, E, device=cuda)
    ft = dgs_synth.make_features(N, D, dt, device=cuda)
    ipc, ixc, ftc = ip.cpu().pin_memory(), ix.cpu().pin_memory(), ft.cpu().pin_memory()
    del ip, ix, ft
    torch.cuda.empty_cache()
    allnodes = torch.arange(N)
    fan = [15, 10, 5]
    seeds = dgs_synth.seed_batches(N, 1024, 40, device=cuda)
    res = {}
    counts = {}
    for name, mod in (("reference", ref), ("this_repo", dgs)):
        smp = mod.classes.P2PCacheSampler(ipc, ixc, torch.Tensor(), allnodes, 0)
        fs = mod.classes.P2PCacheFeatureServer(ftc, allnodes.to(cuda), 0)
        for i in range(5):
            b = smp._CAPI_sample_node_classifiction(seeds[i], fan, False)
            fs._CAPI_get_feature(b[-1][1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        edges = rows = 0
        for i in range(5, 35):
            b = smp._CAPI_sample_node_classifiction(seeds[i], fan, False)
            x = fs._CAPI_get_feature(b[-1][1])
            edges += sum(t[2].numel() for t in b)
            rows += x.shape[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        # sampling only / extract only (back-to-back launches on saved frontiers)
        e0.record()
        fronts = [smp._CAPI_sample_node_classifiction(seeds[i], fan, False)[-1][1].clone() for i in range(5, 35)]
        e1.record()
        torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1) / 30
        outs = [fs._CAPI_get_feature(f) for f in fronts]
        del outs
        torch.cuda.synchronize()
        e0.record()
        outs = [fs._CAPI_get_feature(f) for f in fronts]
        e1.record()
        torch.cuda.synchronize()
        ms_x = e0.elapsed_time(e1) / 30
        del outs
        res[name] = {"ms_per_step": ms, "batches_per_sec": 1e3 / ms, "sampled_edges_per_sec": edges / 30 / (ms * 1e-3),
                     "sampling_ms": ms_s, "extract_ms": ms_x,
                     "extract_algorithmic_gbps": sum(f.numel() for f in fronts) / 30 * (2 * D * 4 + 8) / (ms_x * 1e-3) / 1e9,
                     "rows_per_step": rows / 30, "edges_per_step": edges / 30}
        # deterministic part of the output: hop-1 edge count = sum(min(deg, 5)) for the same seeds
        b = smp._CAPI_sample_node_classifiction(seeds[0], fan, False)
        counts[name] = b[0][2].numel()
        del smp, fs
        torch.cuda.empty_cache()
    assert counts["reference"] == counts["this_repo"]
    for k_ in ("ms_per_step", "sampling_ms", "extract_ms"):
        res["speedup_" + k_] = res["reference"][k_] / res["this_repo"][k_]
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, "ref_gpu_timing.json"), "w"), indent=1)
    print(json.dumps(res))
