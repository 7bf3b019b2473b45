"""Multi-rank worker (launched by tests/test_multigpu.py through torch.distributed.run, one rank per
GPU).  Runs the reference's own 2-rank test scenarios (tests/test_p2p_server.py,
test_build_sampler.py, test_feature_server.py, test_sampler_uniform.py, test_nccl.py) with
assertions, plus a sharded parity check against the CPU oracle.  Prints "RANK r OK"."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dist-gnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dgs  # noqa: E402
import dgs_synth  # noqa: E402
import oracle  # noqa: E402
from DistGNN.dist import create_communicator  # noqa: E402


def t2n(t):
    return t.detach().cpu().numpy()


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    create_communicator(world)
    assert dgs.ops._Test_GetLocalRank() == rank and dgs.ops._Test_GetWorldSize() == world
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_reference_tests.json")))
    G = kat["graph"]

    # ---- tests/test_nccl.py: variable-length all-gather
    n = (rank * 3 + 1) % 11
    got = dgs.ops._Test_NCCLTensorAllGather(torch.ones(n, device=dev) * rank)
    assert len(got) == world
    for r, t in enumerate(got):
        assert t.numel() == (r * 3 + 1) % 11 and bool((t == r).all())

    indptr = torch.tensor(G["indptr"]).pin_memory()
    indices = torch.tensor(G["indices"]).pin_memory()
    probs = torch.tensor(G["probs"]).pin_memory()
    if world == 2:
        # ---- tests/test_p2p_server.py
        cache = torch.tensor(kat["p2p_server"]["cache_nids"][rank]).to(dev)
        sub = dgs.ops._Test_ExtractIndptr(cache, indptr)
        assert sub.tolist() == kat["p2p_server"]["sub_indptr"][rank]
        srv = dgs.classes.TensorP2PServer(sub)
        assert srv._CAPI_get_local_device_tensor().tolist() == kat["p2p_server"]["sub_indptr"][rank]
        for r in range(2):
            assert srv._CAPI_get_device_tensor(r).tolist() == kat["p2p_server"]["sub_indptr"][r]
        # ---- tests/test_build_sampler.py
        k = kat["build_sampler"]
        s = dgs.classes.P2PCacheSampler(indptr, indices, probs, torch.tensor(k["cache_nids"][rank]), rank)
        li, lx, lp = s._CAPI_get_local_cache_structure_tensors()
        if rank == 1:
            assert li.tolist() == k["rank1_local_indptr"] and lx.tolist() == k["rank1_local_indices"]
            assert torch.equal(lp.cpu(), torch.tensor(k["rank1_local_probs"]))
        key, idx, devid = s._CAPI_get_local_cache_hashmap_tensors()
        assert key.numel() == k["hash_capacity"]
        ent = {kk: (d, i) for kk, d, i in zip(key.tolist(), devid.tolist(), idx.tolist()) if kk >= 0}
        for q, d, i in zip(k["queries"], k[f"rank{rank}_dev"], k[f"rank{rank}_idx"]):
            assert ent.get(q, (-1, -1)) == (d, i), (rank, q, ent.get(q), d, i)
        # ---- tests/test_sampler_bias.py / test_sampler_uniform.py
        su = dgs.classes.P2PCacheSampler(indptr, indices, torch.Tensor(), torch.tensor(k["cache_nids"][rank]), rank)
        for smp in (su, s):
            out = smp._CAPI_sample_node_classifiction(torch.tensor([0, 3, 5]).to(dev), [2, 2], False)
            seeds, frontier, row, col = out[0]
            assert row.tolist() == [0, 0, 2, 2] and frontier[:3].tolist() == [0, 3, 5]
            nb = frontier[col].tolist()
            assert set(nb[:2]) <= {1, 2, 3, 4} and set(nb[2:]) <= {6, 7, 8, 9, 10}
            assert len(set(nb[:2])) == 2 and len(set(nb[2:])) == 2
            full = smp._CAPI_sample_node_classifiction(torch.tensor([0, 3, 5]).to(dev), [-1], False)[0]
            assert full[1].tolist() == kat["all_neighbors"]["frontier"]
            assert full[2].tolist() == kat["all_neighbors"]["relabeled_row"]
            assert full[3].tolist() == kat["all_neighbors"]["relabeled_col"]
        # ---- tests/test_feature_server.py
        f = kat["feature_server"]
        feature = torch.arange(0, 100, 1).float().pin_memory().reshape(10, 10)
        fs = dgs.classes.P2PCacheFeatureServer(feature, torch.tensor(f["cache_nids"][rank]).to(dev), rank)
        assert torch.equal(fs._CAPI_get_gpu_feature().cpu(), feature[f["cache_nids"][rank]])
        out = fs._CAPI_get_feature(torch.tensor(f["query"]).to(dev))
        assert torch.equal(out.cpu(), feature[f["query"]])

    # ---- sharded parity on a larger graph: node n on GPU n mod world, 10 % left un-cached (host)
    N, E, D = 40000, 900000, 100
    ip, ix, pr = dgs_synth.make_csr(N, E, seed=31, weights=True, classes=8)
    feat = dgs_synth.feature_rows(torch.arange(N), D)
    ipp, ixp, prp, fp = ip.pin_memory(), ix.pin_memory(), pr.pin_memory(), feat.pin_memory()
    mine = torch.arange(rank, N, world)
    mine = mine[mine % 10 != 9]
    # overlap: every rank also caches the first 500 nodes (local must win)
    mine = torch.unique(torch.cat([mine, torch.arange(500)]))
    maxdeg = int((ip[1:] - ip[:-1]).max())
    seeds = torch.randperm(N, generator=torch.Generator().manual_seed(100 + rank))[:300].to(dev)
    for bias in (False, True):
        smp = dgs.classes.P2PCacheSampler(ipp, ixp, prp if bias else torch.Tensor(), mine, rank)
        exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(ip), t2n(ix), 2)
        for fan in ([-1, -1], [maxdeg, maxdeg]):
            out = smp._CAPI_sample_node_classifiction(seeds, fan, False)
            for a, e in zip(out, exp):
                for x, z in zip(a, e):
                    assert np.array_equal(t2n(x), z)
        out = smp._CAPI_sample_node_classifiction(seeds, [10, 5], False, rng_seed=1)
        cur = seeds.cpu()
        for (s_, f_, r_, c_), k in zip(out, (5, 10)):
            f = f_.cpu()
            rg, cg = cur[r_.cpu()], f[c_.cpu()]
            deg = ip[cur + 1] - ip[cur]
            assert r_.numel() == int(torch.clamp(deg, max=k).sum())
            # every (row, col) is an edge of the graph
            o = 0
            for sd, d in zip(cur.tolist(), deg.tolist()):
                c = min(d, k)
                nb = set(ix[ip[sd]:ip[sd + 1]].tolist())
                assert set(cg[o:o + c].tolist()) <= nb and set(rg[o:o + c].tolist()) <= {sd}
                o += c
            cur = f
        smp.close()
    fs = dgs.classes.P2PCacheFeatureServer(fp, mine.to(dev), rank)
    q = torch.randint(0, N, (50000,), generator=torch.Generator().manual_seed(rank)).to(dev)
    for algo in (0, 1):
        assert torch.equal(fs._CAPI_get_feature(q, algo).cpu(), feat[q.cpu()])
    # lookups resolve like the reference's insertion order says
    lists = [t2n(t) for t in dgs.ops._allgather_tensors(mine.to(dev))]
    key, idx, devid = oracle.hashmap_build(lists, rank)
    ed, ei = oracle.hashmap_lookup(key, idx, devid, t2n(q))
    ok, oi, od = fs._CAPI_get_local_cache_hashmap_tensors()
    ent = {kk: (d, i) for kk, d, i in zip(ok.tolist(), od.tolist(), oi.tolist()) if kk >= 0}
    for qq, d, i in zip(t2n(q)[:5000].tolist(), ed[:5000].tolist(), ei[:5000].tolist()):
        assert ent.get(qq, (-1, -1)) == (d, i)
    dist.barrier()
    fs.close()

    # ---- "selfish" placement: every rank caches the SAME hot nodes (get_cache_nids_selfish), so the
    # all-gathered lists hold world * n ids but only n distinct keys - the table is sized for n
    # (hashmap.cu:20) and every hit must resolve to the local rank
    deg_all = ip[1:] - ip[:-1]
    hot = torch.argsort(deg_all, descending=True)[:N // 3]
    fs = dgs.classes.P2PCacheFeatureServer(fp, hot.to(dev), rank)
    assert fs._mod_world == 0
    assert torch.equal(fs._CAPI_get_feature(q).cpu(), feat[q.cpu()])
    ok, oi, od = fs._CAPI_get_local_cache_hashmap_tensors()
    live = ok >= 0
    assert int(live.sum()) == hot.numel() and bool((od[live] == rank).all())
    smp = dgs.classes.P2PCacheSampler(ipp, ixp, torch.Tensor(), hot, rank)
    out = smp._CAPI_sample_node_classifiction(seeds, [-1, -1], False)
    exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(ip), t2n(ix), 2)
    for a, e in zip(out, exp):
        for x, z in zip(a, e):
            assert np.array_equal(t2n(x), z)
    smp.close()
    dist.barrier()
    fs.close()

    # ---- exact modulo sharding (node n on GPU n % world, slot n // world): arithmetic owner, no
    # location table; shards generated directly on the device (from_device_shard[s] extensions)
    nids, sp, si, spr = dgs_synth.make_shard(N, E, rank, world, seed=31, device=dev, weights=True, classes=8)
    for bias in (False, True):
        smp = dgs.classes.P2PCacheSampler.from_device_shards(sp, si, spr if bias else None, nids, N, rank)
        assert smp._mod_world == world and smp._table is None
        exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(ip), t2n(ix), 2)
        for fan in ([-1, -1], [maxdeg, maxdeg]):
            out = smp._CAPI_sample_node_classifiction(seeds, fan, False)
            for a, e in zip(out, exp):
                for x, z in zip(a, e):
                    assert np.array_equal(t2n(x), z)
        a = smp._CAPI_sample_node_classifiction(seeds, [15, 10, 5], False, rng_seed=3)
        b = smp._CAPI_sample_node_classifiction(seeds, [15, 10, 5], False, rng_seed=3)
        assert all(torch.equal(x, y) for u, v in zip(a, b) for x, y in zip(u, v))
        smp.close()
    fsm = dgs.classes.P2PCacheFeatureServer.from_device_shard(
        dgs_synth.feature_rows(nids, D), nids, N, rank)
    assert fsm._mod_world == world
    for algo in (0, 1, 2):
        assert torch.equal(fsm._CAPI_get_feature(q, algo).cpu(), feat[q.cpu()])
    # id-exchange variant (NCCL all-to-all of ids, owners gather locally, rows come back): same rows,
    # for request counts that differ per rank, an empty request and ids that all have one owner
    gq = torch.Generator().manual_seed(4000 + rank)
    for n_req in (0, 1, 5000 + 777 * rank):
        qe = torch.randint(0, N, (n_req,), generator=gq).to(dev)
        assert torch.equal(fsm.get_feature_exchange(qe).cpu(), feat[qe.cpu()])
    qe = (torch.randint(0, N // world - 1, (3000,), generator=gq) * world + (world - 1)).to(dev)
    assert torch.equal(fsm.get_feature_exchange(qe).cpu(), feat[qe.cpu()])
    assert torch.equal(fsm.get_feature_exchange(qe), fsm._CAPI_get_feature(qe))
    # one-call loader over the sharded sampler + feature server == the separate plugin calls
    smp = dgs.classes.P2PCacheSampler.from_device_shards(sp, si, None, nids, N, rank)
    loader = dgs.classes.BatchLoader(smp, fsm)
    for it in range(3):
        blocks, x, y = loader.load(seeds.cpu().pin_memory(), [10, 5], False, rng_seed=50 + it)
        ref = smp._CAPI_sample_node_classifiction(seeds, [10, 5], False, rng_seed=50 + it)
        assert all(torch.equal(a, b) for u, v in zip(blocks, ref) for a, b in zip(u, v))
        assert y is None and torch.equal(x.cpu(), feat[ref[-1][1].cpu()])
    # B batches per sampling launch over the sharded sampler, sequential (load_many) and two-stream
    # (iter_many, gathers capped at 2 CTAs per SM): both == single calls
    grp = [torch.randperm(N, generator=torch.Generator().manual_seed(7 * rank + j))[:3 * 100].reshape(3, 100)
           for j in range(4)]
    rngf = lambda gi: [900 + 10 * gi + b for b in range(3)]
    many = [loader.load_many(gq.pin_memory(), [10, 5], False, rngf(gi)) for gi, gq in enumerate(grp)]
    piped = list(loader.iter_many([gq.to(dev) for gq in grp], [10, 5], False, rngf, 0, 2))
    for gi in range(4):
        for b in range(3):
            ref = smp._CAPI_sample_node_classifiction(grp[gi][b].to(dev), [10, 5], False, rng_seed=rngf(gi)[b])
            for res in (many[gi][b], piped[gi][b]):
                assert all(torch.equal(a, c) for u, v in zip(res[0], ref) for a, c in zip(u, v))
                assert torch.equal(res[1].cpu(), feat[ref[-1][1].cpu()])
    smp.close()
    dist.barrier()
    fsm.close()
    # ---- full replica on every rank (what the cache policy yields when everything fits in HBM):
    # no location table, every read is local; same results as the sharded layouts
    everything = torch.arange(N)
    for bias in (False, True):
        smp = dgs.classes.P2PCacheSampler(ipp, ixp, prp if bias else torch.Tensor(), everything, rank)
        assert smp._mod_world == -1 and smp._table is None
        exp = oracle.sample_blocks_all_neighbors(t2n(seeds), t2n(ip), t2n(ix), 2)
        for fan in ([-1, -1], [maxdeg, maxdeg]):
            out = smp._CAPI_sample_node_classifiction(seeds, fan, False)
            for a, e in zip(out, exp):
                for x, z in zip(a, e):
                    assert np.array_equal(t2n(x), z)
        key, idx, devid = smp._CAPI_get_local_cache_hashmap_tensors()      # built on demand: local wins
        live = key >= 0
        assert int(live.sum()) == N and bool((devid[live] == rank).all())
        assert torch.equal(idx[live], key[live])
        smp.close()
    fsr = dgs.classes.P2PCacheFeatureServer(fp, everything, rank)
    assert fsr._mod_world == -1
    for algo in (0, 1, 2):
        assert torch.equal(fsr._CAPI_get_feature(q, algo).cpu(), feat[q.cpu()])
    smp = dgs.classes.P2PCacheSampler(ipp, ixp, torch.Tensor(), everything, rank)
    loader = dgs.classes.BatchLoader(smp, fsr)
    blocks, x, _ = loader.load(seeds, [10, 5], False, rng_seed=77)
    ref = smp._CAPI_sample_node_classifiction(seeds, [10, 5], False, rng_seed=77)
    assert all(torch.equal(a, b) for u, v in zip(blocks, ref) for a, b in zip(u, v))
    assert torch.equal(x.cpu(), feat[ref[-1][1].cpu()])
    smp.close()
    dist.barrier()
    fsr.close()
    print(f"RANK {rank} OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
