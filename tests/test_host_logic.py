"""CPU: host-side logic - seed iterator, synthetic generators, shard partitioning, and the N > 1
bootstrap / reduction logic on a world_size-2 gloo group (no GPU)."""
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dgs_synth
from DistGNN.dataloading import SeedGenerator
from DistGNN.dist import create_communicator, exchange_extract, owner_of, partition_seeds


def test_seed_generator_matches_reference_semantics():
    data = torch.arange(10)
    g = SeedGenerator(data, 4)
    batches = [b.tolist() for b in g]
    assert batches == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]] and len(g) == 3 and g.is_finished()
    g = SeedGenerator(data, 4, drop_last=True)
    assert [b.tolist() for b in g] == [[0, 1, 2, 3], [4, 5, 6, 7]] and len(g) == 2
    torch.manual_seed(0)
    g = SeedGenerator(data, 3, shuffle=True)
    seen = torch.cat(list(g))
    assert sorted(seen.tolist()) == list(range(10)) and seen.tolist() != list(range(10))
    # second epoch reshuffles
    again = torch.cat(list(g))
    assert sorted(again.tolist()) == list(range(10))


def test_seed_generator_edges():
    data = torch.arange(12)
    g = SeedGenerator(data, 4)
    assert len(g) == 3 and g.is_finished()            # usable before the first epoch starts
    with pytest.raises(StopIteration):
        next(g)
    it = iter(g)
    first = next(it)
    assert first.data_ptr() == data.data_ptr() and not g.is_finished()    # a view, not a copy
    assert [next(it).tolist(), next(it).tolist()] == [[4, 5, 6, 7], [8, 9, 10, 11]]
    assert g.step == g.last_step == 3 and g.is_finished()
    assert [b.tolist() for b in it][0] == [0, 1, 2, 3]      # iter() starts a new epoch, as in the reference
    assert len(SeedGenerator(data, 5, drop_last=True)) == 2 and len(SeedGenerator(data, 5)) == 3
    assert len(SeedGenerator(data[:0], 5)) == 0 and list(SeedGenerator(data[:0], 5)) == []
    assert [b.tolist() for b in SeedGenerator(data[:1], 5, shuffle=True)] == [[0]]
    with pytest.raises(ValueError):
        SeedGenerator(data, 0)


def test_build_blocks_orders_blocks_input_side_first():
    """Host logic of DistGNN.dataloading.build_blocks (the CSC kernel itself is a GPU test)."""
    from DistGNN.dataloading import NID, build_blocks
    hop0 = (torch.tensor([7, 9]), torch.tensor([7, 9, 3]), torch.tensor([0, 1]), torch.tensor([2, 0]))
    hop1 = (hop0[1], torch.tensor([7, 9, 3, 5]), torch.tensor([0, 2, 2]), torch.tensor([1, 3, 0]))
    blocks = build_blocks([hop0, hop1], capi=object())
    assert [b.num_dst_nodes() for b in blocks] == [3, 2] and [b.num_src_nodes() for b in blocks] == [4, 3]
    assert [b.num_edges() for b in blocks] == [3, 2]
    assert torch.equal(blocks[0].dstdata[NID], blocks[1].srcdata[NID])
    src, dst = blocks[1].edges()
    assert src.tolist() == [2, 0] and dst.tolist() == [0, 1]


def test_synth_is_deterministic_and_shaped():
    ip, ix, pr = dgs_synth.make_csr(5000, 120000, seed=3, weights=True)
    ip2, ix2, pr2 = dgs_synth.make_csr(5000, 120000, seed=3, weights=True)
    assert torch.equal(ip, ip2) and torch.equal(ix, ix2) and torch.equal(pr, pr2)
    deg = ip[1:] - ip[:-1]
    assert ip[0] == 0 and ip[-1] == ix.numel() and (deg >= 0).all()
    assert 0.8 * 120000 < ix.numel() < 1.2 * 120000          # edge count close to the target
    assert (deg == 0).float().mean() > 0.03                    # isolated nodes exist
    assert deg.max() > 20 * deg.float().mean()                 # heavy tail
    assert ix.min() >= 0 and ix.max() < 5000 and (pr > 0).all() and pr.max() <= 4.0
    ip3, _, _ = dgs_synth.make_csr(5000, 120000, seed=4)
    assert not torch.equal(ip, ip3)
    f = dgs_synth.feature_rows(torch.tensor([7, 3, 7]), 100)
    assert f.shape == (3, 100) and torch.equal(f[0], f[2]) and torch.isfinite(f).all()
    assert torch.equal(dgs_synth.make_features(50, 16)[7], dgs_synth.feature_rows(torch.tensor([7]), 16)[0])
    b = dgs_synth.feature_rows(torch.arange(4), 256, torch.bfloat16)
    assert b.dtype == torch.bfloat16 and torch.isfinite(b.float()).all()
    s = dgs_synth.seed_batches(1000, 64, 5)
    assert s.shape == (5, 64) and len(torch.unique(s)) == 320


def test_products_shape_statistics():
    N, E, D, dt = dgs_synth.SHAPES["products"]
    deg = dgs_synth.degrees(N, E)
    assert abs(float(deg.sum()) / E - 1) < 0.05
    assert 10_000 < int(deg.max()) < 40_000


@pytest.mark.parametrize("world", [2, 3, 8])
def test_modulo_shards_tile_the_graph(world):
    N, E = 3000, 70000
    ip, ix, pr = dgs_synth.make_csr(N, E, seed=5, weights=True)
    seen = torch.zeros(N, dtype=torch.bool)
    edges = 0
    for r in range(world):
        nids, sp, si, spr = dgs_synth.make_shard(N, E, r, world, seed=5, weights=True, node_chunk=400)
        assert torch.equal(owner_of(nids, world), torch.full_like(nids, r))
        assert torch.equal(nids, torch.arange(r, N, world))
        seen[nids] = True
        edges += si.numel()
        for j in (0, len(nids) // 2, len(nids) - 1):
            n = int(nids[j])
            assert torch.equal(si[sp[j]:sp[j + 1]], ix[ip[n]:ip[n + 1]])
            assert torch.equal(spr[sp[j]:sp[j + 1]], pr[ip[n]:ip[n + 1]])
    assert seen.all() and edges == ix.numel()


def test_partition_seeds_covers_everything():
    seeds = torch.arange(103)
    parts = [partition_seeds(seeds, r, 4) for r in range(4)]
    assert torch.equal(torch.cat(parts), seeds)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 3


# ------------------------------------------------------------------ world_size 2 over gloo
class _StubOps:
    def __init__(self, rank):
        self.rank = rank
        self.calls = []

    def _CAPI_get_unique_id(self):
        self.calls.append("get")
        return [1000 * (self.rank + 1) + i for i in range(16)]

    def _CAPI_set_nccl(self, nranks, ids, rank):
        self.calls.append(("set", nranks, list(ids), rank))


class _StubCapi:
    def __init__(self, rank):
        self.ops = _StubOps(rank)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _route_torch(nids, world):
    """What dgs_route_ids computes, restated with torch ops (test double for the gloo run)."""
    owner = nids % world
    order = torch.argsort(owner, stable=True)
    inv = torch.empty_like(nids)
    inv[order] = torch.arange(nids.numel(), dtype=nids.dtype)
    return (nids // world)[order], inv, torch.bincount(owner, minlength=world)


def _gather_torch(table, idx):
    return table[idx.long()]


def _worker(rank, world, port, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # (1) bootstrap: the id made by rank 0 reaches every rank, set_nccl gets (world, id, rank)
        capi = _StubCapi(rank)
        ids = create_communicator(world, capi=capi)
        assert ids == [1000 + i for i in range(16)]
        assert capi.ops.calls[-1] == ("set", world, ids, rank)
        assert ("get" in capi.ops.calls) == (rank == 0)
        with pytest.raises(TypeError):
            create_communicator(world, rank, capi=capi)   # the reference's tests pass the rank as group
        # (2) two groups of one: every rank is its own root
        capi2 = _StubCapi(rank)
        sub = [dist.new_group([r]) for r in range(world)]
        ids2 = create_communicator(1, sub[rank], capi=capi2)
        assert ids2 == [1000 * (rank + 1) + i for i in range(16)]
        assert capi2.ops.calls[-1] == ("set", 1, ids2, 0)
        # (3) bench.py's aggregation: time = max over ranks, work = sum over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([100.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        assert float(t) == 10.0 + world - 1 and float(w) == 100.0 * world * (world + 1) / 2
        # (4) sharding: the ranks' modulo shards are disjoint and cover the graph; seeds are split
        N, E = 2000, 40000
        nids, sp, si, _ = dgs_synth.make_shard(N, E, rank, world, seed=9)
        counts = [None] * world
        dist.all_gather_object(counts, (nids.numel(), si.numel(), int(nids.sum())))
        ip, ix, _ = dgs_synth.make_csr(N, E, seed=9)
        assert sum(c[0] for c in counts) == N and sum(c[1] for c in counts) == ix.numel()
        assert sum(c[2] for c in counts) == N * (N - 1) // 2
        mine = partition_seeds(torch.arange(101), rank, world)
        sizes = [None] * world
        dist.all_gather_object(sizes, mine.tolist())
        assert sorted(sum(sizes, [])) == list(range(101))
        # (5) id-exchange extract (north star (4)): ids to the owners, rows back, request order
        # restored - the collective logic with torch restatements of the two native kernels
        # (dgs.ops.route_ids / _CAPI_cuda_index_select), on this rank's modulo shard
        D = 5
        local_rows = dgs_synth.make_features(N, D, nids=nids)
        g = torch.Generator().manual_seed(100 + rank)
        for n_req in (0, 1, 257 + 13 * rank, 0 if rank == 0 else 50, 64 if rank == 0 else 0):
            req = torch.randint(0, N, (n_req,), generator=g)
            got = exchange_extract(req, world, rank, local_rows, route=_route_torch, gather=_gather_torch)
            assert got.shape == (n_req, D) and torch.equal(got, dgs_synth.feature_rows(req, D, torch.float32))
        with pytest.raises(RuntimeError):
            exchange_extract(req, world + 1, rank, local_rows, route=_route_torch, gather=_gather_torch)
        out.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
