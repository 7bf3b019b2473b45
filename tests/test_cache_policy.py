"""Cache policy against golden vectors produced by the REFERENCE's own
python/DistGNN/cache/cache_value.py (tools/make_cache_policy_golden.py): selfish / selfless
placement, the hottest-rank partition and both value estimators, on a 2-rank gloo group (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cache_policy_inputs import COST, MEMS, WORLD, cost_args, inputs
from DistGNN import cache as C

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_cache_policy.npz")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=WORLD)
    try:
        gold = np.load(GOLDEN)
        a = cost_args()
        for bias in (False, True):
            for mem in MEMS:
                graph, sh, fh, probs = inputs(rank, bias)
                sh0, fh0 = sh.clone(), fh.clone()
                tag = f"r{rank}_b{int(bias)}_m{mem}"
                s1, f1 = C.get_cache_nids_selfish(graph, sh, fh, mem, *a, probs=probs)
                v1 = C.compute_total_value_selfish(graph, sh, fh, s1, f1, *a, probs=probs)
                s2, f2 = C.get_cache_nids_selfless(graph, sh, fh, mem, *a, probs=probs)
                v2 = C.compute_total_value_selfless(graph, sh, fh, s2, f2, COST["bandwidth_gpu"],
                                                    COST["bandwidth_nvlink"], WORLD, *a[1:], probs=probs)
                hs, hf = C.get_hot_nids_p2p_global(sh, fh)
                for k, v in (("selfish_s", s1), ("selfish_f", f1), ("selfless_s", s2), ("selfless_f", f2),
                             ("global_s", hs), ("global_f", hf)):
                    assert np.array_equal(v.numpy(), gold[f"{tag}_{k}"]), (tag, k)
                assert np.array_equal(np.array([v1, v2]), gold[f"{tag}_values"]), tag
                assert torch.equal(sh, sh0) and torch.equal(fh, fh0)      # heats are left untouched
        # "auto": both ranks agree, and the answer is the argmax of the all-reduced values
        graph, sh, fh, _ = inputs(rank, False)
        name, s, f = C.choose_cache_policy(graph, sh, fh, MEMS[1], WORLD, cost_model=COST)
        tot = np.zeros(2)
        for r in range(WORLD):
            tot += gold[f"r{r}_b0_m{MEMS[1]}_values"]
        assert name == ("selfish" if tot[0] > tot[1] else "selfless")
        key = "selfish" if name == "selfish" else "selfless"
        assert np.array_equal(s.numpy(), gold[f"r{rank}_b0_m{MEMS[1]}_{key}_s"])
        assert np.array_equal(f.numpy(), gold[f"r{rank}_b0_m{MEMS[1]}_{key}_f"])
        out.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_policies_match_reference_golden_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_space_and_value_helpers():
    graph, sh, fh, _ = inputs(0, True)
    nids = torch.tensor([0, 5, 17, 2999])
    deg = graph["indptr"][nids + 1] - graph["indptr"][nids]
    assert torch.equal(C.get_structure_space(nids, graph), deg * 8 + 8)
    assert torch.equal(C.get_structure_space(nids, graph, probs="probs"), deg * 12 + 8)
    assert C.get_feature_space(graph) == 24 * 4
    assert torch.equal(C.get_node_value(torch.tensor([2.0, 4.0]), 4, 3.0), torch.tensor([1.5, 3.0]))
    with pytest.raises(AssertionError):
        C.get_node_value(sh, 1.5, 1.0)
    s, f = C.get_hot_nids_local(sh, fh)
    assert torch.equal(s, torch.nonzero(sh).flatten()) and torch.equal(f, torch.nonzero(fh).flatten())
    # knapsack: descending value, cut where the prefix sum reaches the capacity
    sn, fn = torch.tensor([10, 11, 12]), torch.tensor([20, 21])
    ss, fs = torch.tensor([100, 100, 100]), torch.tensor([50, 50])
    sv, fv = torch.tensor([5.0, 1.0, 3.0]), torch.tensor([4.0, 2.0])
    cs, cf, used = C.get_cache_nids_local(sn, ss, sv, fn, fs, fv, 260)
    assert cs.tolist() == [10, 12] and cf.tolist() == [20] and used == 250
