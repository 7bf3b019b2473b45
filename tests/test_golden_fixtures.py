"""CPU: the oracle against fixtures written by the reference's own kernels on the GPU box
(tests/golden/ref_*.npz, produced by `DGS_WRITE_GOLDEN=1 pytest tests/test_ref_differential.py`
with oracle/_ref present, see tests/golden/README.md).  Skipped until the fixtures exist."""
import glob
import os

import numpy as np
import pytest
import torch

import dgs_synth
import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    p = os.path.join(G, name + ".npz")
    if not os.path.exists(p):
        pytest.skip(f"{name}.npz not generated yet")
    return np.load(p)


def test_golden_index_select():
    z = _load("ref_index_select")
    feat = dgs_synth.feature_rows(torch.arange(5000), 100).numpy()
    assert np.array_equal(oracle.index_select(feat, z["nids"]), z["out"])


def test_golden_extract_subcsr():
    z = _load("ref_extract_subcsr")
    indptr, _, _ = dgs_synth.make_csr(4000, 90000, seed=21, weights=True)
    assert np.array_equal(oracle.extract_indptr(z["nids"], indptr.numpy()), z["sub_indptr"])


def test_golden_relabel():
    z = _load("ref_relabel")
    u, _ = oracle.relabel([z["seeds"], z["col"]], [z["row"]])
    assert np.array_equal(u, z["unique"])


def test_golden_full_neighbor():
    z = _load("ref_full_neighbor")
    indptr, indices, _ = dgs_synth.make_csr(4000, 90000, seed=21)
    r, c = oracle.sample_all_neighbors(z["seeds"], indptr.numpy(), indices.numpy())
    assert len(r) == int(z["nnz"][0])
    assert np.array_equal(r[:4000], z["row"]) and np.array_equal(c[:4000], z["col"])


def test_golden_feature_server():
    z = _load("ref_feature_server")
    feat = dgs_synth.feature_rows(torch.arange(6000), 100).numpy()
    key, idx, dev = oracle.hashmap_build([z["cache"]], 0)
    out = oracle.extract_p2p(feat, [feat[z["cache"]]], key, idx, dev, z["q"])
    assert np.array_equal(out, z["out"])


def test_golden_sampler_blocks():
    z = _load("ref_sampler_blocks")
    indptr, indices, _ = dgs_synth.make_csr(1200, 6000, seed=23)
    out = oracle.sample_blocks_all_neighbors(z["seeds"], indptr.numpy(), indices.numpy(), 2)
    assert np.array_equal(out[0][1], z["frontier0"])
    assert np.array_equal(out[0][2], z["row0"]) and np.array_equal(out[0][3], z["col0"])
    assert np.array_equal(out[1][1][:5000], z["frontier1"]) and len(out[1][2]) == int(z["nnz1"][0])
