"""Dataset IO: OGB raw files -> the reference's *.pt directory -> load_dataset.

The expected CSC is computed the way the reference computes it
(python/DistGNN/dataloading/dataset_preprocess.py:34-43: scipy `coo_matrix((zeros, (dst, src))).tocsr()`),
so these tests pin `build_csc` to the reference's algorithm bit for bit on the CPU."""
import gzip
import os

import numpy as np
import pytest
import torch
from scipy.sparse import coo_matrix

from DistGNN.dataloading import load_dataset
from DistGNN.dataloading import dataset_preprocess as dp


def _write_csv_gz(path, arr):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    arr = np.asarray(arr)
    if arr.ndim == 1:
        arr = arr[:, None]
    with gzip.open(path, "wt") as f:
        for row in arr:
            f.write(",".join(repr(x) if isinstance(x, float) else str(x) for x in row.tolist()) + "\n")


def _scipy_csc(src, dst, n):
    coo = coo_matrix((np.zeros(len(src)), (dst, src)), shape=(n, n), dtype=np.int64)
    csr = coo.tocsr()
    return csr.indptr.astype(np.int64), csr.indices.astype(np.int64)


def _raw_products(root, n=60, m=400, d=7, seed=0):
    rng = np.random.default_rng(seed)
    edges = rng.integers(0, n, (m, 2))
    edges[:20] = edges[20:40]            # duplicate edges
    edges[40:45, 1] = edges[40:45, 0]    # self loops
    feats = rng.standard_normal((n, d)).astype(np.float32)
    labels = rng.integers(0, 5, n)
    perm = rng.permutation(n)
    _write_csv_gz(os.path.join(root, "raw/edge.csv.gz"), edges)
    _write_csv_gz(os.path.join(root, "raw/node-feat.csv.gz"), feats.astype(np.float64))
    _write_csv_gz(os.path.join(root, "raw/node-label.csv.gz"), labels)
    for k, sl in (("train", perm[:30]), ("valid", perm[30:45]), ("test", perm[45:])):
        _write_csv_gz(os.path.join(root, f"split/sales_ranking/{k}.csv.gz"), sl)
    return edges, feats, labels, perm


def test_build_csc_equals_scipy_tocsr():
    rng = np.random.default_rng(1)
    for n, m in ((1, 0), (5, 3), (50, 600), (300, 5000)):
        src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
        ip, ix = dp.build_csc(src, dst, n)
        eip, eix = _scipy_csc(src, dst, n)
        assert ip.dtype == torch.int64 and ix.dtype == torch.int64
        assert np.array_equal(ip.numpy(), eip) and np.array_equal(ix.numpy(), eix)
    with pytest.raises(ValueError):
        dp.build_csc([0, 7], [1, 2], 5)


def test_process_products_roundtrip(tmp_path):
    root, out = str(tmp_path / "raw_products"), str(tmp_path / "out")
    edges, feats, labels, perm = _raw_products(root)
    meta = dp.process_products(root, out, bias=True)
    n = feats.shape[0]
    src = np.concatenate((edges[:, 0], edges[:, 1]))
    dst = np.concatenate((edges[:, 1], edges[:, 0]))      # symmetrised, dataset_preprocess.py:34-36
    eip, eix = _scipy_csc(src, dst, n)
    graph, num_classes = load_dataset(out, "ogbn-products", with_feature=True, with_probs=True)
    assert set(graph) == {"labels", "indptr", "indices", "train_idx", "features", "probs"}
    assert np.array_equal(graph["indptr"].numpy(), eip) and np.array_equal(graph["indices"].numpy(), eix)
    assert graph["indptr"].dtype == torch.int64 and graph["indices"].dtype == torch.int64
    assert graph["features"].dtype == torch.float32 and torch.equal(graph["features"], torch.from_numpy(feats))
    assert graph["labels"].dtype == torch.int64 and graph["labels"].tolist() == labels.tolist()
    assert graph["train_idx"].tolist() == perm[:30].tolist()
    assert graph["probs"].dtype == torch.float32 and graph["probs"].numel() == len(eix)
    assert bool((graph["probs"] >= 0).all())
    assert num_classes == len(np.unique(labels)) == meta["num_classes"]
    assert meta == {"dataset": "ogbn-products", "num_nodes": n, "num_edges": len(eix),
                    "num_classes": num_classes, "feature_dim": feats.shape[1], "num_train_nodes": 30,
                    "num_valid_nodes": 15, "num_test_nodes": n - 45}
    assert torch.load(os.path.join(out, "valid_idx.pt")).tolist() == perm[30:45].tolist()
    # the loader refuses a directory that holds another dataset (reference: assert, load_dataset.py:9)
    with pytest.raises(RuntimeError):
        load_dataset(out, "ogbn-papers100M")
    g2, _ = load_dataset(out, "ogbn-products", with_feature=False)
    assert "features" not in g2 and "probs" not in g2


def _raw_papers(root, n=80, m=500, d=6, seed=3):
    rng = np.random.default_rng(seed)
    edge_index = rng.integers(0, n, (2, m))
    feats = rng.standard_normal((n, d)).astype(np.float32)
    labels = rng.integers(0, 4, (n, 1)).astype(np.float32)
    labels[rng.random(n) < 0.5] = np.nan          # most papers are unlabeled
    os.makedirs(os.path.join(root, "raw"), exist_ok=True)
    np.savez(os.path.join(root, "raw/data.npz"), node_feat=feats, edge_index=edge_index)
    np.savez(os.path.join(root, "raw/node-label.npz"), node_label=labels)
    perm = rng.permutation(n)
    a, b = n // 2, 3 * n // 4
    for k, sl in (("train", perm[:a]), ("valid", perm[a:b]), ("test", perm[b:])):
        _write_csv_gz(os.path.join(root, f"split/time/{k}.csv.gz"), sl)
    return edge_index, feats, labels, perm


def test_process_papers100M_roundtrip(tmp_path):
    root, out = str(tmp_path / "raw_papers"), str(tmp_path / "out")
    edge_index, feats, labels, perm = _raw_papers(root)
    meta = dp.process_papers100M(root, out)
    n = feats.shape[0]
    eip, eix = _scipy_csc(edge_index[0], edge_index[1], n)    # directed, dataset_preprocess.py:118-119
    graph, num_classes = load_dataset(out, "ogbn-papers100M")
    assert np.array_equal(graph["indptr"].numpy(), eip) and np.array_equal(graph["indices"].numpy(), eix)
    assert graph["labels"].dtype == torch.float32 and graph["labels"].shape == (n,)
    lab = labels[:, 0]
    assert np.array_equal(np.isnan(graph["labels"].numpy()), np.isnan(lab))
    assert num_classes == len(np.unique(lab[~np.isnan(lab)])) == meta["num_classes"]
    assert not os.path.exists(os.path.join(out, "probs.pt"))
    assert meta["num_edges"] == len(eix) and meta["feature_dim"] == feats.shape[1]


def test_generate_papers400M_structure(tmp_path):
    root, out = str(tmp_path / "raw_papers"), str(tmp_path / "out400")
    edge_index, feats, labels, perm = _raw_papers(root, n=40, m=200)
    meta = dp.generate_papers400M(root, out, seed=5)
    n = feats.shape[0]
    graph, _ = load_dataset(out, "ogbn-papers400M")
    ip, ix = graph["indptr"].numpy(), graph["indices"].numpy()
    assert meta["num_nodes"] == 4 * n == len(ip) - 1 and graph["features"].shape == (4 * n, feats.shape[1])
    assert torch.equal(graph["features"][n:2 * n], torch.from_numpy(feats))
    assert graph["train_idx"].tolist() == np.concatenate([perm[:20] + c * n for c in range(4)]).tolist()
    dst = np.repeat(np.arange(4 * n), np.diff(ip))
    have = set(zip(ix.tolist(), dst.tolist()))                 # (src, dst)
    # the copy-to-copy links, paired like the reference pairs them (dataset_preprocess.py:190-208)
    ids = np.arange(n)
    for c in range(4):
        others = np.concatenate([ids + o * n for o in range(4) if o != c])
        for s_, d_ in zip(np.repeat(ids + c * n, 3).tolist(), others.tolist()):
            assert (s_, d_) in have
    # every original edge, in both directions, between SOME pair of copies; nothing else
    mod = {(s_ % n, d_ % n) for s_, d_ in have}
    orig = set(zip(edge_index[0].tolist(), edge_index[1].tolist()))
    links = {(s_ % n, d_ % n) for c in range(4) for s_, d_ in
             zip(np.repeat(ids + c * n, 3).tolist(),
                 np.concatenate([ids + o * n for o in range(4) if o != c]).tolist())}
    assert orig <= mod and {(b, a) for a, b in orig} <= mod
    assert mod <= orig | {(b, a) for a, b in orig} | links
    # rows sorted, duplicates merged
    for r in range(4 * n):
        row = ix[ip[r]:ip[r + 1]]
        assert np.all(np.diff(row) > 0)
