"""OGB raw data -> the reference's on-disk layout (a directory of *.pt tensors + metadata.pt), the
counterpart of python/DistGNN/dataloading/dataset_preprocess.py:9-324.

Same inputs (the OGB `raw/` and `split/` files), same outputs (file names, dtypes, metadata keys),
same graph: the reference builds `coo_matrix((zeros, (dst, src))).tocsr()`, i.e. a CSC whose row n
holds the DISTINCT in-neighbours of n in ascending order (scipy merges duplicate entries);
ogbn-products is symmetrised first (:34-36), ogbn-papers100M is taken as directed (:118-119).
Here the CSC is built with a sort + unique + bincount on whatever device the edge list fits
(`build_csc(..., device="cuda")` on a B200 handles the 3.2 G symmetrised papers100M pairs in HBM
instead of scipy's single-threaded host pass), the result is bit-identical.

  python -m DistGNN.dataloading.dataset_preprocess --dataset ogbn-products --root raw_dir --save-path out [--bias]
"""
import argparse
import os

import numpy as np
import torch

DATASETS = ("ogbn-products", "ogbn-papers100M", "ogbn-papers400M")


def build_csc(src, dst, num_nodes, device=None, chunk=1 << 28):
    """CSC (indptr int64[N+1], indices int64[E']) of the distinct edges src -> dst: row n lists the
    distinct sources of edges into n, ascending - what `coo_matrix((data, (dst, src))).tocsr()` of
    the reference yields (dataset_preprocess.py:37-43)."""
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    if src.numel() != dst.numel():
        raise ValueError("src and dst must have the same length")
    if src.numel() and (int(torch.min(src)) < 0 or int(torch.min(dst)) < 0 or
                        int(torch.max(src)) >= num_nodes or int(torch.max(dst)) >= num_nodes):
        raise ValueError("edge endpoint outside [0, num_nodes)")
    if device is None:
        device = src.device
    # one 64-bit key per edge, sorted: (dst, src) order = row-major CSC with ascending neighbours
    key = dst.to(device) * num_nodes + src.to(device)
    key = torch.unique(key)          # sorted + duplicates merged
    rows = torch.div(key, num_nodes, rounding_mode="floor")
    indices = key - rows * num_nodes
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=key.device)
    torch.cumsum(torch.bincount(rows, minlength=num_nodes), 0, out=indptr[1:])
    return indptr.cpu(), indices.cpu()


def make_edge_probs(num_edges, generator=None):
    """|N(0, 1)| float32 edge weights, as the reference draws them (dataset_preprocess.py:69-71)."""
    return torch.randn((num_edges,), generator=generator).abs().float()


def save_dataset(save_path, name, indptr, indices, features, labels, train_idx, valid_idx, test_idx,
                 bias=False, generator=None):
    """Write the reference's file set: features / labels / indptr / indices / {train,valid,test}_idx
    (+ probs with bias) and metadata.pt with the reference's keys (dataset_preprocess.py:46-91)."""
    os.makedirs(save_path, exist_ok=True)
    features = torch.as_tensor(features).float()
    labels_t = torch.as_tensor(labels)
    tensors = {
        "features": features,
        "labels": labels_t,
        "indptr": torch.as_tensor(indptr).long(),
        "indices": torch.as_tensor(indices).long(),
        "train_idx": torch.as_tensor(train_idx).long(),
        "valid_idx": torch.as_tensor(valid_idx).long(),
        "test_idx": torch.as_tensor(test_idx).long(),
    }
    if bias:
        tensors["probs"] = make_edge_probs(tensors["indices"].numel(), generator)
    for k, t in tensors.items():
        torch.save(t, os.path.join(save_path, k + ".pt"))
    lab = labels_t.double().numpy().reshape(-1)
    meta = {
        "dataset": name,
        "num_nodes": int(features.shape[0]),
        "num_edges": int(tensors["indices"].numel()),
        "num_classes": int(np.unique(lab[~np.isnan(lab)]).shape[0]),
        "feature_dim": int(features.shape[1]),
        "num_train_nodes": int(tensors["train_idx"].numel()),
        "num_valid_nodes": int(tensors["valid_idx"].numel()),
        "num_test_nodes": int(tensors["test_idx"].numel()),
    }
    torch.save(meta, os.path.join(save_path, "metadata.pt"))
    return meta


def _read_csv_gz(path):
    import pandas as pd
    return pd.read_csv(path, compression="gzip", header=None).values


def process_products(dataset_path, save_path, bias=False, device=None):
    """ogbn-products: undirected edge list -> symmetrised CSC; labels int64
    (dataset_preprocess.py:9-91)."""
    edges = _read_csv_gz(os.path.join(dataset_path, "raw/edge.csv.gz")).T
    features = _read_csv_gz(os.path.join(dataset_path, "raw/node-feat.csv.gz"))
    labels = _read_csv_gz(os.path.join(dataset_path, "raw/node-label.csv.gz")).T[0]
    split = {k: _read_csv_gz(os.path.join(dataset_path, f"split/sales_ranking/{k}.csv.gz")).T[0]
             for k in ("train", "valid", "test")}
    src = np.concatenate((edges[0], edges[1]))
    dst = np.concatenate((edges[1], edges[0]))
    indptr, indices = build_csc(src, dst, features.shape[0], device)
    return save_dataset(save_path, "ogbn-products", indptr, indices, features,
                        torch.from_numpy(np.asarray(labels)).long(), split["train"], split["valid"],
                        split["test"], bias)


def process_papers100M(dataset_path, save_path, bias=False, device=None):
    """ogbn-papers100M: directed edge list as given; labels float32 with NaN for unlabeled nodes
    (dataset_preprocess.py:94-173)."""
    data_file = np.load(os.path.join(dataset_path, "raw/data.npz"))
    label_file = np.load(os.path.join(dataset_path, "raw/node-label.npz"))
    features, edge_index = data_file["node_feat"], data_file["edge_index"]
    labels = torch.from_numpy(np.array(label_file["node_label"])).float().squeeze(1)
    split = {k: _read_csv_gz(os.path.join(dataset_path, f"split/time/{k}.csv.gz")).T[0]
             for k in ("train", "valid", "test")}
    indptr, indices = build_csc(edge_index[0], edge_index[1], features.shape[0], device)
    return save_dataset(save_path, "ogbn-papers100M", indptr, indices, features, labels, split["train"],
                        split["valid"], split["test"], bias)


def generate_papers400M(papers100M_path, save_path, bias=False, device=None, seed=None):
    """Four copies of ogbn-papers100M: every original edge (both directions) lands between random
    copies of its endpoints, plus 3 extra out-edges per node into the other three copies
    (dataset_preprocess.py:176-324).  The extra edges are paired exactly like the reference pairs
    them - source list = every id of copy c repeated three times, destination list = the ids of the
    other copies in ascending order - which links node i of copy c to nodes 3 i, 3 i + 1, 3 i + 2 of
    that list, not to its own twins (kept: the point is the same edge multiset, not a nicer graph).
    Features / labels / splits are the originals repeated."""
    data_file = np.load(os.path.join(papers100M_path, "raw/data.npz"))
    label_file = np.load(os.path.join(papers100M_path, "raw/node-label.npz"))
    feats, edge_index = data_file["node_feat"], data_file["edge_index"]
    osrc, odst = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    n, m = feats.shape[0], osrc.shape[0]
    rng = np.random.default_rng(seed)
    sm = rng.integers(0, 4, 2 * m, dtype=np.int64)
    dm = rng.integers(0, 4, 2 * m, dtype=np.int64)
    ids = np.arange(n, dtype=np.int64)
    twin_src = np.concatenate([np.repeat(ids + c * n, 3) for c in range(4)])
    twin_dst = np.concatenate([np.concatenate([ids + o * n for o in range(4) if o != c])
                               for c in range(4)])
    src = np.concatenate([osrc + sm[:m] * n, odst + sm[m:] * n, twin_src])
    dst = np.concatenate([odst + dm[:m] * n, osrc + dm[m:] * n, twin_dst])
    indptr, indices = build_csc(src, dst, 4 * n, device)
    split = {k: _read_csv_gz(os.path.join(papers100M_path, f"split/time/{k}.csv.gz")).T[0]
             for k in ("train", "valid", "test")}
    rep = {k: np.concatenate([v + c * n for c in range(4)]) for k, v in split.items()}
    labels = torch.from_numpy(np.concatenate([label_file["node_label"]] * 4)).float().squeeze(1)
    return save_dataset(save_path, "ogbn-papers400M", indptr, indices, np.concatenate([feats] * 4, 0),
                        labels, rep["train"], rep["valid"], rep["test"], bias)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="ogbn-papers100M", choices=DATASETS)
    ap.add_argument("--root", default="dataset/")
    ap.add_argument("--save-path", default=".")
    ap.add_argument("--bias", action="store_true", default=False)
    ap.add_argument("--device", default=None, help="where the CSC is built (e.g. cuda); default: host")
    args = ap.parse_args(argv)
    fn = {"ogbn-products": process_products, "ogbn-papers100M": process_papers100M,
          "ogbn-papers400M": generate_papers400M}[args.dataset]
    meta = fn(args.root, args.save_path, args.bias, device=args.device)
    print(meta)


if __name__ == "__main__":
    main()
