"""Dataset loader for the reference's on-disk layout
(python/DistGNN/dataloading/load_dataset.py:5-32): a directory of *.pt tensors + metadata.pt."""
import os

import torch


def load_dataset(path, dataset_name, with_feature=True, with_probs=False):
    """Returns (graph_tensors: dict, num_classes).  Keys: indptr, indices, [probs], [features],
    labels, train_idx (each a CPU tensor, ready for _CAPI_tensor_pin_memory)."""
    meta = torch.load(os.path.join(path, "metadata.pt"))
    if meta["dataset"] != dataset_name:
        raise RuntimeError(f"{path} holds dataset {meta['dataset']!r}, not {dataset_name!r}")
    keys = ["indptr", "indices", "labels", "train_idx"]
    if with_probs:
        keys.append("probs")
    if with_feature:
        keys.append("features")
    graph = {k: torch.load(os.path.join(path, k + ".pt")) for k in keys}
    return graph, meta["num_classes"]
