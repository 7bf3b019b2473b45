"""``DistGNN.dataloading``: seed batching, on-disk graph loading and DGL-free message-flow blocks.

`SeedGenerator` and `load_dataset` keep the reference's names (python/DistGNN/dataloading);
`build_blocks` / `Block` replace the `dgl.create_block` step of the reference's training script
(example/graphsage/node_classification.py:18-28) with a CSC built by one sm_100a kernel."""
from .blocks import NID, Block, build_blocks
from .dataloader import SeedGenerator
from .load_dataset import load_dataset
from . import dataset_preprocess

__all__ = ["NID", "Block", "build_blocks", "SeedGenerator", "load_dataset", "dataset_preprocess"]
