from .dataloader import SeedGenerator
from .load_dataset import load_dataset
from .blocks import NID, Block, build_blocks
