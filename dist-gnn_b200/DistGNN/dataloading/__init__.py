from .dataloader import SeedGenerator
from .load_dataset import load_dataset
