"""Seed batch iterator (python/DistGNN/dataloading/dataloader.py:4-45)."""
import torch


class SeedGenerator(object):
    """Iterates `data` in batches of `batch_size`; with shuffle=True the data is re-permuted (on its
    own device) at the start of every epoch."""

    def __init__(self, data: torch.Tensor, batch_size: int, shuffle: bool = False,
                 drop_last: bool = False):
        self.data = data
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.step = 0
        self.last_step = 0

    def __iter__(self):
        if self.shuffle:
            perm = torch.randperm(self.data.shape[0], device=self.data.device)
            self.data = self.data[perm]
        self.step = 0
        n = self.data.shape[0]
        self.last_step = n // self.batch_size if self.drop_last else -(-n // self.batch_size)
        return self

    def __next__(self):
        if self.step >= self.last_step:
            raise StopIteration
        lo = self.step * self.batch_size
        self.step += 1
        return self.data[lo:lo + self.batch_size]

    def __len__(self):
        return self.last_step

    def is_finished(self):
        return self.step >= self.last_step
