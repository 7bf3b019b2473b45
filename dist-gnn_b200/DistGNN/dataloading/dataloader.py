"""Seed batch iterator with the interface of the reference's SeedGenerator
(python/DistGNN/dataloading/dataloader.py:4-45): iterate -> consecutive batches of the seed tensor,
re-permuted on its own device at the start of every epoch when shuffle=True; `len()` = batches per
epoch; `is_finished()`; the public counters `step` / `last_step`.

Batches are `narrow()` views of one tensor (no copy); unlike the reference `len()` and
`is_finished()` also work before the first epoch has started."""
import torch


def _batches_per_epoch(num_items, batch_size, drop_last):
    full, rest = divmod(int(num_items), int(batch_size))
    return full if (drop_last or rest == 0) else full + 1


class SeedGenerator(object):

    def __init__(self, data: torch.Tensor, batch_size: int, shuffle: bool = False,
                 drop_last: bool = False):
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        self.data, self.batch_size = data, int(batch_size)
        self.shuffle, self.drop_last = bool(shuffle), bool(drop_last)
        self.last_step = _batches_per_epoch(data.shape[0], self.batch_size, self.drop_last)
        self.step = self.last_step          # nothing to hand out until an epoch is started

    def __iter__(self):
        n = self.data.shape[0]
        if self.shuffle and n > 1:
            self.data = self.data.index_select(0, torch.randperm(n, device=self.data.device))
        self.last_step = _batches_per_epoch(n, self.batch_size, self.drop_last)
        self.step = 0
        return self

    def __next__(self):
        if self.is_finished():
            raise StopIteration
        start = self.step * self.batch_size
        self.step += 1
        return self.data.narrow(0, start, min(self.batch_size, self.data.shape[0] - start))

    def __len__(self):
        return self.last_step

    def is_finished(self):
        return self.step >= self.last_step
