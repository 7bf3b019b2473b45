"""Message-flow blocks from the sampler's output, without DGL.

The reference's caller turns every sampled hop into `dgl.create_block((coo_col, coo_row),
num_src_nodes=|frontier|, num_dst_nodes=|seeds|)` and stores the global ids as srcdata / dstdata
(example/graphsage/node_classification.py:18-28).  DGL is not a dependency here; `Block` carries the
same information - edges src -> dst in local ids, `srcdata[NID]`, `dstdata[NID]` - and the CSC a
message-passing layer consumes, built by one tiny kernel because the sampler already emits the
edges grouped by destination."""
import torch

NID = "_ID"     # dgl.NID


class Block:
    """One hop: `num_src_nodes()` frontier nodes -> `num_dst_nodes()` seeds."""

    def __init__(self, seeds, frontier, coo_row, coo_col, capi):
        self._capi = capi
        self._row, self._col = coo_row, coo_col
        self.srcdata = {NID: frontier}
        self.dstdata = {NID: seeds}
        self._indptr = None

    def num_src_nodes(self):
        return self.srcdata[NID].numel()

    def num_dst_nodes(self):
        return self.dstdata[NID].numel()

    def num_edges(self):
        return self._row.numel()

    def edges(self):
        """(src, dst) local ids, i.e. the (coo_col, coo_row) handed to dgl.create_block."""
        return self._col, self._row

    def csc(self):
        """(indptr over dst nodes, src ids, edge ids) - adj_tensors('csc') of the DGL block."""
        if self._indptr is None:
            self._indptr = self._capi.ops.coo_rows_to_indptr(self._row, self.num_dst_nodes())
        return self._indptr, self._col, torch.arange(self.num_edges(), dtype=self._col.dtype,
                                                     device=self._col.device)

    def in_degrees(self):
        indptr = self.csc()[0]
        return indptr[1:] - indptr[:-1]


def build_blocks(batch, capi=None):
    """build_blocks of the reference's training script: the sampler's list of
    (seeds, frontier, coo_row, coo_col), seed-side hop first, becomes a list of blocks with the
    input-side hop first (the order the model consumes them)."""
    if capi is None:
        import dgs as capi
    blocks = []
    for seeds, frontier, coo_row, coo_col in batch:
        blocks.insert(0, Block(seeds, frontier, coo_row, coo_col, capi))
    return blocks
