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
        self._src = self._eid = None

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
            try:
                # the sampler emits edges grouped by destination when the hop's seeds are distinct
                self._indptr = self._capi.ops.coo_rows_to_indptr(self._row, self.num_dst_nodes(),
                                                                 check_sorted=True)
                self._src = self._col
                self._eid = torch.arange(self.num_edges(), dtype=self._col.dtype,
                                         device=self._col.device)
            except RuntimeError as e:
                if "ascending" not in str(e):
                    raise
                # duplicate seeds in the first hop: rows are first-occurrence ids, not ascending ->
                # stable sort by destination (what dgl.create_block's COO->CSC conversion does)
                order = torch.argsort(self._row, stable=True)
                deg = torch.bincount(self._row.long(), minlength=self.num_dst_nodes())
                indptr = torch.zeros(self.num_dst_nodes() + 1, dtype=self._row.dtype,
                                     device=self._row.device)
                indptr[1:] = torch.cumsum(deg, 0)
                self._indptr = indptr
                self._src = self._col[order]
                self._eid = order.to(self._col.dtype)
        return self._indptr, self._src, self._eid

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
