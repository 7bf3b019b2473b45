"""NVLink evidence for the peer-load gather (2 GPUs, ONE process, so it can run under ncu):
the feature table lives on cuda:1, the gather kernel runs on cuda:0 and reads it through a plain
peer mapping (cudaDeviceEnablePeerAccess) - the same in-kernel ld.global peer loads the sharded
feature server issues on VMM-mapped shards.

  python tools/nvlink_probe.py                       # GB/s per row size / algo, expected link bytes
  ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,... python tools/nvlink_probe.py --once --dim 100

Expected received bytes per launch: R * row_bytes of payload; what the link really moves is whole
128-byte lines, i.e. R * 128 * ceil((offset % 128 + row_bytes) / 128) - 512 bytes for a 400-byte
row at any 16-byte-aligned offset."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402

import dgs  # noqa: E402
from dgs import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=0, help="feature dim (fp32); 0 = sweep 100 and 128")
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--table-rows", type=int, default=4_000_000)
    ap.add_argument("--algos", default="1,2,3,7")
    ap.add_argument("--once", action="store_true", help="one launch per (dim, algo): for ncu")
    args = ap.parse_args()
    d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
    torch.cuda.set_device(0)
    _lib.check(_lib.lib().dgs_enable_peer_access(1), "peer access")
    out = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for dim in ([args.dim] if args.dim else [100, 128]):
        table = torch.randn(args.table_rows, dim, device=d1)
        rb = dim * 4
        g = torch.Generator().manual_seed(dim)
        nids = torch.randint(0, args.table_rows, (args.rows,), generator=g).to(d0)
        off = (nids.cpu() * rb) % 128
        lines = ((off + rb + 127) // 128).sum().item() * 128
        for algo in [int(a) for a in args.algos.split(",")]:
            reps = 1 if args.once else 10
            x = dgs.ops._CAPI_cuda_index_select(table, nids, algo)
            torch.cuda.synchronize()
            if not args.once:
                assert torch.equal(x.cpu(), table[nids.to(d1)].cpu())
            e0.record()
            for _ in range(reps):
                x = dgs.ops._CAPI_cuda_index_select(table, nids, algo)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            r = {"row_bytes": rb, "rows": args.rows, "algo": algo, "ms": ms,
                 "payload_bytes": args.rows * rb, "line_bytes_128": lines,
                 "payload_gbps": args.rows * rb / ms / 1e6, "line_gbps": lines / ms / 1e6}
            out.append(r)
            print(json.dumps(r), file=sys.stderr, flush=True)
        del table
    print(json.dumps(out))


if __name__ == "__main__":
    main()
