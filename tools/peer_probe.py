"""Peer-gather probe (2 GPUs, ONE process): random row gather on cuda:0 out of a table that lives on
cuda:1 (plain peer mapping via cudaDeviceEnablePeerAccess), for growing table sizes - separates the
cost of NVLink random reads from the cost of CUDA-IPC mappings."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch, dgs

d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
torch.cuda.set_device(0)
from dgs import _lib
_lib.check(_lib.lib().dgs_enable_peer_access(1), 'peer access')
D = 128
for gb in (0.5, 2, 8, 28):
    rows = int(gb * 1e9 / (D * 4))
    for where in ("peer", "local"):
        table = torch.empty((rows, D), dtype=torch.float32, device=d1 if where == "peer" else d0)
        R = 1_000_000
        nids = torch.randint(0, rows, (R,), device=d0)
        for algo in (1, 2):
            outs = [dgs.ops._CAPI_cuda_index_select(table, nids, algo) for _ in range(3)]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs = [dgs.ops._CAPI_cuda_index_select(table, nids, algo) for _ in range(3)]
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(json.dumps({"table_gb": gb, "where": where, "algo": algo, "ms": round(ms, 4),
                              "row_read_gbps": round(R * D * 4 / ms / 1e6, 1)}), flush=True)
            del outs
        del table
        torch.cuda.empty_cache()
