"""Does cudaLimitMaxL2FetchGranularity change the random-row gather?  (products shape, 400-B rows)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch, dgs, dgs_synth
from dgs import _lib
dev = torch.device("cuda", 0)
for shape in ("products", "papers_like"):
    if shape == "products":
        N, D = 2_449_029, 100
    else:
        N, D = 20_000_000, 128
    ft = torch.empty((N, D), dtype=torch.float32, device=dev).normal_()
    qs = [torch.randint(0, N, (192_000 if shape == "products" else 1_000_000,), device=dev) for _ in range(8)]
    for gran in (64, 32, 128, 64):
        _lib.check(_lib.lib().dgs_set_l2_fetch_granularity(gran))
        for algo in (1, 2):
            outs = [dgs.ops._CAPI_cuda_index_select(ft, q, algo) for q in qs]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for rep in range(3):
                outs = [dgs.ops._CAPI_cuda_index_select(ft, q, algo) for q in qs]
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 24
            rows = qs[0].numel()
            print(json.dumps({"shape": shape, "l2_fetch_granularity": gran, "algo": algo, "us": round(ms * 1e3, 2),
                              "algorithmic_gbps": round(rows * (2 * D * 4 + 8) / ms / 1e6, 1)}), flush=True)
    del ft
    torch.cuda.empty_cache()
