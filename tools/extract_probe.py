"""Feature-extract kernel variants on one GPU: GB/s (algorithmic: rows * (2 * row_bytes + 8)) for
row sizes 400 B (products) and 512 B (papers100M / friendster), random and sequential row ids.

  python tools/extract_probe.py [--algos 1,2,3,4,5,6] [--rows 192000,1000000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dist-gnn_b200"))
import torch  # noqa: E402

import dgs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algos", default="1,2,3,7")
    ap.add_argument("--rows", default="192000,1000000")
    ap.add_argument("--table-rows", type=int, default=2_449_029)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--l2gran", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (32/64/128); 0 = leave")
    ap.add_argument("--tile-rows", default="0", help="algo 1: rows per CTA tile to sweep (0 = default 64)")
    ap.add_argument("--ctas", default="8", help="resident CTAs per SM the gather grids are sized for, to sweep")
    ap.add_argument("--dims", default="100,128")
    ap.add_argument("--patterns", default="random,sequential")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if args.l2gran:
        from dgs import _lib
        _lib.check(_lib.lib().dgs_set_l2_fetch_granularity(args.l2gran))
    peak = 6535.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from dgs import _lib as _l
    for dim in [int(d) for d in args.dims.split(",")]:
        table = torch.randn(args.table_rows, dim, device=dev)
        rb = dim * 4
        for R in [int(x) for x in args.rows.split(",")]:
            g = torch.Generator().manual_seed(R)
            ids = {"random": [torch.randint(0, args.table_rows, (R,), generator=g).to(dev) for _ in range(4)],
                   "sequential": [torch.arange(i * 1000, i * 1000 + R, device=dev) for i in range(4)]}
            knobs = [(int(t), int(c)) for t in args.tile_rows.split(",") for c in args.ctas.split(",")]
            for pat, qs, algo, (tile, ctas) in [(p_, q_, a_, kn) for p_, q_ in ids.items()
                                                if p_ in args.patterns.split(",")
                                                for a_ in [int(a) for a in args.algos.split(",")]
                                                for kn in (knobs if a_ == 1 else knobs[:1])]:
                _l.check(_l.lib().dgs_set_gather_tile_rows(tile))
                _l.check(_l.lib().dgs_set_gather_ctas_per_sm(ctas))
                if True:
                    outs = [dgs.ops._CAPI_cuda_index_select(table, qs[i % 4], algo) for i in range(args.reps)]
                    del outs
                    torch.cuda.synchronize()
                    e0.record()
                    outs = [dgs.ops._CAPI_cuda_index_select(table, qs[i % 4], algo) for i in range(args.reps)]
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.reps
                    gbs = R * (2 * rb + 8) / (ms * 1e-3) / 1e9
                    assert torch.equal(outs[0], table[qs[0]])
                    del outs
                    r = {"row_bytes": rb, "rows": R, "ids": pat, "algo": algo, "tile_rows": tile, "ctas_per_sm": ctas,
                         "us": ms * 1e3, "gbs": gbs, "frac_of_measured_peak": gbs / peak}
                    res.append(r)
                    print(f"{rb} B x {R:8d} {pat:10s} algo {algo} tile {tile:2d} ctas {ctas}: {ms * 1e3:8.1f} us {gbs:7.0f} GB/s "
                          f"{gbs / peak:.3f}", file=sys.stderr, flush=True)
        del table
    print(json.dumps(res))


if __name__ == "__main__":
    main()
