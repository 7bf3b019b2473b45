"""Turn an `ncu --set full` report into the small JSON summary committed under profiles/.

  python tools/ncu_summary.py gpurun_out/r02_full.ncu-rep|raw.csv profiles/r02_ncu_full_summary.json \
      fused_batch=multi_batch_kernel gather_rows=gather_rows_kernel mb_pick=mb_pick_kernel ...

Each NAME=REGEX selects the launches whose kernel name matches REGEX; the summary keeps, per launch,
the metrics that DESIGN.md / bench.py quote (time, DRAM bytes, throughput %, occupancy, issue
activity, stall reasons, sectors)."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
    "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__inst_executed.sum",
    "nvlrx__bytes.sum", "nvltx__bytes.sum", "lts__t_sectors_srcunit_tex_aperture_peer.sum",
    "lts__t_sectors_srcunit_tex_aperture_peer_op_read.sum",
    "lts__t_sectors_srcunit_tex_aperture_sysmem.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    groups = [a.split("=", 1) for a in sys.argv[3:]]
    if rep.endswith(".csv"):     # already exported with `ncu -i x.ncu-rep --page raw --csv` on the GPU box
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    res = {name: [] for name, _ in groups}
    for r in rows[2:]:
        for name, rx in groups:
            if re.search(rx, r[ki]):
                d = {}
                for i, h in enumerate(hdr):
                    if h in KEEP and r[i] != "":
                        d[h] = f"{r[i]} {units[i]}".strip()
                d["kernel"] = r[ki][:160]
                res[name].append(d)
                break
    json.dump(res, open(out, "w"), indent=1)
    for name, v in res.items():
        print(name, len(v), [x.get("gpu__time_duration.sum") for x in v][:6])


if __name__ == "__main__":
    main()
