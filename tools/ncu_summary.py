"""Summarise an `ncu --set full` report into the JSON committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r01d_prof.ncu-rep profiles/r01d_ncu_full_summary.json

Reads the report with `ncu -i <rep> --page raw --csv` (one row per profiled launch) and keeps the
metrics DESIGN.md / bench.py quote: duration, DRAM bytes (-> roofline.traffic), pipe utilisation,
issue-slot utilisation, occupancy limits, shared-memory bank conflicts, top stall reasons.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
        if name in res:
            continue
        d = {}
        for k in KEEP:
            if k in col:
                d[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        stalls = []
        for h, i in col.items():
            if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        d["top_stalls_warps_per_issue"] = {n: round(v, 3) for v, n in stalls[:6]}
        res[name] = d
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
