"""Turn an `ncu --set full` report into the two artefacts the repo tracks:

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01d_ncu_full_metrics.csv [--traffic]

  * a CSV with one row per profiled launch and the metrics DESIGN.md quotes (duration, DRAM bytes,
    pipe utilisation, issue-slot utilisation, stall reasons);
  * with --traffic, profiles/ncu_traffic.json: mean DRAM bytes (read + write) per launch per kernel,
    which bench.py reports as roofline.traffic.

Runs `ncu -i <rep> --page raw --csv` (the report is read, nothing is profiled), so it works in the
CPU-only build container.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
UNIT_TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    cols = ["ID", "Kernel Name", "Block Size", "Grid Size"] + [k for k in KEEP if k in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in data:
            w.writerow([r[idx[c]] for c in cols])
    print(f"{out}: {len(data)} launches, {len(cols)} columns")
    if "--traffic" in sys.argv:
        agg = {}
        for r in data:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip()
            b = sum(float(r[idx[m]].replace(",", "")) * UNIT_TO_BYTES[units[idx[m]]]
                    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            agg.setdefault(name, []).append(b)
        tj = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v),
                  "source": f"{out} (ncu --set full --clock-control none, profiles/run_kernels.py all)"}
              for k, v in agg.items()}
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
        with open(path, "w") as f:
            json.dump(tj, f, indent=1)
        print(path, "written:", ", ".join(tj))


if __name__ == "__main__":
    main()
