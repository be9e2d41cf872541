"""Summarise ncu --set full captures into the JSON kept under profiles/ (the .ncu-rep files stay in gpurun_out/: too large).

  python tools/ncu_summary.py out.json label=gpurun_out/x.ncu-rep [label=...]

Reads each report with `ncu -i <rep> --page raw --csv` and keeps the metrics DESIGN.md and bench.py's roofline.traffic quote.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed.sum": "thread_instructions",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__pcsamp_sample_buffer_full": None,
}


def read_report(path):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        return [{"error": r.stderr[-300:]}]
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for i, name in enumerate(hdr):
            k = KEEP.get(name)
            if not k or i >= len(row):
                continue
            try:
                d[k] = float(row[i].replace(",", ""))
            except ValueError:
                continue
            d[k + "_unit"] = units[i]
        out.append(d)
    return out


def main():
    out_path, items = sys.argv[1], sys.argv[2:]
    res = {"source": "ncu --set full --clock-control none --import-source on; one launch per capture (see tools/gpu_r2_*.sh for the command lines)", "kernels": {}}
    for it in items:
        label, path = it.split("=", 1)
        res["kernels"][label] = read_report(path)
    json.dump(res, open(out_path, "w"), indent=1)
    for label, ks in res["kernels"].items():
        for k in ks:
            print(label, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in k.items() if not a.endswith("_unit")})


if __name__ == "__main__":
    main()
