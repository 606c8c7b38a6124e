"""Reduce an .ncu-rep (ncu --set full) to the handful of metrics DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json "description" """
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg"]


def main():
    rep, out, desc = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        m = {}
        for name, unit, v in zip(hdr, units, vals):
            short = name.split("TriageCompute.")[-1]
            if short in KEYS or "pcsamp_warps_issue_stalled" in name:
                if v not in ("", "0", "n/a"):
                    m[short] = [v, unit]
        res.append({"kernel": rec.get("Kernel Name", ""), "grid": rec.get("Grid Size", ""), "block": rec.get("Block Size", ""), "metrics": m})
    with open(out, "w") as f:
        json.dump({"source": rep, "description": desc, "launches": res}, f, indent=1)
    print("wrote", out, len(res), "launch(es)")


if __name__ == "__main__":
    main()
