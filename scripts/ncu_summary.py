"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.json + .md lines."""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_elapsed",
    "smsp__inst_executed.sum": "warp_inst",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}


def main(rep, out_json):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT.get(units[i], 1) if KEYS[h] in ("duration", "dram_read", "dram_write") else 1
                d[KEYS[h]] = v
        if "dram_read" in d:
            d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0)
        kernels.append(d)
    json.dump({"report": rep, "kernels": kernels}, open(out_json, "w"), indent=1)
    for k in kernels:
        print(json.dumps(k))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
