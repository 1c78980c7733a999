#!/usr/bin/env python3
"""Key numbers of one `ncu --set full` capture as JSON (reads `ncu -i rep --page raw --csv`).
usage: ncu_summary.py <rep.ncu-rep> <kernel label> <input_bytes> <output_bytes> "<command the capture came from>" """
import csv, io, json, subprocess, sys
rep, label, inb, outb, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
def get(name, scale=None):
    if name not in h: return None
    i = h.index(name)
    try: x = float(v[i].replace(",", ""))
    except ValueError: return v[i]
    unit = u[i]
    if scale == "bytes":
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    if scale == "smem_kb":
        x *= {"byte/block": 1 / 1024, "Kbyte/block": 1000 / 1024, "Mbyte/block": 1e6 / 1024}.get(unit, 1)
    if scale == "ms":
        x *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)
    return x
shared = get("smsp__inst_executed_op_shared.sum") or 0
out = {
    "source": cmd, "kernel": label, "input_bytes": inb, "output_bytes": outb,
    "duration_ms": get("gpu__time_duration.sum", "ms"),
    "dram_bytes_read": get("dram__bytes_read.sum", "bytes"), "dram_bytes_write": get("dram__bytes_write.sum", "bytes"),
    "dram_throughput_pct_of_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "warp_instructions": get("smsp__inst_executed.sum"),
    "inst_per_cycle_per_sm": get("sm__inst_executed.avg.per_cycle_active"),
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "threads_per_warp_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "lsu_wavefronts_pct_of_peak": get("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active") or get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_bank_conflict_wavefronts": get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
    "registers_per_thread": get("launch__registers_per_thread"), "grid": get("launch__grid_size"), "block": get("launch__block_size"),
    "dynamic_smem_kb": get("launch__shared_mem_per_block_dynamic", "smem_kb"),
    "stalls_per_issue": {n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(float(v[i]), 3)
                          for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and float(v[i]) >= 0.05},
}
if out["warp_instructions"] and inb:
    out["warp_instructions_per_input_byte"] = out["warp_instructions"] / inb
if out["duration_ms"]:
    out["algorithmic_GBps"] = (inb + outb) / out["duration_ms"] / 1e6
print(json.dumps(out, indent=1))
