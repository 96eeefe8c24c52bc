"""Condenses an .ncu-rep (raw page) into the per-kernel metric table kept under profiles/."""
import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__warps_eligible.avg.per_cycle_active"]
ki = hdr.index("Kernel Name")
with open(out, "w") as f:
    f.write(f"# {title}\n")
    for r in rows[2:]:
        f.write(f"## {r[ki]}\n")
        for h, u, v in zip(hdr, units, r):
            tensor = h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed") or \
                h == "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
            if h in keys or tensor or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                try:
                    if h.startswith("smsp__average_warps") and float(v.replace(",", "")) < 0.02:
                        continue
                except ValueError:
                    pass
                f.write(f"{h}\t{u}\t{v}\n")
