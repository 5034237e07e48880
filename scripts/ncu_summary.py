"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-launch metrics cited in DESIGN.md / bench.py."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [hdr.index("Kernel Name")] + [hdr.index(w) for w in WANT if w in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[c] + (f" [{units[c]}]" if units[c] else "") for c in cols])
for r in rows[2:]:
    w.writerow([r[c][:80] for c in cols])
