"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the markdown tables kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_step.ncu-rep "<title line>" > profiles/rNN_ncu_summary.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__grid_size", "launch__cluster_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]

rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
head, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(head)}
print("# " + title + "\n")
for r in data:
    print("## " + r[col["Kernel Name"]][:100] + "\n")
    print("| metric | unit | value |\n|---|---|---|")
    for m in METRICS:
        if m in col:
            print("| %s | %s | %s |" % (m, units[col[m]], r[col[m]]))
    print()
