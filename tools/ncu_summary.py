"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the markdown tables kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_step.ncu-rep "<title line>" [--traffic PREC profiles/rNN_ncu_traffic.json]
           > profiles/rNN_ncu_summary.md

--traffic also writes the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE NeRF-MLP query of precision mode PREC
(nerf_fast_kernel + the guard-band mlp_exact_kernel<0> launch that follows it) as JSON: bench.py reports that file's value as
`roofline.traffic`, so the number on the bench line is always the newest committed capture, never a constant."""
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
traffic = sys.argv[sys.argv.index("--traffic") + 1: sys.argv.index("--traffic") + 3] if "--traffic" in sys.argv else None
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout   # a raw-page CSV exported on the GPU box, or the report itself
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

if traffic:
    import json
    import os

    def mb(r, name):
        v, unit = float(r[col[name]]), units[col[name]].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]

    per = {}
    for r in data:
        k = r[col["Kernel Name"]]
        key = "fast" if "nerf_fast_kernel" in k else ("guard" if "mlp_exact_kernel<0>" in k or "mlp_exact_kernel<(int)0>" in k else None)
        if key and key not in per:   # first launch of each
            per[key] = {"read": mb(r, "dram__bytes_read.sum"), "write": mb(r, "dram__bytes_write.sum"),
                        "ms": float(r[col["gpu__time_duration.sum"]]) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(units[col["gpu__time_duration.sum"]].lower().replace("second", "s").replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1.0)}
    total = sum(v["read"] + v["write"] for v in per.values())
    out = {traffic[0]: {"bytes": total, "kernels": per, "source": os.path.basename(rep),
                        "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, one 800x800x64 NeRF-MLP query"}}
    prev = {}
    if os.path.exists(traffic[1]):
        prev = json.load(open(traffic[1]))
    prev.update(out)
    json.dump(prev, open(traffic[1], "w"), indent=1)
