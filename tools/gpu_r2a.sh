#!/bin/bash
# Round-2 first GPU pass: parity tests (old + new), smoke, bench N=1 (+ reference arm)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
tail -n 30 gpurun_out/t_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 5 gpurun_out/bench.err
cat gpurun_out/bench.json gpurun_out/bench_ref.json
