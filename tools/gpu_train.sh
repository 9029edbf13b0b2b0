#!/bin/bash
# training-step pass: parity tests of the training slice, then the 4096-ray step (eager / graph, A/B switches), per-kernel breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "train or fused or gradient or jvp or depthnet_literal or core_optimization or graphed or plugin or adam or gemm" > gpurun_out/t_train.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary_train.txt
(timeout 300 python tools/bench_train.py 30; timeout 300 python tools/bench_train.py 30 --graph; B200NERF_TARGET_SMS=0 timeout 300 python tools/bench_train.py 30 --graph; B200NERF_PDL=0 timeout 300 python tools/bench_train.py 30 --graph; B200NERF_TRAIN_BRANCHES=literal timeout 300 python tools/bench_train.py 30 --graph) > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?" | tee -a gpurun_out/summary_train.txt
timeout 300 python tools/train_breakdown.py 4096 > gpurun_out/train_breakdown.txt 2>&1; echo "train breakdown rc=$?" | tee -a gpurun_out/summary_train.txt
grep -n "passed\|failed\|FAILED\|Error\|worst\|config #" gpurun_out/t_train.log | cut -c1-300
cut -c1-200 gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err; grep -v "^--\|Self C\|^$" gpurun_out/train_breakdown.txt | tail -22 | cut -c1-60,150-215
