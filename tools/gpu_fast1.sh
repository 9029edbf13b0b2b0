#!/bin/bash
# first GPU contact of the throughput kernel: single-CTA variant, CTA-pair variant, then the regression suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
B200NERF_FAST_NCTA=1 timeout 300 python tools/fast_check.py 4096 64 > gpurun_out/fast_n1_small.log 2>&1; echo "n1 small rc=$?"
B200NERF_FAST_NCTA=2 timeout 300 python tools/fast_check.py 4096 64 > gpurun_out/fast_n2_small.log 2>&1; echo "n2 small rc=$?"
B200NERF_FAST_NCTA=1 SKIP_SPLIT_TIMING=1 timeout 300 python tools/fast_check.py 640000 64 > gpurun_out/fast_n1_big.log 2>&1; echo "n1 big rc=$?"
B200NERF_FAST_NCTA=2 SKIP_SPLIT_TIMING=1 timeout 300 python tools/fast_check.py 640000 64 > gpurun_out/fast_n2_big.log 2>&1; echo "n2 big rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/fast_n1_small.log gpurun_out/fast_n2_small.log gpurun_out/fast_n1_big.log gpurun_out/fast_n2_big.log
tail -n 5 gpurun_out/t_gpu.log
