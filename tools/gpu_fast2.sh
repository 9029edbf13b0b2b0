#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fast_check.py 4096 64 > gpurun_out/fast_small.log 2>&1; echo "small rc=$?"
SKIP_SPLIT_TIMING=1 timeout 300 python tools/fast_check.py 640000 64 > gpurun_out/fast_big.log 2>&1; echo "big rc=$?"
tail -n 9 gpurun_out/fast_small.log gpurun_out/fast_big.log
python tools/profile_fast.py 131072 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nerf_fast -s 2 -c 1 -f -o gpurun_out/prof_fast python tools/profile_fast.py 131072 > gpurun_out/ncu_fast.log 2>&1
echo "ncu rc=$?"
