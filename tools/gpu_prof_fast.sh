#!/bin/bash
mkdir -p gpurun_out
python tools/profile_fast.py 131072 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nerf_fast -s 2 -c 1 -f -o gpurun_out/prof_fast python tools/profile_fast.py 131072 > gpurun_out/ncu_fast.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/prof_plain.log gpurun_out/ncu_fast.log
