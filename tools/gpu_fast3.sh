#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fast_check.py 4096 64 > gpurun_out/fast_small.log 2>&1; echo "small rc=$?"
SKIP_SPLIT_TIMING=1 timeout 300 python tools/fast_check.py 640000 64 > gpurun_out/fast_big.log 2>&1; echo "big rc=$?"
grep -E "fp16 vs|guarded|identical|fp16:|fast:" gpurun_out/fast_small.log gpurun_out/fast_big.log
B200NERF_LIB=nerf_sampling_b200/libb200nerf_tl.so timeout 200 python tools/timeline.py > gpurun_out/timeline.log 2>&1; echo "timeline rc=$?"
