#!/bin/bash
# First contact with the GPU: layout self-test first, then the parity suite, each under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "umma" > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not umma and not full_size" > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "full_size" > gpurun_out/t_full.log 2>&1; echo "full rc=$?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/t_umma.log; tail -60 gpurun_out/t_parity.log; tail -30 gpurun_out/t_full.log
