mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "training or depthnet_literal or jvp or optimization or graphed" > gpurun_out/t_train.log 2>&1; echo "pytest train rc=$?"
tail -n 3 gpurun_out/t_train.log
timeout 300 python tools/tgemm_timeline.py 2>&1 | tail -n 4
timeout 300 python tools/bench_train.py 20 2>&1 | tail -n 1
timeout 300 python tools/train_breakdown.py 4096 2>&1 | tail -n 3 | cut -c1-1500
