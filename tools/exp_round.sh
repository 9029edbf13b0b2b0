mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "training_gemm" > gpurun_out/t_gemm.log 2>&1; echo "pytest gemm rc=$?"
tail -n 15 gpurun_out/t_gemm.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "training or depthnet_literal or jvp or optimization or graphed" > gpurun_out/t_train.log 2>&1; echo "pytest train rc=$?"
tail -n 15 gpurun_out/t_train.log
timeout 300 python tools/bench_train.py 20 2>&1 | tail -n 3
B200NERF_TRAIN_GEMM=fp32 timeout 300 python tools/bench_train.py 20 2>&1 | tail -n 1
timeout 300 python tools/train_breakdown.py 4096 2>&1 | cut -c1-150 | grep -v "^---" | head -24
