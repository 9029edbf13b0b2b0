#!/bin/bash
# Single-GPU pass of a round: parity tests, smoke, bench (+ reference arm, + other precision modes), ncu launch list, training A/B numbers.
# Everything lands in gpurun_out/ (which gpurun merges back only when it stays under 64 MiB: the ncu --set full captures live in
# tools/gpu_profile.sh, which exports them to CSV on the box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --prec split --no-cpu-baseline --no-extra > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err; echo "bench split rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --prec fp16 --no-cpu-baseline --no-extra > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench fp16 rc=$?" | tee -a gpurun_out/summary.txt
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" | tee -a gpurun_out/summary.txt
# training step A/B: default (eager, graph), then one switch at a time back to the previous form, then the 512-ray shard of an 8-GPU job
(timeout 300 python tools/bench_train.py 100; timeout 300 python tools/bench_train.py 200 --graph;
 B200NERF_TRAIN_CHAIN=gemm timeout 300 python tools/bench_train.py 200 --graph;
 B200NERF_TRAIN_JVP=fp32 timeout 300 python tools/bench_train.py 200 --graph;
 B200NERF_SPLIT_BACKWARD=0 timeout 300 python tools/bench_train.py 200 --graph;
 B200NERF_TGEMM=reg timeout 300 python tools/bench_train.py 200 --graph;
 B200NERF_TARGET_SMS=0 timeout 300 python tools/bench_train.py 200 --graph;
 B200NERF_TRAIN_CHAIN=gemm B200NERF_TRAIN_JVP=fp32 B200NERF_SPLIT_BACKWARD=0 B200NERF_TGEMM=reg B200NERF_TARGET_SMS=128 timeout 300 python tools/bench_train.py 200 --graph;
 timeout 300 python tools/bench_train.py 200 --graph --rays 512;
 B200NERF_TRAIN_GEMM=fp32 timeout 300 python tools/bench_train.py 30) > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python tools/train_timeline.py 4096 > gpurun_out/train_timeline_4096.txt 2>&1
timeout 300 python tools/train_timeline.py 512 > gpurun_out/train_timeline_512.txt 2>&1
timeout 300 python tools/train_breakdown.py 4096 > gpurun_out/train_breakdown.txt 2>&1; echo "train breakdown rc=$?" | tee -a gpurun_out/summary.txt
timeout 120 python tools/tgemm_timeline.py > gpurun_out/tgemm_timeline.txt 2>&1
(timeout 300 python tools/profile_composite.py 64; timeout 300 python tools/profile_composite.py 32; timeout 300 python tools/profile_composite.py 128) > gpurun_out/bench_composite.txt 2>&1; echo "composite rc=$?" | tee -a gpurun_out/summary.txt
tail -n 4 gpurun_out/t_gpu.log gpurun_out/smoke.log
cat gpurun_out/bench_train.json
