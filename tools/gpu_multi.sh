#!/bin/bash
# multi-GPU pass: bench at N = $1 ranks (weak scaling value, e2e with the gather, e2e_api = render_path(shard="views"), shard_parity,
# strong scaling, config4 / config5 at N ranks)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?" | tee gpurun_out/summary_n$N.txt
tail -n 8 gpurun_out/bench_n$N.err | cut -c1-300
grep "^{" gpurun_out/bench_n$N.json | cut -c1-400
