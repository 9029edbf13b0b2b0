#!/bin/bash
# Full GPU round: parity suite, smoke, bench (both arms), ncu launch list + full capture of the two roofline kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --prec bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
if [ "$1" != "noprof" ]; then
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 3 -c 1 -o gpurun_out/prof_mlp -f python tools/profile_step.py > gpurun_out/ncu_mlp.log 2>&1
echo "ncu mlp rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite -s 1 -c 1 -o gpurun_out/prof_comp -f python tools/profile_step.py > gpurun_out/ncu_comp.log 2>&1
echo "ncu comp rc=$?" | tee -a gpurun_out/summary.txt
fi
tail -15 gpurun_out/t_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; cat gpurun_out/bench_bf16.json; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench.err
