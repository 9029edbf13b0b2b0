cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t16.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary16.txt
grep -E "^(FAILED|E  )|passed|failed" gpurun_out/t16.log | cut -c1-250 | head -30 >> gpurun_out/summary16.txt
for n in 4096 512; do
  timeout 300 python tools/train_timeline.py $n > gpurun_out/tl11_$n.txt 2>&1
  echo "$n: $(head -1 gpurun_out/tl11_$n.txt)" >> gpurun_out/summary16.txt
done
for sm in 120 136 144; do
  B200NERF_TARGET_SMS=$sm timeout 300 python tools/bench_train.py 300 --graph 2>/dev/null | tail -1 | cut -c1-110 | sed "s/^/sms=$sm /" >> gpurun_out/summary16.txt
done
cat gpurun_out/summary16.txt
