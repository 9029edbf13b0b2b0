cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "depthnet or train or split_backward or core_optimization or plugin_training or graphed or jvp" -s > gpurun_out/t15.log 2>&1; echo "pytest rc=$?" > gpurun_out/summary15.txt
grep -E "^(FAILED|E  )|passed|failed|worst|config #5|split backward" gpurun_out/t15.log | cut -c1-250 | head -30 >> gpurun_out/summary15.txt
for n in 4096 512; do
  timeout 300 python tools/train_timeline.py $n > gpurun_out/tl10_$n.txt 2>&1
  echo "$n: $(head -1 gpurun_out/tl10_$n.txt)" >> gpurun_out/summary15.txt
  B200NERF_TRAIN_CHAIN=gemm timeout 300 python tools/train_timeline.py $n 2>/dev/null | head -1 | sed "s/^/$n gemm: /" >> gpurun_out/summary15.txt
done
cat gpurun_out/summary15.txt
