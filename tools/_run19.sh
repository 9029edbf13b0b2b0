cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/t19.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|E  )|passed|failed|DepthNet training form" gpurun_out/t19.log | cut -c1-250 | head -20
