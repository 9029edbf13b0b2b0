cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "depthnet_literal" -s > gpurun_out/t17.log 2>&1
grep -E "DepthNet training form|passed|failed|AssertionError" gpurun_out/t17.log | cut -c1-400
