cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_mid.json
