#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one_step.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench
from nerf_sampling_b200 import ops, training
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT
dev = torch.device("cuda", 0)
models = bench.build_models(dev, PREC_FAST); models[2].precision = PREC_SPLIT
tr, kw = bench.make_trainer(models, dev); kw["model_mode"] = "train"
ro, rd, _ = ops.get_rays(bench.H, bench.W, bench.intrinsics(), bench.pose_for_step(0), dev)
sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(0))[:4096].to(dev)
rays = (ro[sel].contiguous(), rd[sel].contiguous()); target = torch.rand(4096, 3, generator=torch.Generator().manual_seed(1)).to(dev)
opt = training.Adam(list(models[2].parameters()), lr=1e-4)
for i in range(3): tr.core_optimization_loop(opt, kw, rays, i, target)
torch.cuda.synchronize(); print("ok")
PY
python /tmp/one_step.py > gpurun_out/one_step.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"chain_fwd|chain_bwd" -c 2 -f -o gpurun_out/prof_chain python /tmp/one_step.py > gpurun_out/ncu_chain.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_chain.log
