#!/bin/bash
# ncu --set full captures of the render step (fast NeRF kernel, guard band, DepthNet, composite) and of one training step (grouped
# 3xTF32 GEMMs + chain kernels).  The reports are exported to CSV (raw page) on the box and deleted: gpurun merges at most 64 MiB back.
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"nerf_fast|composite|mlp_exact" --launch-skip 4 -c 4 -f -o gpurun_out/prof_step python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "ncu step rc=$?" | tee gpurun_out/summary_prof.txt
ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null; rm -f gpurun_out/prof_step.ncu-rep
python tools/profile_train.py > gpurun_out/plain3.log 2>&1 &&
# per eager step: chain_fwd, 2 tgemm (branch product, cat_layers.0), catchain_pack, mlp_exact<4> (cat chain), <2> <3> (d raw / d z), <5> (Jacobian chain),
# tgemm (J_0 W_cat0), 2 mlp_exact<0> (guard bands of the target), tgemm (G), chain_bwd, tgemm_reg (weight gradients), chain_du = 16 launches
ncu --set full --clock-control none -k regex:"tgemm|chain|mlp_exact" --launch-skip 16 -c 16 -f -o gpurun_out/prof_train python tools/profile_train.py > gpurun_out/ncu_train.log 2>&1
echo "ncu train rc=$?" | tee -a gpurun_out/summary_prof.txt
ncu -i gpurun_out/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null; rm -f gpurun_out/prof_train.ncu-rep
ls -la gpurun_out | head -40
