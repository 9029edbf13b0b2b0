"""Two eager training steps (config #5 shapes) for ncu: `ncu --set full -k regex:"tgemm|chain|mlp_exact" --launch-skip 16 -c 16 python
tools/profile_train.py` captures the 16 matching launches of the second step (see tools/gpu_profile.sh for the order)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import ops, training  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT  # noqa: E402
from nerf_sampling_b200.trainers import DepthNetTrainer  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
coarse, fine, dn = bench.build_models(dev, PREC_FAST)
dn.precision = PREC_SPLIT
tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=False,
                     white_bkgd=True, device=str(dev), n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                     input_dims_embed=3, perturb=0.0)
tr.H, tr.W, tr.K, tr.chunk = bench.H, bench.W, bench.intrinsics(), 32768
kw = dict(network_fn=coarse, network_fine=fine, depth_network=dn, network_query_fn=None, N_samples=64, N_importance=128,
          trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=0.0, lindisp=True, ndc=False, near=2.0, far=6.0,
          use_viewdirs=True, model_mode="train")
ro, rd, _ = ops.get_rays(bench.H, bench.W, bench.intrinsics(), bench.pose_for_step(0), dev)
sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(0))[:n_total].to(dev)
rays = (ro[sel].contiguous(), rd[sel].contiguous())
target = torch.rand(n_total, 3, generator=torch.Generator().manual_seed(1)).to(dev)
opt = training.Adam(list(dn.parameters()), lr=1e-4)
for i in range(2):
    out = tr.core_optimization_loop(opt, kw, rays, i, target)
torch.cuda.synchronize()
print("ok", float(out[0]))
