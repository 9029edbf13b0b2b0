"""One launch of the throughput NeRF kernel on a reduced view, for ncu (tools/gpu_prof_fast.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import ops  # noqa: E402
from nerf_sampling_b200.packing import PREC_FP16, PREC_SPLIT, PackedNeRF  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
S = 64
dev = torch.device("cuda", 0)
coarse, fine, dn = bench.build_models(dev, PREC_SPLIT)
pk = PackedNeRF(fine.state_dict(), dev, PREC_FP16)
ro, rd, vd = ops.get_rays(bench.H, bench.W, bench.intrinsics(), bench.pose_for_step(0), dev)
ro, rd, vd = ro[:n_rays].contiguous(), rd[:n_rays].contiguous(), vd[:n_rays].contiguous()
z = ops.place_samples(ops.depthnet_forward(dn.packed(), ro, rd), S, "uniform", 0.1)
for _ in range(3):
    raw = ops.nerf_mlp(pk, vd, rays_o=ro, rays_d=rd, z=z)
torch.cuda.synchronize()
print("ok", float(raw.abs().mean()))
