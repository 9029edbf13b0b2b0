"""Two 800x800 DepthNet render steps (1 warm-up + 1) -- the short command ncu wraps (see tools/gpu_profile.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import ops  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT  # noqa: E402

prec = PREC_SPLIT if "--split" in sys.argv else PREC_FAST
dev = torch.device("cuda", 0)
coarse, fine, dn = bench.build_models(dev, prec)
dn.precision = PREC_SPLIT
K = bench.intrinsics()
for i in range(2):
    ro, rd, vd = ops.get_rays(bench.H, bench.W, K, bench.pose_for_step(i), dev)
    out = ops.render_depthnet(dn.packed(), fine.packed(), ro, rd, vd, bench.S, "uniform", bench.DISTANCE)
torch.cuda.synchronize()
print("ok", float(out["rgb"].mean()))
