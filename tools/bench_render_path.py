"""render_path (the reference-facing multi-view API, nerf_utils.py:258-360) on N 800x800 poses: views/s and rays/s end to
end, host images included (pinned buffers, async D2H overlapped with the next view)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200.nerf_pytorch import nerf_utils  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT  # noqa: E402
from nerf_sampling_b200.trainers import DepthNetTrainer  # noqa: E402

n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
coarse, fine, dn = bench.build_models(dev, PREC_FAST)
dn.precision = PREC_SPLIT
tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=False,
                     white_bkgd=True, device=str(dev), n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                     input_dims_embed=3, distance=0.1, sampling_mode="uniform", n_depth_samples=64)
kw = dict(network_fn=coarse, network_fine=fine, depth_network=dn, network_query_fn=None, N_samples=64, N_importance=128,
          trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=False, lindisp=True, ndc=False, near=2.0, far=6.0,
          use_viewdirs=True, model_mode="test")
K = bench.intrinsics()
poses = torch.stack([bench.pose_spherical(-180.0 + 1.8 * i, -30.0, 4.031) for i in range(n_views)])
with torch.no_grad():
    warm = nerf_utils.render_path(poses, [bench.H, bench.W, float(K[0][0])], K, 32768, kw)   # also warms the pinned-memory cache
    del warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rgbs, disps, _ = nerf_utils.render_path(poses, [bench.H, bench.W, float(K[0][0])], K, 32768, kw)
    dt = time.perf_counter() - t0
print(json.dumps({"metric": "rays_per_sec", "value": n_views * bench.H * bench.W / dt, "ms_per_view": 1e3 * dt / n_views, "views": n_views,
                  "api": "nerf_utils.render_path(chunk=32768) -> numpy rgbs/disps on the host", "shape": list(rgbs.shape)}))
