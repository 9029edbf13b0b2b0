"""Vanilla hierarchical baseline (BASELINE config #4): 800x800 view, 64 coarse + 128 fine samples with sample_pdf, all 192
re-evaluated by the fine NeRF, through the reference-facing API (render_test with trainer.use_full_nerf).  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200.nerf_pytorch import nerf_utils  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT  # noqa: E402
from nerf_sampling_b200.trainers import DepthNetTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
coarse, fine, dn = bench.build_models(dev, PREC_FAST)
dn.precision = PREC_SPLIT
tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=False,
                     white_bkgd=True, device=str(dev), n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                     input_dims_embed=3, distance=0.1, sampling_mode="uniform", n_depth_samples=64, use_full_nerf=True)
kw = dict(network_fn=coarse, network_fine=fine, depth_network=dn, network_query_fn=None, N_samples=64, N_importance=128,
          trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=False, lindisp=True, ndc=False, near=2.0, far=6.0,
          use_viewdirs=True, model_mode="test")
H = W = bench.H
K = bench.intrinsics()
with torch.no_grad():
    for i in range(2):
        nerf_utils.render_test(H, W, K, chunk=H * W // 4, c2w=bench.pose_for_step(i), **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rgb, disp, ex = nerf_utils.render_test(H, W, K, chunk=H * W // 4, c2w=bench.pose_for_step(2 + i), **kw)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flop = 256 * 1_186_816 * H * W
print(json.dumps({"metric": "rays_per_sec", "value": H * W / ms * 1e3, "ms_per_view": ms, "algorithmic_tflops": flop / ms / 1e9,
                  "frac_of_measured_peak": flop / ms / 1e9 / bench.peaks()["tf_sustained"],
                  "config": "vanilla hierarchical: 64 coarse + 128 fine (192 re-evaluated), sample_pdf, 800x800, 1 B200, chunk = 160,000 rays"}))
