"""Kernel timeline of ONE graphed training step (start, duration, stream, name), from the torch profiler: shows which launches
sit on the critical path and what overlaps what.  usage: python tools/train_timeline.py [n_rays] [--eager]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import ops, training  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT  # noqa: E402
from nerf_sampling_b200.trainers import DepthNetTrainer  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
eager = "--eager" in sys.argv
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
if eager:
    run = lambda i: tr.core_optimization_loop(opt, kw, rays, i, target)  # noqa: E731
else:
    graphed = training.GraphedTrainStep(tr, opt, kw, n_total)
    run = lambda i: graphed(rays, target)  # noqa: E731
for i in range(5):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(50):
    run(i)
e1.record()
torch.cuda.synchronize()
print("rays %d, %s: %.3f ms/step" % (n_total, "eager" if eager else "graph", e0.elapsed_time(e1) / 50))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    run(100)
    torch.cuda.synchronize()
    run(101)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time > 0]
evs.sort(key=lambda e: e.time_range.start)
# the second step: everything after the largest idle gap
starts = [e.time_range.start for e in evs]
gap_i = max(range(1, len(evs)), key=lambda i: starts[i] - evs[i - 1].time_range.end)
evs = evs[gap_i:]
t0 = evs[0].time_range.start
streams = {}
end_prev = {}
for e in evs:
    sid = getattr(e, "device_resource_id", None)
    if sid is None:
        sid = getattr(e, "device_index", 0)
    streams.setdefault(sid, len(streams))
    name = e.name.replace("b200::", "").split("(")[0][:44]
    print("%8.1f %7.1f  s%d  %s" % (e.time_range.start - t0, e.device_time, streams[sid], name))
print("step span %.1f us" % (max(e.time_range.end for e in evs) - t0))
