"""DepthNet training step (BASELINE config #5): 4096 rays per step, data parallel over the ranks of torch.distributed.
One step = Trainer.core_optimization_loop: hierarchical target from the frozen NeRFs (64 + 192 evaluations per ray on the
tensor-core kernels), DepthNet fp32 forward/backward, colour gradient through the frozen fine NeRF, gradient all-reduce,
Adam.  Prints one JSON line (steps/s and rays/s); launched like bench.py (python or torch.distributed.run)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import ops, training  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT  # noqa: E402
from nerf_sampling_b200.trainers import DepthNetTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
use_graph = "--graph" in sys.argv
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
coarse, fine, dn = bench.build_models(dev, PREC_FAST)
dn.precision = PREC_SPLIT
tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=False,
                     white_bkgd=True, device=str(dev), n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                     input_dims_embed=3, perturb=0.0)
tr.H, tr.W, tr.K, tr.chunk = bench.H, bench.W, bench.intrinsics(), 32768
kw = dict(network_fn=coarse, network_fine=fine, depth_network=dn, network_query_fn=None, N_samples=64, N_importance=128,
          trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=0.0, lindisp=True, ndc=False, near=2.0, far=6.0,
          use_viewdirs=True, model_mode="train")
n_total = int(sys.argv[sys.argv.index("--rays") + 1]) if "--rays" in sys.argv else 4096
per = n_total // world
ro, rd, _ = ops.get_rays(bench.H, bench.W, bench.intrinsics(), bench.pose_for_step(0), dev)
sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(0))[:n_total][rank * per:(rank + 1) * per].to(dev)
rays = (ro[sel].contiguous(), rd[sel].contiguous())
target = torch.rand(n_total, 3, generator=torch.Generator().manual_seed(1))[rank * per:(rank + 1) * per].to(dev)
opt = training.Adam(list(dn.parameters()), lr=1e-4)
if use_graph:
    graphed = training.GraphedTrainStep(tr, opt, kw, per)
    run = lambda i: graphed(rays, target)  # noqa: E731
else:
    run = lambda i: tr.core_optimization_loop(opt, kw, rays, i, target)  # noqa: E731
for i in range(3):
    out = run(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    out = run(3 + i)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": "train_steps_per_sec", "value": 1e3 / float(ms), "ms_per_step": float(ms), "rays_per_step": n_total,
                      "rays_per_sec": n_total * 1e3 / float(ms), "n_gpus": world, "cuda_graph": use_graph, "loss": float(out[0]), "depth_net_loss": float(out[1]),
                      "config": "DepthNet training step, %d rays/batch, 64+128 hierarchical target, data parallel (BASELINE config #5)" % n_total,
                      "switches": {k: v for k, v in os.environ.items() if k.startswith("B200NERF_")}}))
if world > 1:
    dist.destroy_process_group()
