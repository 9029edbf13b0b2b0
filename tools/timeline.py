"""Clock-stamp timeline of cluster 0's first units of the throughput kernel (needs a -DB200NERF_TIMELINE build):
    B200NERF_LIB=nerf_sampling_b200/libb200nerf_tl.so python tools/timeline.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import _lib, ops  # noqa: E402
from nerf_sampling_b200.packing import PREC_FP16, PREC_SPLIT, PackedNeRF  # noqa: E402

n_rays = 131072
S = 64
dev = torch.device("cuda", 0)
coarse, fine, dn = bench.build_models(dev, PREC_SPLIT)
pk = PackedNeRF(fine.state_dict(), dev, PREC_FP16)
ro, rd, vd = ops.get_rays(bench.H, bench.W, bench.intrinsics(), bench.pose_for_step(0), dev)
ro, rd, vd = ro[:n_rays].contiguous(), rd[:n_rays].contiguous(), vd[:n_rays].contiguous()
z = ops.place_samples(ops.depthnet_forward(dn.packed(), ro, rd), S, "uniform", 0.1)
L = _lib.lib()
L.b200nerf_debug_set_timeline.argtypes = [C.c_void_p]
L.b200nerf_debug_set_timeline.restype = None
for _ in range(2):
    ops.nerf_mlp(pk, vd, rays_o=ro, rays_d=rd, z=z)
buf = torch.zeros(3 * 80 * 4, dtype=torch.int64, device=dev)
L.b200nerf_debug_set_timeline(buf.data_ptr())
ops.nerf_mlp(pk, vd, rays_o=ro, rays_d=rd, z=z)
torch.cuda.synchronize()
L.b200nerf_debug_set_timeline(None)
t = buf.cpu().reshape(3, 80, 4)
t0 = int(t[0, 0, 0])
names = ["mma", "epi_h0", "epi_h1"]
print("idx unit step slot | mma: wait_begin wait_end issue_end (wait, issue) | epi_h0: poll acc_full done (wait, work) | epi_h1 ...")
for i in range(20, 80):
    u, r = divmod(i, 20)
    s, slot = divmod(r, 2)
    m = [int(x) - t0 for x in t[0, i, :3]]
    e0 = [int(x) - t0 for x in t[1, i, :3]]
    e1 = [int(x) - t0 for x in t[2, i, :3]]
    print(f"{i:3d} u{u} s{s} x{slot} | mma {m[0]:8d} {m[1]:8d} {m[2]:8d} (wait {m[1]-m[0]:5d} issue {m[2]-m[1]:5d}) | "
          f"e0 {e0[1]:8d} {e0[2]:8d} (wait {e0[1]-e0[0]:5d} work {e0[2]-e0[1]:5d}) | e1 (wait {e1[1]-e1[0]:5d} work {e1[2]-e1[1]:5d}) "
          f"| acc_full->seen {e0[1]-m[2]:5d}  arrive->mma {m[1]-max(e0[2],e1[2]) if i+2<80 else 0}")
