"""Diagnostic: activations saved by the DepthNet training forward, fused split-precision chain vs per-layer 3xTF32 products."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_sampling_b200 import _lib, ops, training  # noqa: E402
from nerf_sampling_b200.packing import PREC_FAST  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 117
dev = torch.device("cuda", 0)
_, _, dn = bench.build_models(dev, PREC_FAST)
params = training.depthnet_params(dn)
hidden, cat = training.depthnet_arch(dn)
L = _lib.lib()
ro, rd, _ = ops.get_rays(bench.H, bench.W, bench.intrinsics(), bench.pose_for_step(0), dev)
sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(0))[:n].to(dev)
ro, rd = ro[sel].contiguous(), rd[sel].contiguous()
ints = lambda v: (C.c_int * len(v))(*v)  # noqa: E731
ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
nf = L.b200nerf_depthnet_train_ws_floats(n, len(hidden), ints(hidden), len(cat), ints(cat))
take = lambda f: (f + 63) // 64 * 64  # noqa: E731
off = take(n * 252) + 3 * len(hidden) * take(n * 256)
out = {}
for mode in ("gemm", "fused"):
    os.environ["B200NERF_TRAIN_CHAIN"] = mode
    ws = torch.full((nf,), float('nan'), device=dev) if len(sys.argv) > 2 else torch.zeros(nf, device=dev)
    z = torch.empty(n, 1, device=dev)
    _lib.check(L.b200nerf_depthnet_train_fwd(ptrs(params), len(hidden), ints(hidden), len(cat), ints(cat), ro.data_ptr(), rd.data_ptr(), n,
                                             float(dn.sphere_radius), float(dn.near), float(dn.far), ws.data_ptr(), z.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    dz = (torch.linspace(0.5, 1.5, n, device=dev)).contiguous()
    grads = [torch.zeros_like(p) for p in params]
    _lib.check(L.b200nerf_depthnet_train_bwd(ptrs(params), len(hidden), ints(hidden), len(cat), ints(cat), n, float(dn.near), float(dn.far),
                                             ws.data_ptr(), dz.data_ptr(), ptrs(grads), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out[mode + "_g"] = grads
    a = [ws[off + j * take(n * 256): off + j * take(n * 256) + n * 256].reshape(n, 256).clone() for j in range(len(cat))]
    o2 = off + len(cat) * take(n * 256)
    out[mode] = (a, ws[o2 + take(n): o2 + take(n) + n].clone(), z.clone())
for j in range(len(cat)):
    g, f = out["gemm"][0][j], out["fused"][0][j]
    d = (g - f).abs()
    print("a_%d: max|a| %.3e  max abs diff %.3e  rms diff %.3e  sign mismatches %d  worst row %d" % (
        j, float(g.abs().max()), float(d.max()), float(d.pow(2).mean().sqrt()), int(((g > 0) != (f > 0)).sum()), int(d.max(1).values.argmax())))
print("s: max diff %.3e; z: max diff %.3e" % (float((out["gemm"][1] - out["fused"][1]).abs().max()), float((out["gemm"][2] - out["fused"][2]).abs().max())))

names = [k for k, _ in dn.named_parameters()]
worst = []
for k, g, f in zip(names, out["gemm_g"], out["fused_g"]):
    worst.append((float((g - f).abs().max()) / (float(g.abs().max()) + 1e-20), float((g - f).norm()) / (float(g.norm()) + 1e-20), k))
worst.sort(reverse=True)
print("NaN gradients: gemm %d tensors, fused %d tensors" % (sum(bool(torch.isnan(g).any()) for g in out["gemm_g"]), sum(bool(torch.isnan(g).any()) for g in out["fused_g"])))
print("one-pass backward after the two forwards, worst tensors (max-abs rel, L2 rel):")
for w in worst[:8]:
    print("  %.2e %.2e %s" % w)
