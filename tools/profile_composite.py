"""raw2outputs kernel alone on 800x800x64 synthetic inputs (timing with CUDA events; wrap in ncu for counters)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sampling_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n, S = 640000, int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn(n, S, 4, device=dev, generator=g)
z = torch.sort(2 + 4 * torch.rand(n, S, device=dev, generator=g), -1).values
rd = torch.randn(n, 3, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    out = ops.composite(raw, z, rd, True, want_alphas=False)
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.composite(raw, z, rd, True, want_alphas=False)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print("composite S=%d: %.4f ms median (includes the output allocations of ops.composite), %.1f GB/s algorithmic" % (S, ms, (24 * S + 36) * n / ms / 1e6))
