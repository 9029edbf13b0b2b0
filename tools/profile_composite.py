"""raw2outputs kernel alone on 800x800xS synthetic inputs (timing with CUDA events; wrap in ncu for counters).

usage: python tools/profile_composite.py [S] [heat]   -- heat=1 runs a 30 ms tensor-core load before every timed launch so the
composite is measured at the power-capped clocks it sees inside a render step.  B200NERF_COMPOSITE_LDG=1 selects the
register-staged kernel for S = 32 / 64 / 128 (default there: the TMA-staged kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sampling_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
n, S = 640000, int(sys.argv[1]) if len(sys.argv) > 1 else 64
heat = len(sys.argv) > 2 and sys.argv[2] == "1"
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn(n, S, 4, device=dev, generator=g)
z = torch.sort(2 + 4 * torch.rand(n, S, device=dev, generator=g), -1).values
rd = torch.randn(n, 3, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rgb = torch.empty(n, 3, device=dev)
disp, acc, depth = (torch.empty(n, device=dev) for _ in range(3))
w = torch.empty(n, S, device=dev)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
L = _lib.lib()


def run():
    _lib.check(L.b200nerf_composite_fwd(raw.data_ptr(), z.data_ptr(), rd.data_ptr(), None, n, S, 1, rgb.data_ptr(), disp.data_ptr(),
                                        acc.data_ptr(), depth.data_ptr(), w.data_ptr(), None, st))


for _ in range(3):
    run()
ts = []
for _ in range(10):
    if heat:
        for _ in range(30):
            a @ a
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
kind = "ldg" if os.environ.get("B200NERF_COMPOSITE_LDG") == "1" else "tma"
print("composite[%s%s] S=%d: %.4f ms median, min %.4f, %.1f GB/s algorithmic (%.3f of 6464.9)" % (
    kind, ",heated" if heat else "", S, ms, min(ts), (24 * S + 36) * n / ms / 1e6, (24 * S + 36) * n / ms / 1e6 / 6464.9))
# cross-check against the torch formulation on a slice
ref = ops.composite(raw[:4096], z[:4096], rd[:4096], True, want_alphas=False)
print("max |rgb - first-4096 recompute| %.2e" % float((ref[0] - rgb[:4096]).abs().max()))
