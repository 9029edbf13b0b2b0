"""Phase time stamps of one CTA of the grouped 3xTF32 GEMM (csrc/tgemm.cuh) on the training shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sampling_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
L = _lib.lib()
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
L.b200nerf_debug_set_tgemm_timeline(dbg.data_ptr())
st = torch.cuda.current_stream().cuda_stream
n = 4096
for name, (M, N, K, a, sam, sak, b, sbk, sbn, ldc) in {
    "fwd 4096x256x256": (n, 256, 256, torch.randn(n, 256, device=dev), 256, 1, torch.randn(256, 256, device=dev), 1, 256, 256),
    "fwd 4096x256x1024": (n, 256, 1024, torch.randn(n, 1024, device=dev), 1024, 1, torch.randn(256, 1024, device=dev), 1, 1024, 256),
    "dgrad 4096x256x256": (n, 256, 256, torch.randn(n, 256, device=dev), 256, 1, torch.randn(256, 256, device=dev), 256, 1, 256),
    "wgrad 256x256x4096": (256, 256, n, torch.randn(n, 256, device=dev), 1, 256, torch.randn(n, 256, device=dev), 256, 1, 256),
}.items():
    c = torch.zeros(M, N, device=dev)
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.b200nerf_debug_sgemm(M, N, K, a.data_ptr(), sam, sak, b.data_ptr(), sbk, sbn, c.data_ptr(), ldc, 0, None, 0, 0.0, 0, st))
        e1.record()
        torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    print("%s: event %.1f us | ready +%.2f us, chunks at %s, last MMA retired +%.2f, stored +%.2f" % (
        name, 1e3 * e0.elapsed_time(e1), (t[1] - t[0]) / 1e3, " ".join("%.2f" % ((x - t[0]) / 1e3) for x in t[4:12] if x > t[0]),
        (t[2] - t[0]) / 1e3, (t[3] - t[0]) / 1e3))
