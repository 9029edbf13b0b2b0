"""nerf_sampling_b200 -- the render_rays hot path of nerf-sampling on B200 (sm_100a).

PyTorch owns device memory, streams and ``torch.distributed``; everything on the data path is a hand-written CUDA
kernel in ``libb200nerf.so`` reached through the C ABI in ``include/b200nerf.h``.  The sub-packages mirror the
reference's module layout (``nerf_pytorch``, ``depth_nets``, ``trainers``).
"""

from . import _build, _lib, ops, packing  # noqa: F401
from ._build import build  # noqa: F401

__all__ = ["build", "ops", "packing"]
