"""Mirror of ``nerf_sampling.depth_nets``."""
from .depth_net import DepthNet  # noqa: F401
