"""Mirror of ``nerf_sampling.nerf_pytorch`` for the render_rays hot path (same names, same call contracts)."""
