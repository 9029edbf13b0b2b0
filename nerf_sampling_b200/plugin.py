"""The drop-in: the reference's own ``DepthNetTrainer`` driving the B200 render_rays path.

``experiments/run.py:150-151`` and ``experiments/render.py:257-264`` build the trainer through the YAML plugin hook
``load_obj_from_config({"module": ..., "kwargs": ...})`` (``nerf_pytorch/utils.py:12-21``) and call ``.train(N_iters)``.
Pointing ``module`` at ``nerf_sampling_b200.plugin.B200DepthNetTrainer`` keeps everything of the host application that is
not the hot path -- ``load_data`` (Blender loader), ``train`` (``Trainer.py:712-787``), ``log`` (``:263-398``),
``update_learning_rate``, checkpoint naming, wandb / optuna reporting -- and replaces what §8 of SURVEY.md puts on the path:

================================  =====================================================================================
``create_nerf_model``             state_dict-compatible ``NeRF`` / ``DepthNet`` shells over the packed tensor-core images,
                                  the fused Adam; ``200000.tar`` checkpoints load verbatim (sampling_trainer.py:54-122)
``render``                        this repo's ``render_path`` (overlapped D2H, threaded PNG, optional rank sharding)
``core_optimization_loop``        fused training render + backward + (all-reduce) + Adam (Trainer.py:506-544)
``sample_random_ray_batch``       device-side ray / pixel selection (Trainer.py:400-475)
``run_network`` / ``raw2outputs`` /
``sample_coarse_points`` /
``sample_fine_points`` /
``_sample_points``                one C-ABI kernel each (Trainer.py:553-710, 789-806; sampling_trainer.py:153-230)
================================  =====================================================================================

While ``train()`` runs, the render entry points of the reference's ``nerf_utils`` module (``render``, ``render_test``,
``render_path``, ``render_rays``, ``render_rays_test``) are swapped for this repo's, so the test-set / video renders that the
untouched ``Trainer.log`` issues land on the fused path as well; they are restored on exit.

When ``nerf_sampling`` is not importable the class derives from this repo's mirror trainer instead (same operators, its own
small ``train`` loop; ``load_data`` must then be supplied by a subclass).
"""

from __future__ import annotations

import contextlib

from .nerf_pytorch import nerf_utils as _nu
from .nerf_pytorch.trainers.Trainer import Trainer as _MirrorTrainer
from .trainers.sampling_trainer import DepthNetTrainer as _MirrorDepthNetTrainer

try:  # the host application (the reference package), when it is installed next to us
    from nerf_sampling.nerf_pytorch import nerf_utils as _ref_nerf_utils
    from nerf_sampling.trainers import DepthNetTrainer as _Base

    HAVE_REFERENCE = True
except ImportError:
    _ref_nerf_utils = None
    _Base = _MirrorDepthNetTrainer
    HAVE_REFERENCE = False

_SWAPPED = ("render", "render_test", "render_path", "render_rays", "render_rays_test", "batchify_rays", "batchify_rays_test")


@contextlib.contextmanager
def reference_entry_points_on_b200():
    """Swap the reference's ``nerf_utils`` render entry points for this repo's for the duration of the block."""
    if _ref_nerf_utils is None:
        yield
        return
    saved = {k: getattr(_ref_nerf_utils, k) for k in _SWAPPED}
    try:
        for k in _SWAPPED:
            setattr(_ref_nerf_utils, k, getattr(_nu, k))
        yield
    finally:
        for k, v in saved.items():
            setattr(_ref_nerf_utils, k, v)


class B200DepthNetTrainer(_Base):
    """``module: nerf_sampling_b200.plugin.B200DepthNetTrainer`` in ``experiments/configs/lego.yaml``."""

    # ---- operators on the hot path: the mirror's bodies, bound to the host application's trainer type
    create_nerf_model = _MirrorDepthNetTrainer.create_nerf_model
    raw2outputs = _MirrorDepthNetTrainer.raw2outputs
    run_network = _MirrorTrainer.run_network
    _sample_points = _MirrorTrainer._sample_points
    sample_coarse_points = _MirrorTrainer.sample_coarse_points
    sample_fine_points = _MirrorTrainer.sample_fine_points
    sample_random_ray_batch = _MirrorTrainer.sample_random_ray_batch
    core_optimization_loop = _MirrorTrainer.core_optimization_loop
    render_and_backward = _MirrorTrainer.render_and_backward
    reduce_and_step = _MirrorTrainer.reduce_and_step
    _device_image = _MirrorTrainer._device_image
    _host_poses = _MirrorTrainer._host_poses

    render = _MirrorTrainer.render

    def train(self, N_iters=200000 + 1):
        """The host application's own loop (Trainer.py:712-787), its ``nerf_utils`` render calls routed to this repo."""
        with reference_entry_points_on_b200():
            return super().train(N_iters)
