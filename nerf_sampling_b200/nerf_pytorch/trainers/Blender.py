"""Mirror of ``nerf_pytorch/trainers/Blender.py``: near=2, far=6, white background flag."""

from .Trainer import Trainer


class BlenderTrainer(Trainer):
    def __init__(self, half_res, white_bkgd, testskip=8, near=2.0, far=6.0, **kwargs):
        self.half_res = half_res
        self.testskip = testskip
        self.white_bkgd = white_bkgd
        self.near = near
        self.far = far
        super().__init__(**kwargs)
