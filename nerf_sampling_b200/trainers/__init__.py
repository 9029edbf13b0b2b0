"""Mirror of ``nerf_sampling.trainers``."""
from .sampling_trainer import DepthNetTrainer  # noqa: F401
