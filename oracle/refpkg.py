"""The UNMODIFIED reference as an importable package (TEST / BENCH INFRASTRUCTURE, never used by product code).

``install()`` is the recipe: ``pip install --no-index --no-deps --target baseline/_ref`` of a scratch copy of
``/root/reference`` (the tree is read-only and setuptools writes ``build/`` next to it).  ``baseline/_ref`` is git-ignored
(no reference source enters the history) but travels to the GPU box with the snapshot, so the ``-m gpu`` tests and
``bench.py --impl reference`` can run the reference itself there.  ``import_reference()`` puts it on ``sys.path`` and
stubs the three third-party modules the reference imports at module scope that this image lacks (imageio, optuna,
matplotlib -- none is touched on the render_rays path).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
STUBBED = ["imageio", "optuna", "optuna.samplers", "optuna.trial", "optuna.exceptions", "matplotlib", "matplotlib.pyplot",
           "matplotlib.figure", "matplotlib.axes"]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "nerf_sampling", "trainers", "sampling_trainer.py"))


def install(force: bool = False) -> bool:
    """Install the reference into baseline/_ref when its source tree is present (build container only)."""
    if available() and not force:
        return True
    if not os.path.isdir(os.path.join(REF_SRC, "nerf_sampling")):
        return False
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--upgrade", "--target", REF_DIR, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("reference install failed:\n" + res.stdout + res.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return available()


class _Stub(types.ModuleType):
    """Attribute access yields sub-stubs, calls return None; dunder lookups raise so ``inspect`` keeps working."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


def import_reference() -> bool:
    """Make ``import nerf_sampling`` resolve to the reference; False when baseline/_ref is absent."""
    if not available():
        return False
    import importlib

    import torch  # noqa: F401  (must be imported before the stubs: torch probes optional modules by name)

    for name in STUBBED:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import nerf_sampling  # noqa: F401

    return True


def build_reference_trainer(basedir: str, device: str = "cpu", seed: int = 42, **over):
    """``DepthNetTrainer`` of the reference with random-init weights, built the way experiments/run.py builds it
    (seed 42, n_layers=10, layer_width=256, sphere_radius=2; run.py:104-151) -> (trainer, render_kwargs_train, render_kwargs_test)."""
    import torch
    from nerf_sampling.trainers import DepthNetTrainer

    kw = dict(dataset_type="blender", basedir=basedir, expname="exp", no_batching=True, datadir="unused", device=device,
              N_rand=1024, white_bkgd=True, half_res=True, input_dims_embed=3, use_viewdirs=True, N_importance=128,
              N_samples=64, n_layers=10, layer_width=256, sphere_radius=2.0, depth_net_lr=1e-4, train_depth_net_only=True,
              distance=0.1, sampling_mode="uniform", n_depth_samples=32)
    kw.update(over)
    os.makedirs(os.path.join(basedir, kw["expname"]), exist_ok=True)
    torch.manual_seed(seed)
    tr = DepthNetTrainer(**kw)
    _opt, _sopt, rk_train, rk_test = tr.create_nerf_model()
    return tr, rk_train, rk_test


if __name__ == "__main__":
    print("installed" if install(force="--force" in sys.argv) else "reference source tree not present")
