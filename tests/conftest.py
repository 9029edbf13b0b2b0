import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def weights_checksum(*dicts) -> float:
    tot = 0.0
    for d in dicts:
        for k in sorted(d):
            tot += float(d[k].double().abs().sum())
    return tot


def load_golden(name: str):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))


@pytest.fixture(scope="session")
def oracle_models():
    """Coarse NeRF, fine NeRF, DepthNet params re-created from seed 42; the checksum stored in the fixtures
    proves they are the weights the reference itself initialised when the goldens were generated."""
    from oracle import nerf_oracle as O

    coarse, fine, dn = O.init_models(42)
    want = float(load_golden("g1")["weights_checksum"])
    got = weights_checksum(coarse, fine, dn)
    assert abs(got - want) <= 1e-9 * want, "seeded init no longer reproduces the reference's weights"
    return coarse, fine, dn


@pytest.fixture(scope="session")
def lib():
    import nerf_sampling_b200 as pkg

    pkg.build()
    from nerf_sampling_b200 import _lib

    return _lib.lib()


@pytest.fixture(scope="session", params=["fast", "split"])
def b200_models(request, oracle_models):
    """The product's module shells loaded with the same weights, on the GPU, once per NeRF precision mode:
    "fast" (fp16 single pass + split-precision guard band, the default) and "split" (bf16 hi+lo everywhere)."""
    from nerf_sampling_b200.depth_nets import DepthNet
    from nerf_sampling_b200.nerf_pytorch.run_nerf_helpers import NeRF
    from nerf_sampling_b200.packing import PREC_FAST, PREC_SPLIT

    coarse, fine, dn = oracle_models
    dev = torch.device("cuda")

    def nerf(sd):
        m = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        m.load_state_dict(sd)
        m.precision = PREC_FAST if request.param == "fast" else PREC_SPLIT
        return m.to(dev)

    d = DepthNet(hidden_sizes=[256] * 10, cat_hidden_sizes=[256] * 10, sphere_radius=2.0)
    d.load_state_dict(dn)
    return nerf(coarse), nerf(fine), d.to(dev)
