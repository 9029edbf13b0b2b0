"""The oracle against the fixtures the reference itself produced (tests/golden/make_golden.py).

make_golden.py asserts BIT-equality oracle == reference in the build container.  Here (any machine, no reference)
the tolerance is a few ulp because a different host CPU may pick a different sgemm kernel."""

import numpy as np
import torch

from conftest import load_golden
from oracle import nerf_oracle as O

TOL = 2e-5


def t(x):
    return torch.from_numpy(np.asarray(x))


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape
    assert bool((((a - b).abs() <= tol) | (torch.isnan(a) & torch.isnan(b))).all()), float((a - b).abs().max())


def test_g1_tiny_view_all_intermediates(oracle_models):
    coarse, fine, dn = oracle_models
    g = load_golden("g1")
    H, W, S = int(g["H"]), int(g["W"]), int(g["S"])
    c2w = O.pose_spherical(30.0, -30.0, 4.0)[:3, :4]
    with torch.no_grad():
        o = O.render_view(H, W, O.intrinsics(H, W), c2w, coarse, fine, dn, n_depth_samples=S, sampling_mode="uniform",
                          distance=float(g["distance"]))
    close(o["rays_o"], t(g["rays_o"]), 0)
    close(o["rays_d"], t(g["rays_d"]), 1e-7)
    close(o["z_mean"].reshape(-1, 1), t(g["z_mean"]))
    close(o["depth_net_z_vals"], t(g["z"]))
    close(o["depth_net_pts"], t(g["pts"]), 1e-4)
    close(o["raw"].reshape(-1, S, 4), t(g["raw"]), 1e-4)
    close(o["depth_net_weights"], t(g["weights"]), 1e-4)
    close(o["depth_net_rgb_map"], t(g["rgb"]), 1e-4)


def test_g2_config1_view(oracle_models):
    """BASELINE config #1: 200x200, DepthNet + 32 uniform samples."""
    coarse, fine, dn = oracle_models
    g = load_golden("g2")
    H, W = int(g["H"]), int(g["W"])
    c2w = O.pose_spherical(30.0, -30.0, 4.0)[:3, :4]
    with torch.no_grad():
        o = O.render_view(H, W, O.intrinsics(H, W), c2w, coarse, fine, dn, n_depth_samples=32, sampling_mode="uniform", distance=0.1)
    close(o["z_mean"].reshape(H, W), t(g["z_mean"]))
    # the last interval is 1e10 long, so alpha_last is a step function of sign(sigma_last): skip rays whose
    # sigma_last is within round-off of zero (none on the machine that generated the fixture)
    safe = torch.from_numpy(np.abs(g["sigma_last"]) > 1e-5).reshape(H, W)
    assert safe.float().mean() > 0.999
    d = (o["depth_net_rgb_map"] - t(g["rgb"])).abs().max(-1).values
    assert float(d[safe].max()) <= 1e-4


def test_g3_placement(oracle_models):
    g = load_golden("g3")
    ro, rd, mean = t(g["rays_o"]), t(g["rays_d"]), t(g["mean"])
    pts, z = O.place_samples(ro, rd, mean, 13, "gaussian", 0.3, noise=t(g["noise"]))
    close(z, t(g["z_gauss"]), 0)
    close(pts, t(g["pts_gauss"]), 0)
    for S in (1, 2, 3, 13, 32, 33, 64):
        _, z = O.place_samples(ro, rd, mean, S, "depth_only" if S == 1 else "uniform", 0.25)
        close(z, t(g[f"z_uniform_{S}"]), 0)


def test_g4_hierarchical(oracle_models):
    coarse, fine, dn = oracle_models
    g = load_golden("g4")
    packed, *_ = O.prepare_rays(int(g["H"]), int(g["W"]), None, rays=(t(g["rays_o"]), t(g["rays_d"])))
    with torch.no_grad():
        h = O.hierarchical(packed, coarse, fine)
    close(h["z_coarse"], t(g["z_coarse"]), 0)
    close(h["weights_coarse"], t(g["weights_coarse"]), 1e-4)
    close(h["z_samples"], t(g["z_samples"]), 1e-3)
    close(h["rgb_fine"], t(g["rgb"]), 1e-3)
    # sample_pdf on the reference's own (bins, weights): integer indices must be bit-exact
    zc, wc = t(g["z_coarse"]), t(g["weights_coarse"])
    mid = 0.5 * (zc[..., 1:] + zc[..., :-1])
    zs, inds = O.sample_pdf(mid, wc[..., 1:-1], 128, det=True, return_inds=True)
    assert torch.equal(inds, t(g["inds"]))
    close(zs, t(g["z_samples"]), 1e-6)


def test_g5_training_render(oracle_models):
    coarse, fine, dn = oracle_models
    g = load_golden("g5")
    packed, *_ = O.prepare_rays(800, 800, None, rays=(t(g["rays_o"]), t(g["rays_d"])))
    with torch.no_grad():
        o = O.render_rays_train(packed, coarse, fine, dn)
    close(o["depth_net_z_vals"], t(g["z_dn"]))
    close(o["max_z_vals"], t(g["max_z"]), 1e-3)
    close(o["depth_net_rgb_map"], t(g["rgb"]), 1e-4)
    loss = torch.mean((o["depth_net_rgb_map"] - t(g["target"])) ** 2)
    assert abs(float(loss) - float(g["img_loss"])) < 1e-6


def test_raw2outputs_single_sample_quirk():
    """S == 1: empty per-sample tensors, colour = sigmoid(rgb), disp = 1e10 (sampling_trainer.py:178-180,:220-221)."""
    raw = torch.tensor([[[0.3, -0.2, 1.0, 5.0]]])
    rgb, disp, acc, depth, dens, alphas, w = O.raw2outputs(raw, torch.tensor([[3.0]]), torch.tensor([[0.0, 0.0, 1.0]]))
    assert w.shape == (1, 0) and alphas.shape == (1, 0)
    assert torch.allclose(rgb, torch.sigmoid(raw[:, 0, :3]))
    assert float(acc) == 0.0 and float(depth) == 0.0 and float(disp) == 1e10
