"""fp32 restatement of nerf-sampling's ``render_rays`` hot path (TEST INFRASTRUCTURE).

This module is the parity oracle.  It restates, in plain functional torch
(fp32, device-agnostic: CPU by default, a CUDA device when a test wants the
same arithmetic at full size), what the reference computes on the path

    DepthNet -> sample placement -> positional encoding + NeRF MLP -> raw2outputs

plus the vanilla hierarchical sampler and the training-time render.  Every
function cites the reference ``file:line`` it follows (paths relative to
``/root/reference/nerf_sampling``).  It is pinned against the reference itself:
``tests/golden/make_golden.py`` imports the real reference in the build
container, runs both on the same seeded inputs, asserts bit-equality on CPU and
writes the fixtures that ``tests/test_oracle_golden.py`` re-checks wherever the
reference is absent (the GPU box).

Product code never imports this file.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------
# Scene helpers (camera model used by every benchmark configuration)
# ----------------------------------------------------------------------------

LEGO_CAMERA_ANGLE_X = 0.6911112070083618  # transforms_*.json of the lego scene


def pose_spherical(theta: float, phi: float, radius: float) -> torch.Tensor:
    """Camera-to-world matrix on a sphere (nerf_pytorch/load_blender.py:8-43)."""
    t = torch.tensor(
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]]
    ).float()
    ph = phi / 180.0 * np.pi
    rp = torch.tensor(
        [
            [1, 0, 0, 0],
            [0, np.cos(ph), -np.sin(ph), 0],
            [0, np.sin(ph), np.cos(ph), 0],
            [0, 0, 0, 1],
        ]
    ).float()
    th = theta / 180.0 * np.pi
    rt = torch.tensor(
        [
            [np.cos(th), 0, -np.sin(th), 0],
            [0, 1, 0, 0],
            [np.sin(th), 0, np.cos(th), 0],
            [0, 0, 0, 1],
        ]
    ).float()
    flip = torch.tensor(
        np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]),
        dtype=torch.float32,
    )
    return flip @ (rt @ (rp @ t))


def intrinsics(H: int, W: int, camera_angle_x: float = LEGO_CAMERA_ANGLE_X) -> np.ndarray:
    """focal rule load_blender.py:80-82, K matrix trainers/Trainer.py:142."""
    focal = 0.5 * W / np.tan(0.5 * camera_angle_x)
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])


def get_rays(H: int, W: int, K, c2w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pinhole rays, pixel centres without +0.5 (run_nerf_helpers.py:187-202)."""
    dev = c2w.device
    xs = torch.linspace(0, W - 1, W, device=dev)
    ys = torch.linspace(0, H - 1, H, device=dev)
    i, j = torch.meshgrid(xs, ys, indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack(
        [(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1
    )
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def prepare_rays(H, W, K, c2w=None, rays=None, near=2.0, far=6.0):
    """[o, d, near, far, viewdir] rows for a view or a ray batch
    (nerf_utils.py:156-188 with use_viewdirs=True, ndc=False, no static cam)."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    viewdirs = viewdirs.reshape(-1, 3).float()
    sh = rays_d.shape
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    nr = near * torch.ones_like(rays_d[..., :1])
    fr = far * torch.ones_like(rays_d[..., :1])
    packed = torch.cat([rays_o, rays_d, nr, fr, viewdirs], -1)
    return packed, rays_o, rays_d, sh


# ----------------------------------------------------------------------------
# Random-init weights in the reference's construction order
# ----------------------------------------------------------------------------


def _linear(params: Params, name: str, fan_in: int, fan_out: int) -> None:
    lin = torch.nn.Linear(fan_in, fan_out)  # same default init / RNG draws as the reference
    params[name + ".weight"] = lin.weight.detach().clone()
    params[name + ".bias"] = lin.bias.detach().clone()


def init_nerf(D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,)) -> Params:
    """state_dict of NeRF(use_viewdirs=True) in creation order (run_nerf_helpers.py:67-107)."""
    p: Params = {}
    _linear(p, "pts_linears.0", input_ch, W)
    for i in range(D - 1):
        _linear(p, f"pts_linears.{i + 1}", W + input_ch if i in skips else W, W)
    _linear(p, "views_linears.0", input_ch_views + W, W // 2)
    _linear(p, "feature_linear", W, W)
    _linear(p, "alpha_linear", W, 1)
    _linear(p, "rgb_linear", W // 2, 3)
    return p


def init_depthnet(hidden: List[int], cat_hidden: List[int], multires: int = 10) -> Params:
    """state_dict of DepthNet in creation order (depth_nets/depth_net.py:37-108)."""
    d3 = 3 + 3 * 2 * multires  # 63
    d6 = 6 + 6 * 2 * multires  # 126
    p: Params = {}
    _linear(p, "origin_layers.0", 2 * d3, hidden[0])
    _linear(p, "direction_layers.0", 2 * d3, hidden[0])
    _linear(p, "intersection_layers.0", 2 * d6, hidden[0])
    for i, size in enumerate(hidden[:-1]):
        _linear(p, f"origin_layers.{i + 1}", size + d3, hidden[i + 1])
        _linear(p, f"direction_layers.{i + 1}", size + d3, hidden[i + 1])
    for i, size in enumerate(hidden[:-1]):
        _linear(p, f"intersection_layers.{i + 1}", size + d6, hidden[i + 1])
    _linear(p, "cat_layers.0", hidden[-1] * 3 + d3 + d3 + d6, cat_hidden[0])
    for i, size in enumerate(cat_hidden[:-1]):
        _linear(p, f"cat_layers.{2 * (i + 1)}", size, cat_hidden[i + 1])  # odd slots are LeakyReLU
    _linear(p, "to_depth.0", cat_hidden[-1], 1)
    return p


def init_models(seed: int = 42, n_layers: int = 10, width: int = 256, fine: bool = True):
    """Coarse NeRF, fine NeRF, DepthNet -- the order DepthNetTrainer.create_nerf_model
    builds them (nerf_utils.py:393-430 then trainers/sampling_trainer.py:67-74),
    after ``torch.manual_seed(seed)`` as experiments/run.py:111 does."""
    torch.manual_seed(seed)
    coarse = init_nerf()
    fine_p = init_nerf() if fine else None
    dn = init_depthnet([width] * n_layers, [width] * n_layers)
    return coarse, fine_p, dn


def params_to(params: Optional[Params], device) -> Optional[Params]:
    return None if params is None else {k: v.to(device) for k, v in params.items()}


# ----------------------------------------------------------------------------
# Operators
# ----------------------------------------------------------------------------


def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]
    (run_nerf_helpers.py:15-63; log-sampled bands are exact powers of two)."""
    bands = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires, device=x.device)
    out = [x]
    for f in bands:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def nerf_forward(p: Params, x: torch.Tensor, input_ch=63, input_ch_views=27, skips=(4,)) -> torch.Tensor:
    """8x256 trunk, skip-cat after layer 4, view branch (run_nerf_helpers.py:109-134)."""
    pts, views = torch.split(x, [input_ch, input_ch_views], dim=-1)
    h = pts
    n_trunk = len([k for k in p if k.startswith("pts_linears.") and k.endswith(".weight")])
    for i in range(n_trunk):
        h = F.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([pts, h], -1)
    alpha = F.linear(h, p["alpha_linear.weight"], p["alpha_linear.bias"])
    feature = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])
    h = torch.cat([feature, views], -1)
    h = F.relu(F.linear(h, p["views_linears.0.weight"], p["views_linears.0.bias"]))
    rgb = F.linear(h, p["rgb_linear.weight"], p["rgb_linear.bias"])
    return torch.cat([rgb, alpha], -1)


def run_network(pts: torch.Tensor, viewdirs: torch.Tensor, p: Params, netchunk: int = 1024 * 64) -> torch.Tensor:
    """Encode + query in netchunk slices (trainers/Trainer.py:789-806, nerf_utils.py:45-55)."""
    flat = pts.reshape(-1, pts.shape[-1])
    emb = embed(flat, 10)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    emb = torch.cat([emb, embed(dirs, 4)], -1)
    outs = [nerf_forward(p, emb[i : i + netchunk]) for i in range(0, emb.shape[0], netchunk)]
    out = torch.cat(outs, 0)
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])


def solve_quadratic(a, b, c):
    """Roots [(-b-sqrt(D))/2a, (-b+sqrt(D))/2a], NaN when D<0 (nerf_pytorch/utils.py:159-179)."""
    delta = b**2 - 4 * a * c
    pm = torch.stack([torch.ones_like(delta), -torch.ones_like(delta)])
    return (-b - (pm * torch.sqrt(delta))) / (2 * a)


def sphere_intersections(origin, direction, radius: torch.Tensor):
    """Line/sphere hits, centre 0 (nerf_pytorch/utils.py:182-217). Returns t [n,2], pts [n,2,3]."""
    b = 2 * (direction * origin).sum(dim=1)
    c = torch.norm(origin, dim=1) ** 2 - radius**2  # (reference writes radius.T: a no-op on a 1-D tensor)
    a = (direction * direction).sum(dim=1)
    t = solve_quadratic(a, b, c).T
    pts = origin.unsqueeze(1) + t.unsqueeze(2) * direction.unsqueeze(1)
    return t, pts


def depthnet_forward(p: Params, rays_o, rays_d, sphere_radius=2.0, near=2, far=6, multires=10):
    """DepthNet.forward (depth_nets/depth_net.py:117-169).

    The three branches apply NO activation: ``nn.LeakyReLU(x)`` at :140,:148,:156
    builds a module and drops it.  cat_layers are Linear+LeakyReLU(0.01)."""
    radius = torch.tensor([sphere_radius], device=rays_o.device)
    e_o = embed(rays_o, multires)
    e_d = embed(rays_d, multires)
    _, hits = sphere_intersections(rays_o, rays_d, radius)
    e_i = embed(torch.flatten(hits, start_dim=1), multires)

    def branch(name, e):
        x = e
        i = 0
        while f"{name}.{i}.weight" in p:
            x = F.linear(torch.cat([x, e], -1), p[f"{name}.{i}.weight"], p[f"{name}.{i}.bias"])
            i += 1
        return x

    x = torch.cat(
        [branch("origin_layers", e_o), branch("direction_layers", e_d), branch("intersection_layers", e_i), e_o, e_d, e_i],
        -1,
    )
    i = 0
    while f"cat_layers.{i}.weight" in p:
        x = F.leaky_relu(F.linear(x, p[f"cat_layers.{i}.weight"], p[f"cat_layers.{i}.bias"]), 0.01)
        i += 2
    s = torch.sigmoid(F.linear(x, p["to_depth.0.weight"], p["to_depth.0.bias"]))
    return near * (1 - s) + far * s


def place_samples(rays_o, rays_d, mean, n_samples=32, mode="gaussian", std=0.1, noise=None):
    """sample_points_around_mean (nerf_pytorch/utils.py:220-244).

    ``noise`` ([N, n_samples-1], standard normal) replaces the reference's
    ``torch.randn`` draw so gaussian mode is reproducible; uniform mode clips to
    [2, 6], gaussian does not."""
    if mode == "depth_only":
        z = mean
    elif mode == "gaussian":
        if noise is None:
            noise = torch.randn(mean.shape[0], n_samples - 1, device=mean.device)
        z, _ = torch.cat([mean + std * noise, mean], dim=-1).sort(dim=-1)
    elif mode == "uniform":
        grid = torch.linspace(-std, std, steps=n_samples - 1, device=mean.device)
        z, _ = torch.cat([mean + grid.view(1, -1).expand(mean.size(0), -1), mean], dim=-1).sort(dim=-1)
        z = torch.clip(z, 2, 6)
    else:
        raise ValueError(mode)
    return rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None], z


def raw2alpha(raw, dists):
    """1 - exp(-relu(sigma) * delta) (nerf_utils.py:27-42)."""
    return 1.0 - torch.exp(-F.relu(raw) * dists)


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0.0, white_bkgd=True, noise=None):
    """Alpha compositing (trainers/sampling_trainer.py:153-230).

    Returns (rgb_map, disp_map, acc_map, depth_map, density, alphas, weights)."""
    dev = raw.device
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.tensor([1e10], device=dev).expand(dists[..., :1].shape)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    nz = 0.0
    if raw_noise_std > 0.0:
        nz = (torch.randn(raw[..., 3].shape, device=dev) if noise is None else noise) * raw_noise_std
    density = raw[..., 3]
    alphas = raw2alpha(density + nz, dists)
    trans = torch.cumprod(
        torch.cat([torch.ones((alphas.shape[0], 1), device=dev), 1.0 - alphas + 1e-10], -1), -1
    )[:, :-1]
    weights = alphas * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / (torch.sum(weights, -1) + 1e-10))
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    if weights.shape[-1] == 0:
        # S == 1 quirk (:220-221): the 1e10 pad is sized from the EMPTY slice dists[..., :1], so
        # dists/alphas/weights are [N,0], acc = depth = 0, disp = 1e10 and the colour is the plain
        # sum of sigmoid(rgb) over the single sample.  This is what the training render returns.
        rgb_map = torch.sum(rgb, -2)
    return rgb_map, disp_map, acc_map, depth_map, density, alphas, weights


def sample_pdf(bins, weights, n_samples, det=True, u=None, return_inds=False):
    """Inverse-CDF sampling (run_nerf_helpers.py:250-293).  ``u`` overrides the draw."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, steps=n_samples, device=bins.device)
            u = u.expand(list(cdf.shape[:-1]) + [n_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [n_samples], device=bins.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])
    if return_inds:
        return samples, inds
    return samples


def coarse_z(near, far, n_rays, n_samples=64, lindisp=True, t_rand=None):
    """Stratified coarse depths (trainers/Trainer.py:603-627); lindisp=True is the
    reference default (Trainer.py:49).  ``t_rand`` ([N,S] in [0,1)) enables the jitter."""
    t = torch.linspace(0.0, 1.0, steps=n_samples, device=near.device)
    if not lindisp:
        z = near * (1.0 - t) + far * t
    else:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    z = z.expand([n_rays, n_samples])
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def hierarchical(packed, coarse: Params, fine: Optional[Params], n_samples=64, n_importance=128,
                 lindisp=True, white_bkgd=True, raw_noise_std=0.0):
    """Vanilla NeRF coarse + fine pass, perturb=0 (nerf_utils.py:497-611 ->
    trainers/Trainer.py:579-710).  Returns a dict of every intermediate."""
    rays_o, rays_d, viewdirs = packed[:, 0:3], packed[:, 3:6], packed[:, -3:]
    bounds = packed[..., 6:8].reshape(-1, 1, 2)
    near, far = bounds[..., 0], bounds[..., 1]
    n = packed.shape[0]
    zc = coarse_z(near, far, n, n_samples, lindisp)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * zc[..., :, None]
    raw_c = run_network(pts, viewdirs, coarse)
    rgb_c, disp_c, acc_c, _, _, _, w_c = raw2outputs(raw_c, zc, rays_d, raw_noise_std, white_bkgd)
    mid = 0.5 * (zc[..., 1:] + zc[..., :-1])
    z_samp, inds = sample_pdf(mid, w_c[..., 1:-1], n_importance, det=True, return_inds=True)
    z_samp = z_samp.detach()
    zf, _ = torch.sort(torch.cat([zc, z_samp], -1), -1)
    pts_f = rays_o[..., None, :] + rays_d[..., None, :] * zf[..., :, None]
    raw_f = run_network(pts_f, viewdirs, coarse if fine is None else fine)
    rgb_f, disp_f, acc_f, depth_f, dens_f, alphas_f, w_f = raw2outputs(raw_f, zf, rays_d, raw_noise_std, white_bkgd)
    return dict(z_coarse=zc, raw_coarse=raw_c, weights_coarse=w_c, rgb_coarse=rgb_c, disp_coarse=disp_c,
                z_samples=z_samp, inds=inds, z_fine=zf, pts_fine=pts_f, raw_fine=raw_f, rgb_fine=rgb_f,
                disp_fine=disp_f, acc_fine=acc_f, depth_fine=depth_f, density_fine=dens_f,
                alphas_fine=alphas_f, weights_fine=w_f)


def render_rays_test(packed, coarse: Params, fine: Optional[Params], dn: Params, *, n_depth_samples=32,
                     sampling_mode="uniform", distance=0.1, mode="depthnet", noise=None,
                     n_samples=64, n_importance=128, lindisp=True, compare_nerf=False) -> Dict[str, torch.Tensor]:
    """Inference render of one ray chunk (nerf_utils.py:736-876).

    ``mode``: "depthnet" (default path), "full_nerf" (use_full_nerf), "max_pts"
    (use_nerf_max_pts).  The DepthNet path calls raw2outputs with misspelled
    kwargs (:858-865), i.e. zero noise and white background -- restated here."""
    rays_o, rays_d, viewdirs = packed[:, 0:3], packed[:, 3:6], packed[:, -3:]
    ret: Dict[str, torch.Tensor] = {}
    if mode in ("full_nerf", "max_pts") or compare_nerf:
        h = hierarchical(packed, coarse, fine, n_samples, n_importance, lindisp)
        top = h["weights_fine"].argmax(dim=1, keepdim=True)
        max_z = torch.gather(h["z_fine"], 1, top)
        max_w = torch.gather(h["weights_fine"], 1, top)
        rgb = torch.sigmoid(h["raw_fine"][..., :3])
        max_rgb = torch.gather(rgb, 1, top.unsqueeze(-1).expand(-1, 1, 3)).squeeze()
        max_pts = rays_o[..., None, :] + rays_d[..., None, :] * max_z[..., :, None]
        ret.update(max_z_vals=max_z, max_pts=max_pts, max_weights=max_w, top_indices=top)
    if mode == "max_pts":
        rgb_map, disp, w, pts, z = max_rgb, torch.zeros_like(max_rgb), max_w, max_pts, max_z
    elif mode == "full_nerf":
        rgb_map, disp, w, pts, z = h["rgb_fine"], h["disp_fine"], h["weights_fine"], h["pts_fine"], h["z_fine"]
    else:
        z_mean = depthnet_forward(dn, rays_o, rays_d)
        pts, z = place_samples(rays_o, rays_d, z_mean, n_depth_samples, sampling_mode, distance, noise)
        raw = run_network(pts, viewdirs, coarse if fine is None else fine)
        rgb_map, disp, acc, depth, _, alphas, w = raw2outputs(raw, z, rays_d, 0.0, True)
        ret.update(z_mean=z_mean, raw=raw, acc_map=acc, depth_map=depth, alphas=alphas)
    ret.update(depth_net_rgb_map=rgb_map, depth_net_weights=w, depth_net_disp_map=disp,
               depth_net_z_vals=z, depth_net_pts=pts)
    return ret


def render_rays_train(packed, coarse: Params, fine: Optional[Params], dn: Params, *, n_samples=64,
                      n_importance=128, lindisp=True, white_bkgd=True) -> Dict[str, torch.Tensor]:
    """Training render, perturb=0 (nerf_utils.py:614-733): hierarchical target + DepthNet, 1 sample/ray."""
    rays_o, rays_d, viewdirs = packed[:, 0:3], packed[:, 3:6], packed[:, -3:]
    h = hierarchical(packed, coarse, fine, n_samples, n_importance, lindisp, white_bkgd)
    top = h["weights_fine"].argmax(dim=1, keepdim=True)
    max_z = torch.gather(h["z_fine"], 1, top)
    z_dn = depthnet_forward(dn, rays_o, rays_d)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_dn[..., :, None]
    raw = run_network(pts, viewdirs, coarse if fine is None else fine)
    rgb_map, disp, *_ = raw2outputs(raw, z_dn, rays_d, 0.0, True)
    return dict(depth_net_rgb_map=rgb_map, depth_net_disp_map=disp, depth_net_z_vals=z_dn, max_z_vals=max_z,
                depth_net_pts=pts, max_pts=rays_o[..., None, :] + rays_d[..., None, :] * max_z[..., :, None],
                raw=raw, top_indices=top)


def render_view(H, W, K, c2w, coarse, fine, dn, chunk=1024 * 32, **kw) -> Dict[str, torch.Tensor]:
    """render_test for one camera (nerf_utils.py:191-255): chunk loop + reshape to [H,W,...]."""
    packed, rays_o, rays_d, sh = prepare_rays(H, W, K, c2w=c2w)
    outs: Dict[str, List[torch.Tensor]] = {}
    for i in range(0, packed.shape[0], chunk):
        r = render_rays_test(packed[i : i + chunk], coarse, fine, dn, **kw)
        for k, v in r.items():
            outs.setdefault(k, []).append(v)
    res = {k: torch.cat(v, 0) for k, v in outs.items()}
    res = {k: v.reshape(list(sh[:-1]) + list(v.shape[1:])) for k, v in res.items()}
    res["rays_o"], res["rays_d"] = rays_o, rays_d
    return res


def psnr(img: torch.Tensor, target: torch.Tensor) -> float:
    """-10 log10(mse) (run_nerf_helpers.py:9-10)."""
    return float(-10.0 * math.log10(float(torch.mean((img - target) ** 2))))
