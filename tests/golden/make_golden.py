"""Generate the golden fixtures by running the REAL reference (build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/nerf_sampling`` unmodified (stubbing the absent
imageio / optuna / matplotlib modules, none of which is touched on the hot
path), builds ``DepthNetTrainer`` with random-init weights under
``torch.manual_seed(42)`` exactly as ``experiments/run.py`` does, runs the
reference's own ``render_test`` / ``render`` / operators on seeded synthetic
rays on CPU, asserts that ``oracle/nerf_oracle.py`` reproduces every output
BIT-EXACTLY, and writes the reference outputs to ``tests/golden/*.npz``.

The GPU box has no ``/root/reference``; there the fixtures (plus weights
re-created from the seed, guarded by a checksum stored in the fixture) pin the
oracle.
"""

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


for _m in ["imageio", "optuna", "optuna.samplers", "optuna.trial", "optuna.exceptions", "matplotlib",
           "matplotlib.pyplot", "matplotlib.figure", "matplotlib.axes"]:
    sys.modules[_m] = _Stub(_m)

from nerf_sampling.nerf_pytorch import nerf_utils, run_nerf_helpers, utils as ref_utils  # noqa: E402
from nerf_sampling.nerf_pytorch.load_blender import pose_spherical as ref_pose_spherical  # noqa: E402
from nerf_sampling.trainers import DepthNetTrainer  # noqa: E402

from oracle import nerf_oracle as O  # noqa: E402


def weights_checksum(*dicts) -> float:
    tot = 0.0
    for d in dicts:
        for k in sorted(d):
            tot += float(d[k].double().abs().sum())
    return tot


def build_trainer(tmp, **over):
    kw = dict(dataset_type="blender", basedir=tmp, expname="exp", no_batching=True, datadir="unused", device="cpu",
              N_rand=1024, white_bkgd=True, half_res=True, input_dims_embed=3, use_viewdirs=True, N_importance=128,
              N_samples=64, n_layers=10, layer_width=256, sphere_radius=2.0, depth_net_lr=1e-4,
              train_depth_net_only=True, distance=0.1, sampling_mode="uniform", n_depth_samples=32)
    kw.update(over)
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    torch.manual_seed(42)
    tr = DepthNetTrainer(**kw)
    opt, sopt, rk_train, rk_test = tr.create_nerf_model()
    return tr, rk_train, rk_test


def same(a, b, name):
    a = a.detach() if torch.is_tensor(a) else torch.as_tensor(a)
    b = b.detach() if torch.is_tensor(b) else torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    ok = torch.equal(a, b) or bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    assert ok, f"oracle != reference for {name}: max abs diff {(a.double() - b.double()).abs().max()}"


def np_(x):
    return x.detach().cpu().numpy()


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    with tempfile.TemporaryDirectory() as tmp, torch.no_grad():
        tr, rk_train, rk_test = build_trainer(tmp)
        ref_coarse = {k: v.detach() for k, v in rk_test["network_fn"].state_dict().items()}
        ref_fine = {k: v.detach() for k, v in rk_test["network_fine"].state_dict().items()}
        ref_dn = {k: v.detach() for k, v in rk_test["depth_network"].state_dict().items()}
        coarse, fine, dn = O.init_models(42)
        for a, b, n in [(ref_coarse, coarse, "coarse"), (ref_fine, fine, "fine"), (ref_dn, dn, "depthnet")]:
            assert sorted(a.keys()) == sorted(b.keys()), n
            for k in a:
                same(a[k], b[k], f"{n}.{k}")
        csum = weights_checksum(coarse, fine, dn)
        print("weights identical to the reference's; checksum", csum)

        # ---- pose / intrinsics ------------------------------------------------
        c2w = ref_pose_spherical(30.0, -30.0, 4.0)
        same(c2w, O.pose_spherical(30.0, -30.0, 4.0), "pose_spherical")

        def ref_view(H, W, **trainer_over):
            for k, v in trainer_over.items():
                setattr(tr, k, v)
            K = O.intrinsics(H, W)
            rgb, disp, ex = nerf_utils.render_test(H, W, K, chunk=tr.chunk, c2w=c2w[:3, :4], **rk_test)
            return K, rgb, disp, ex

        # ---- G1: tiny view with every intermediate ---------------------------
        H = W = 16
        K, rgb, disp, ex = ref_view(H, W, n_depth_samples=8, sampling_mode="uniform", distance=0.1)
        o = O.render_view(H, W, K, c2w[:3, :4], coarse, fine, dn, n_depth_samples=8, sampling_mode="uniform", distance=0.1)
        same(rgb, o["depth_net_rgb_map"], "g1.rgb")
        same(disp, o["depth_net_disp_map"], "g1.disp")
        same(ex["depth_net_z_vals"], o["depth_net_z_vals"], "g1.z")
        same(ex["depth_net_pts"], o["depth_net_pts"], "g1.pts")
        same(ex["depth_net_weights"], o["depth_net_weights"], "g1.weights")
        same(ex["rays_o"], o["rays_o"], "g1.rays_o")
        same(ex["rays_d"], o["rays_d"], "g1.rays_d")
        # operator-level references for the same rays
        ro, rd = ex["rays_o"], ex["rays_d"]
        z_mean = rk_test["depth_network"](ro, rd)
        same(z_mean, O.depthnet_forward(dn, ro, rd), "g1.z_mean")
        vd = rd / torch.norm(rd, dim=-1, keepdim=True)
        raw = rk_test["network_query_fn"](ex["depth_net_pts"].reshape(-1, 8, 3), vd, rk_test["network_fine"])
        same(raw, O.run_network(ex["depth_net_pts"].reshape(-1, 8, 3), vd, fine), "g1.raw")
        r2o = tr.raw2outputs(raw, ex["depth_net_z_vals"].reshape(-1, 8), rd)
        o2o = O.raw2outputs(raw, ex["depth_net_z_vals"].reshape(-1, 8), rd)
        for a, b, n in zip(r2o, o2o, ["rgb", "disp", "acc", "depth", "density", "alphas", "weights"]):
            same(a, b, "g1.raw2outputs." + n)
        out["g1"] = dict(H=H, W=W, S=8, distance=0.1, rays_o=np_(ro), rays_d=np_(rd), z_mean=np_(z_mean),
                         z=np_(ex["depth_net_z_vals"]), pts=np_(ex["depth_net_pts"]), raw=np_(raw),
                         weights=np_(ex["depth_net_weights"]), rgb=np_(rgb), disp=np_(disp),
                         acc=np_(r2o[2]), depth=np_(r2o[3]), alphas=np_(r2o[5]))

        # ---- G2: BASELINE config #1 (200x200, 32 uniform samples) ------------
        H = W = 200
        K, rgb, disp, ex = ref_view(H, W, n_depth_samples=32, sampling_mode="uniform", distance=0.1)
        o = O.render_view(H, W, K, c2w[:3, :4], coarse, fine, dn, n_depth_samples=32, sampling_mode="uniform", distance=0.1)
        same(rgb, o["depth_net_rgb_map"], "g2.rgb")
        same(disp, o["depth_net_disp_map"], "g2.disp")
        same(ex["depth_net_z_vals"], o["depth_net_z_vals"], "g2.z")
        out["g2"] = dict(H=H, W=W, S=32, distance=0.1, rgb=np_(rgb), disp=np_(disp),
                         z_mean=np_(o["z_mean"]).reshape(H, W), sigma_last=np_(o["raw"][..., -1, 3]))

        # ---- G3: gaussian placement with host-supplied noise + odd sizes -----
        g = torch.Generator().manual_seed(7)
        n = 37
        ro3 = c2w[:3, 3].expand(n, 3).contiguous()
        rd3 = ex["rays_d"].reshape(-1, 3)[torch.randperm(H * W, generator=g)[:n]].contiguous()
        mean3 = rk_test["depth_network"](ro3, rd3)
        noise = torch.randn(n, 12, generator=g)
        torch.manual_seed(123)
        # reference draws torch.randn internally: reproduce by re-seeding and drawing the same tensor
        torch.manual_seed(123)
        pts_ref, z_ref = ref_utils.sample_points_around_mean(ro3, rd3, mean3, n_samples=13, mode="gaussian", std=0.3)
        torch.manual_seed(123)
        noise_ref = torch.randn(n, 12)
        pts_o, z_o = O.place_samples(ro3, rd3, mean3, 13, "gaussian", 0.3, noise=noise_ref)
        same(z_ref, z_o, "g3.z")
        same(pts_ref, pts_o, "g3.pts")
        zs = {}
        for S in (1, 2, 3, 13, 32, 33, 64):
            mode = "depth_only" if S == 1 else "uniform"
            pr, zr = ref_utils.sample_points_around_mean(ro3, rd3, mean3, n_samples=S, mode=mode, std=0.25)
            po, zo = O.place_samples(ro3, rd3, mean3, S, mode, 0.25)
            same(zr, zo, f"g3.uniform{S}.z")
            same(pr, po, f"g3.uniform{S}.pts")
            zs[f"z_uniform_{S}"] = np_(zr)
        out["g3"] = dict(rays_o=np_(ro3), rays_d=np_(rd3), mean=np_(mean3), noise=np_(noise_ref), z_gauss=np_(z_ref),
                         pts_gauss=np_(pts_ref), **zs)

        # ---- G4: vanilla hierarchical (config #4 shape), 24x24 ---------------
        H = W = 24
        K, rgb, disp, ex = ref_view(H, W, use_full_nerf=True)
        tr.use_full_nerf = False
        packed, ro4, rd4, _ = O.prepare_rays(H, W, K, c2w=c2w[:3, :4])
        h = O.hierarchical(packed, coarse, fine)
        same(rgb.reshape(-1, 3), h["rgb_fine"], "g4.rgb")
        same(disp.reshape(-1), h["disp_fine"], "g4.disp")
        same(ex["depth_net_z_vals"].reshape(-1, 192), h["z_fine"], "g4.z_fine")
        same(ex["depth_net_weights"].reshape(-1, 192), h["weights_fine"], "g4.weights_fine")
        # operator-level sample_pdf reference
        mid = 0.5 * (h["z_coarse"][..., 1:] + h["z_coarse"][..., :-1])
        zs_ref = run_nerf_helpers.sample_pdf(mid, h["weights_coarse"][..., 1:-1], 128, det=True)
        same(zs_ref, h["z_samples"], "g4.z_samples")
        out["g4"] = dict(H=H, W=W, rays_o=np_(ro4), rays_d=np_(rd4), z_coarse=np_(h["z_coarse"]),
                         weights_coarse=np_(h["weights_coarse"]), z_samples=np_(h["z_samples"]),
                         inds=np_(h["inds"]).astype(np.int64), z_fine=np_(h["z_fine"]),
                         weights_fine=np_(h["weights_fine"]), rgb=np_(rgb).reshape(-1, 3), disp=np_(disp).reshape(-1),
                         raw_coarse=np_(h["raw_coarse"]))

        # compare_nerf / use_nerf_max_pts modes
        K, rgb_m, disp_m, ex_m = ref_view(H, W, use_nerf_max_pts=True)
        tr.use_nerf_max_pts = False
        om = O.render_rays_test(packed, coarse, fine, dn, mode="max_pts")
        same(rgb_m.reshape(-1, 3), om["depth_net_rgb_map"], "g4.max_rgb")
        same(ex_m["max_z_vals"].reshape(-1, 1), om["max_z_vals"], "g4.max_z")
        out["g4"].update(max_rgb=np_(rgb_m).reshape(-1, 3), max_z=np_(ex_m["max_z_vals"]).reshape(-1, 1),
                         top_indices=np_(om["top_indices"]).astype(np.int64))

    # ---- G5: training render (config #5 shape, 96 rays, perturb=0) ----------
    with tempfile.TemporaryDirectory() as tmp:
        tr, rk_train, rk_test = build_trainer(tmp, perturb=0.0)
        rk_train["perturb"] = 0.0
        H = W = 800
        K = O.intrinsics(H, W)
        tr.H, tr.W, tr.K = H, W, K
        c2w = ref_pose_spherical(30.0, -30.0, 4.0)
        with torch.no_grad():
            ro_all, rd_all = run_nerf_helpers.get_rays(H, W, K, c2w[:3, :4])
        sel = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:96]
        ro5 = ro_all.reshape(-1, 3)[sel].contiguous()
        rd5 = rd_all.reshape(-1, 3)[sel].contiguous()
        target = torch.rand(96, 3, generator=torch.Generator().manual_seed(1))
        batch_rays = torch.stack([ro5, rd5], 0)
        for m in (rk_train["network_fn"], rk_train["network_fine"]):
            ref_utils.freeze_model(m)
        rgb, disp, ex = nerf_utils.render(H, W, K, chunk=tr.chunk, rays=batch_rays, retraw=True, **rk_train)
        with torch.no_grad():
            packed5, *_ = O.prepare_rays(H, W, K, rays=(ro5, rd5))
            coarse, fine, dn = O.init_models(42)
            ot = O.render_rays_train(packed5, coarse, fine, dn)
        same(rgb, ot["depth_net_rgb_map"], "g5.rgb")
        same(ex["max_z_vals"], ot["max_z_vals"], "g5.max_z")
        same(ex["depth_net_z_vals"], ot["depth_net_z_vals"], "g5.z_dn")
        img_loss = torch.mean((rgb - target) ** 2)
        dn_loss = torch.nn.functional.mse_loss(ex["depth_net_z_vals"], ex["max_z_vals"])
        dn_loss.backward(retain_graph=True)
        img_loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in rk_train["depth_network"].named_parameters()}
        gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
        out["g5"] = dict(rays_o=np_(ro5), rays_d=np_(rd5), target=np_(target), rgb=np_(rgb), disp=np_(disp),
                         max_z=np_(ex["max_z_vals"]), z_dn=np_(ex["depth_net_z_vals"]),
                         img_loss=np.float64(img_loss.item()), dn_loss=np.float64(dn_loss.item()),
                         grad_norm=np.float64(gnorm),
                         grad_to_depth_w=np_(grads["to_depth.0.weight"]),
                         grad_cat0_b=np_(grads["cat_layers.0.bias"]),
                         grad_origin0_b=np_(grads["origin_layers.0.bias"]),
                         grad_inter9_b=np_(grads["intersection_layers.9.bias"]))
        print("g5 losses", img_loss.item(), dn_loss.item(), "grad norm", gnorm)

    for name, d in out.items():
        d["weights_checksum"] = np.float64(csum)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **d)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
