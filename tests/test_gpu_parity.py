"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.  Needs a B200."""

import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
RGB_TOL = 1e-3      # north_star: rgb / depth max-abs error
RAW_TOL = {1: 1e-4, 3: 5e-4}  # pre-activation (r,g,b,sigma) vs the reference: split precision / fp16 single pass
SIGMA_GUARD = 1e-5  # rays whose last sigma is this close to 0 sit on the alpha_last step function


def cu(x):
    return torch.as_tensor(np.asarray(x)).to(DEV)


def scene_rays(H, W, theta=30.0, phi=-30.0, radius=4.0):
    c2w = O.pose_spherical(theta, phi, radius)[:3, :4]
    packed, ro, rd, _ = O.prepare_rays(H, W, O.intrinsics(H, W), c2w=c2w)
    return c2w, packed


# --------------------------------------------------------------------------------------------- tensor-core plumbing
@pytest.mark.parametrize("K,N", [(16, 16), (64, 256), (128, 128), (32, 48)])
def test_umma_selftest(lib, K, N):
    """Operand layout, descriptors, instruction descriptor and TMEM read-back of the MLP kernels."""
    from nerf_sampling_b200 import ops

    g = torch.Generator().manual_seed(K * 1000 + N)
    a = torch.randn(128, K, generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randn(N, K, generator=g).to(torch.bfloat16).to(DEV)
    d = ops.umma_selftest(a, b)
    torch.cuda.synchronize()
    want = a.double() @ b.double().T
    assert float((d.double() - want).abs().max()) < 1e-3 * max(1.0, float(want.abs().max()))


def _debug_sgemm(lib, M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, beta, bias, act, slope, force_fp32):
    from nerf_sampling_b200 import _lib

    _lib.check(lib.b200nerf_debug_sgemm(M, N, K, A.data_ptr(), sAm, sAk, B.data_ptr(), sBk, sBn, C.data_ptr(), ldc, beta,
                                        None if bias is None else bias.data_ptr(), act, slope, force_fp32,
                                        torch.cuda.current_stream().cuda_stream))


@pytest.mark.parametrize("n", [4096, 200, 37])
def test_training_gemm_3xtf32_all_layouts(lib, n):
    """csrc/tgemm.cuh (tcgen05 kind::tf32, hi/lo split) on the three operand layouts of the training passes (Linear forward
    with a column offset into W and beta / bias / LeakyReLU, input gradient, split-K weight gradient) against fp64 torch;
    the error must be fp32-grade: within 6x of the CUDA-core fp32 kernel's own rounding error (or 4e-6 of the output scale)."""
    g = torch.Generator().manual_seed(n)
    rnd = lambda *s: torch.randn(*s, generator=g).to(DEV)  # noqa: E731
    for (M_out, K_in, off, ldw) in ((256, 252, 63, 400), (256, 63, 0, 126), (128, 256, 0, 256), (70, 126, 126, 252 + 7)):
        X, W, bias, Y0 = rnd(n, 300)[:, :K_in], rnd(M_out, ldw), rnd(M_out), rnd(n, M_out)
        # forward: Y = Y0 + X W[:, off:off+K]^T + b, LeakyReLU
        want = torch.nn.functional.leaky_relu((Y0.double() + X.double() @ W[:, off:off + K_in].double().T + bias.double()), 0.01)
        errs = []
        for force in (0, 1):
            Y = Y0.clone()
            _debug_sgemm(lib, n, M_out, K_in, X, X.stride(0), 1, W[:, off:], 1, ldw, Y, M_out, 1, bias, 1, 0.01, force)
            errs.append(float((Y.double() - want).abs().max()))
        scale = float(want.abs().max())
        assert errs[0] <= max(6 * errs[1], 4e-6 * scale), ("fwd", M_out, K_in, errs)
        # input gradient: dX = dY W[:, off:off+K]
        dY = rnd(n, M_out)
        want = dY.double() @ W[:, off:off + K_in].double()
        errs = []
        for force in (0, 1):
            dX = torch.full((n, K_in), 7.0, device=DEV)
            _debug_sgemm(lib, n, K_in, M_out, dY, M_out, 1, W[:, off:], ldw, 1, dX, K_in, 0, None, 0, 0.0, force)
            errs.append(float((dX.double() - want).abs().max()))
        assert errs[0] <= max(6 * errs[1], 4e-6 * float(want.abs().max())), ("dgrad", M_out, K_in, errs)
        # weight gradient into a column block of dW: dW[:, off:off+K] = dY^T X (long reduction over the rays: split K)
        want = dY.double().T @ X.double()
        errs = []
        for force in (0, 1):
            dW = torch.full((M_out, ldw), 3.0, device=DEV)
            _debug_sgemm(lib, M_out, K_in, n, dY, 1, M_out, X, X.stride(0), 1, dW[:, off:], ldw, 0, None, 0, 0.0, force)
            errs.append(float((dW[:, off:off + K_in].double() - want).abs().max()))
            assert bool((dW[:, :off] == 3.0).all()) and bool((dW[:, off + K_in:] == 3.0).all())  # neighbours untouched
        assert errs[0] <= max(6 * errs[1], 4e-6 * float(want.abs().max())), ("wgrad", M_out, K_in, errs)


# --------------------------------------------------------------------------------------------- small operators
def test_get_rays_matches_oracle(lib):
    from nerf_sampling_b200 import ops

    for H, W in ((16, 16), (37, 53)):
        c2w = O.pose_spherical(30.0, -30.0, 4.0)[:3, :4]
        K = O.intrinsics(H, W)
        ro, rd, vd = ops.get_rays(H, W, K, c2w)
        _, oro, ord_, _ = O.prepare_rays(H, W, K, c2w=c2w)
        packed = O.prepare_rays(H, W, K, c2w=c2w)[0]
        assert torch.equal(ro.cpu(), oro)
        assert float((rd.cpu() - ord_).abs().max()) <= 2e-7
        assert float((vd.cpu() - packed[:, 8:11]).abs().max()) <= 2e-7


def test_placement_bit_exact(lib):
    from nerf_sampling_b200 import ops

    g = load_golden("g3")
    mean = cu(g["mean"])
    for S in (1, 2, 3, 13, 32, 33, 64):
        mode = "depth_only" if S == 1 else "uniform"
        z = ops.place_samples(mean, S, mode, 0.25)
        assert torch.equal(z.cpu(), torch.from_numpy(g[f"z_uniform_{S}"])), S
    z = ops.place_samples(mean, 13, "gaussian", 0.3, noise=cu(g["noise"]))
    assert torch.equal(z.cpu(), torch.from_numpy(g["z_gauss"]))
    pts = ops.points(cu(g["rays_o"]), cu(g["rays_d"]), z)
    assert torch.equal(pts.cpu(), torch.from_numpy(g["pts_gauss"]))
    # a ray that misses the sphere has a NaN depth: NaN must survive the clip like torch.clip
    m2 = mean.clone()
    m2[0] = float("nan")
    z = ops.place_samples(m2, 8, "uniform", 0.1)
    assert bool(torch.isnan(z[0]).all()) and not bool(torch.isnan(z[1:]).any())


@pytest.mark.parametrize("n,S", [(1, 2), (5, 3), (33, 8), (100, 32), (257, 64), (64, 192), (19, 1), (40, 70),
                                 # TMA-staged kernel (S = 32 / 64 / 128, >= one 2048-sample tile): ragged last tile, several tiles per CTA
                                 (4097, 64), (70001, 64), (3001, 32), (40000, 32), (1501, 128)])
def test_composite_matches_oracle(lib, n, S):
    from nerf_sampling_b200 import ops

    g = torch.Generator().manual_seed(n * 7 + S)
    raw = torch.randn(n, S, 4, generator=g) * 2
    raw[..., 3] = torch.randn(n, S, generator=g) * 20  # mix of transparent / opaque samples
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), -1).values
    if S > 2:
        z[:, 1] = z[:, 0]  # zero-length interval, as the duplicated mean produces
    rd = torch.randn(n, 3, generator=g)
    for white in (True, False):
        want = O.raw2outputs(raw, z, rd, 0.0, white)
        rgb, disp, acc, depth, w, alphas = ops.composite(raw.to(DEV), z.to(DEV), rd.to(DEV), white)
        for name, a, b in (("rgb", rgb, want[0]), ("acc", acc, want[2]), ("depth", depth, want[3]), ("weights", w, want[6]),
                           ("alphas", alphas, want[5])):
            assert a.shape == b.shape, name
            assert float((a.cpu() - b).abs().max()) <= 2e-6 if a.numel() else True, name
        rel = ((disp.cpu() - want[1]).abs() / want[1].abs().clamp_min(1e-10)).max()
        assert float(rel) <= 1e-5
    noise = torch.randn(n, S, generator=g)
    want = O.raw2outputs(raw, z, rd, 0.5, True, noise=noise)
    rgb, *_ = ops.composite(raw.to(DEV), z.to(DEV), rd.to(DEV), True, noise=(noise * 0.5).to(DEV))
    if S > 1:
        assert float((rgb.cpu() - want[0]).abs().max()) <= 2e-6


def test_composite_empty_batch(lib):
    from nerf_sampling_b200 import ops

    rgb, disp, acc, depth, w, a = ops.composite(torch.zeros(0, 8, 4, device=DEV), torch.zeros(0, 8, device=DEV),
                                                torch.zeros(0, 3, device=DEV))
    assert rgb.shape == (0, 3) and w.shape == (0, 8)


# --------------------------------------------------------------------------------------------- DepthNet
def test_depthnet_matches_oracle_and_golden(lib, oracle_models, b200_models):
    _, _, dn = oracle_models
    _, _, b_dn = b200_models
    g = load_golden("g1")
    ro, rd = cu(g["rays_o"]).reshape(-1, 3), cu(g["rays_d"]).reshape(-1, 3)
    with torch.no_grad():
        z = b_dn(ro, rd)
    assert z.shape == (ro.shape[0], 1)
    assert float((z.cpu() - torch.from_numpy(g["z_mean"])).abs().max()) <= 2e-5
    # ragged size (not a multiple of the 128-row tile) + a ray that misses the sphere
    _, packed = scene_rays(37, 41)
    ro2, rd2 = packed[:, 0:3].contiguous(), packed[:, 3:6].contiguous()
    rd2[5] = torch.tensor([0.0, 1.0, 0.0])
    with torch.no_grad():
        want = O.depthnet_forward(dn, ro2, rd2)
        got = b_dn(ro2.to(DEV), rd2.to(DEV)).cpu()
    assert torch.isnan(want[5]) and torch.isnan(got[5])
    ok = ~torch.isnan(want)
    assert float((got[ok] - want[ok]).abs().max()) <= 2e-5


# --------------------------------------------------------------------------------------------- NeRF MLP
def test_nerf_mlp_matches_golden_raw(lib, b200_models):
    _, b_fine, _ = b200_models
    g = load_golden("g1")
    ro, rd = cu(g["rays_o"]).reshape(-1, 3), cu(g["rays_d"]).reshape(-1, 3)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    z = cu(g["z"]).reshape(-1, int(g["S"]))
    raw = b_fine.query(vd, rays_o=ro, rays_d=rd, z=z)
    want = torch.from_numpy(g["raw"])
    assert raw.shape == want.shape
    tol = RAW_TOL[b_fine.precision]
    assert float((raw.cpu() - want).abs().max()) <= tol
    # explicit sample positions (the run_network entry point)
    raw2 = b_fine.query(vd, pts=cu(g["pts"]).reshape(-1, int(g["S"]), 3))
    assert float((raw2.cpu() - want).abs().max()) <= tol


@pytest.mark.parametrize("n,S", [(1, 1), (3, 5), (130, 1), (77, 13), (512, 64), (1031, 7), (2, 640)])
def test_nerf_mlp_ragged_shapes(lib, oracle_models, b200_models, n, S):
    coarse, _, _ = oracle_models
    b_coarse, _, _ = b200_models
    g = torch.Generator().manual_seed(n + S)
    pts = (torch.rand(n, S, 3, generator=g) * 8 - 4)
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    with torch.no_grad():
        want = O.run_network(pts, vd, coarse)
    got = b_coarse.query(vd.to(DEV), pts=pts.to(DEV)).cpu()
    assert float((got - want).abs().max()) <= RAW_TOL[b_coarse.precision]


def test_nerf_mlp_nan_stays_in_its_row(lib, oracle_models, b200_models):
    """A sample at NaN (ray that missed the sphere -> NaN depth) must poison only its own output row."""
    coarse, _, _ = oracle_models
    b_coarse, _, _ = b200_models
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(300, 4, 3, generator=g) * 4 - 2
    pts[7, 2, 1] = float("nan")
    vd = torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1)
    with torch.no_grad():
        want = O.run_network(pts, vd, coarse)
    got = b_coarse.query(vd.to(DEV), pts=pts.to(DEV)).cpu()
    bad = torch.isnan(got).any(-1)
    assert bool(bad[7, 2]) and int(bad.sum()) == 1
    assert float((got[~bad] - want[~bad]).abs().max()) <= RAW_TOL[b_coarse.precision]


def test_fast_kernel_guard_band_vs_split(lib, oracle_models):
    """fp16 single pass + guard band against the split-precision kernel on one 200x200x32 view: the guard band
    must contain every last-of-ray sample whose sigma sign differs, and re-evaluated samples are split-exact."""
    from nerf_sampling_b200 import ops
    from nerf_sampling_b200.packing import PREC_FAST, PREC_FP16, PREC_SPLIT, PackedDepthNet, PackedNeRF

    _, fine, dn = oracle_models
    _, packed = scene_rays(200, 200)
    ro, rd, vd = (packed[:, a:b].contiguous().to(DEV) for a, b in ((0, 3), (3, 6), (8, 11)))
    z = ops.place_samples(ops.depthnet_forward(PackedDepthNet(dn, DEV, PREC_SPLIT), ro, rd), 32, "uniform", 0.1)
    raw_s = ops.nerf_mlp(PackedNeRF(fine, DEV, PREC_SPLIT), vd, rays_o=ro, rays_d=rd, z=z)
    raw_h = ops.nerf_mlp(PackedNeRF(fine, DEV, PREC_FP16), vd, rays_o=ro, rays_d=rd, z=z)
    raw_f = ops.nerf_mlp(PackedNeRF(fine, DEV, PREC_FAST), vd, rays_o=ro, rays_d=rd, z=z)
    ws = ops.nerf_mlp.last_guard_ws
    cnt = int(ws[0])
    assert 0 < cnt < 20000                                    # a band, not everything
    assert float((raw_h - raw_s).abs().max()) <= RAW_TOL[3]   # fp16 single pass is close everywhere ...
    sign = lambda r: r[:, -1, 3] > 0                          # noqa: E731
    flips_h = sign(raw_h) != sign(raw_s)
    assert int((sign(raw_f) != sign(raw_s)).sum()) == 0       # ... and the guard band removes the sign flips
    lst = ws[4 : 4 + cnt].long()
    assert bool(((lst % 32) == 31).all())                     # only last-of-ray samples are flagged
    flagged = torch.zeros(ro.shape[0], dtype=torch.bool, device=DEV)
    flagged[lst // 32] = True
    assert bool(flagged[flips_h].all())                       # every fp16 sign flip was inside the band
    assert torch.equal(raw_f.reshape(-1, 4)[lst], raw_s.reshape(-1, 4)[lst])
    rest = torch.ones(raw_f.numel() // 4, dtype=torch.bool, device=DEV)
    rest[lst] = False
    assert torch.equal(raw_f.reshape(-1, 4)[rest], raw_h.reshape(-1, 4)[rest])


# --------------------------------------------------------------------------------------------- whole path
def render_b200(b200_models, H, W, S, chunk=1024 * 32, **flags):
    from nerf_sampling_b200.nerf_pytorch import nerf_utils
    from nerf_sampling_b200.trainers import DepthNetTrainer

    b_coarse, b_fine, b_dn = b200_models
    tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=True,
                         white_bkgd=True, device=DEV, n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                         input_dims_embed=3, distance=0.1, sampling_mode="uniform", n_depth_samples=S, **flags)
    kw = dict(network_fn=b_coarse, network_fine=b_fine, depth_network=b_dn, network_query_fn=None, N_samples=64,
              N_importance=128, trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=False, lindisp=True, ndc=False,
              near=2.0, far=6.0, use_viewdirs=True, model_mode="test")
    c2w = O.pose_spherical(30.0, -30.0, 4.0)
    with torch.no_grad():
        return nerf_utils.render_test(H, W, O.intrinsics(H, W), chunk=chunk, c2w=c2w[:3, :4], **kw)


def test_render_tiny_view_matches_golden(lib, b200_models):
    g = load_golden("g1")
    rgb, disp, ex = render_b200(b200_models, int(g["H"]), int(g["W"]), int(g["S"]))
    assert rgb.shape == (16, 16, 3) and disp.shape == (16, 16)
    assert torch.equal(ex["depth_net_z_vals"].cpu() != ex["depth_net_z_vals"].cpu(), torch.zeros(16, 16, 8, dtype=torch.bool))
    assert float((ex["depth_net_z_vals"].cpu() - torch.from_numpy(g["z"])).abs().max()) <= 2e-5
    assert float((ex["depth_net_weights"].cpu() - torch.from_numpy(g["weights"])).abs().max()) <= RGB_TOL
    assert float((rgb.cpu() - torch.from_numpy(g["rgb"])).abs().max()) <= RGB_TOL
    rel = ((disp.cpu() - torch.from_numpy(g["disp"])).abs() / torch.from_numpy(g["disp"]).abs()).max()
    assert float(rel) <= 1e-3


def test_render_config1_matches_golden(lib, b200_models):
    """BASELINE config #1 (200x200, 32 samples) against the reference's own output."""
    g = load_golden("g2")
    rgb, disp, ex = render_b200(b200_models, 200, 200, 32)
    want = torch.from_numpy(g["rgb"])
    err = (rgb.cpu() - want).abs().max(-1).values
    knife = torch.from_numpy(np.abs(g["sigma_last"]) < SIGMA_GUARD).reshape(200, 200)
    assert float(knife.float().mean()) < 0.002
    assert float(err[~knife].max()) <= RGB_TOL, f"max err {float(err[~knife].max())}"
    target = torch.rand(200, 200, 3, generator=torch.Generator().manual_seed(1))
    assert abs(O.psnr(rgb.cpu(), target) - O.psnr(want, target)) <= 0.05
    # chunk size must not change the result (the reference is chunk-invariant too); honour `chunk` literally for this check
    from nerf_sampling_b200.nerf_pytorch import nerf_utils

    saved, nerf_utils.COALESCE_RAYS = nerf_utils.COALESCE_RAYS, 0
    try:
        rgb2, _, _ = render_b200(b200_models, 200, 200, 32, chunk=4096)
    finally:
        nerf_utils.COALESCE_RAYS = saved
    assert torch.equal(rgb, rgb2)


def test_render_full_size_vs_oracle_on_device(lib, oracle_models, b200_models):
    """BASELINE config #2 (800x800, 64 samples): oracle arithmetic evaluated by torch on the GPU (fp32, TF32 off)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    coarse, fine, dn = (O.params_to(p, DEV) for p in oracle_models)
    H = W = 800
    rgb, disp, ex = render_b200(b200_models, H, W, 64, chunk=H * W)
    c2w = O.pose_spherical(30.0, -30.0, 4.0).to(DEV)
    with torch.no_grad():
        o = O.render_view(H, W, O.intrinsics(H, W), c2w[:3, :4], coarse, fine, dn, chunk=32768, n_depth_samples=64,
                          sampling_mode="uniform", distance=0.1)
    assert float((ex["depth_net_z_vals"] - o["depth_net_z_vals"]).abs().max()) <= 5e-5
    err = (rgb - o["depth_net_rgb_map"]).abs().max(-1).values
    knife = o["raw"][..., -1, 3].abs() < SIGMA_GUARD
    frac = float(knife.float().mean())
    print(f"knife-edge rays (|sigma_last| < {SIGMA_GUARD}): {int(knife.sum())} of {H * W} ({100 * frac:.3f}%); "
          f"max rgb err elsewhere {float(err[~knife].max()):.2e}; flipped {int((err > RGB_TOL).sum())}")
    assert frac < 0.002
    assert float(err[~knife].max()) <= RGB_TOL
    target = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1)).to(DEV)
    assert abs(O.psnr(rgb, target) - O.psnr(o["depth_net_rgb_map"], target)) <= 0.05
    # size-independent properties: weights are a sub-probability distribution, depths sorted and clipped
    w, z = ex["depth_net_weights"], ex["depth_net_z_vals"]
    assert bool((w >= 0).all()) and float(w.sum(-1).max()) <= 1 + 1e-5
    assert bool((z[..., 1:] >= z[..., :-1]).all()) and float(z.min()) >= 2.0 and float(z.max()) <= 6.0


# --------------------------------------------------------------------------------------------- vanilla hierarchical path
def test_coarse_depths_bit_exact(lib):
    from nerf_sampling_b200 import ops

    g = load_golden("g4")
    n = g["z_coarse"].shape[0]
    near, far = torch.full((n, 1), 2.0, device=DEV), torch.full((n, 1), 6.0, device=DEV)
    z = ops.coarse_z(near, far, n, 64, True)
    assert torch.equal(z.cpu(), torch.from_numpy(g["z_coarse"]))
    for lindisp in (True, False):
        tr = torch.rand(n, 64, generator=torch.Generator().manual_seed(5))
        want = O.coarse_z(near.cpu(), far.cpu(), n, 64, lindisp, t_rand=tr)
        got = ops.coarse_z(near, far, n, 64, lindisp, tr.to(DEV))
        assert torch.equal(got.cpu(), want), lindisp


def test_sample_pdf_indices_bit_exact_on_reference_inputs(lib):
    """Integer-derived outputs: searchsorted indices on the reference's own (bins, weights)."""
    from nerf_sampling_b200 import ops

    g = load_golden("g4")
    zc, wc = cu(g["z_coarse"]), cu(g["weights_coarse"])
    zs, z_all, inds = ops.sample_pdf_merge(zc, wc, 128, return_inds=True)
    want_inds = torch.from_numpy(g["inds"])
    mism = int((inds.cpu() != want_inds).sum())
    print(f"sample_pdf index mismatches: {mism} of {want_inds.numel()}")
    assert mism == 0
    assert float((zs.cpu() - torch.from_numpy(g["z_samples"])).abs().max()) <= 2e-6
    assert float((z_all.cpu() - torch.from_numpy(g["z_fine"])).abs().max()) <= 2e-6
    assert bool((z_all[:, 1:] >= z_all[:, :-1]).all())
    # the merge of the two sorted runs is exactly the sort of their concatenation (same multiset, bit for bit)
    assert torch.equal(z_all, torch.sort(torch.cat([zc, zs], -1), -1).values)
    # ties: every sample identical (u constant), and samples that coincide with coarse depths
    u_const = torch.full((zc.shape[0], 24), 0.5)
    zs2, za2, _ = ops.sample_pdf_merge(zc, wc, 24, u=u_const.to(DEV))
    assert torch.equal(za2, torch.sort(torch.cat([zc, zs2], -1), -1).values)
    # generic entry point (bins, weights) + random u (unsorted samples)
    mid = 0.5 * (zc[:, 1:] + zc[:, :-1])
    u = torch.rand(zc.shape[0], 40, generator=torch.Generator().manual_seed(9))
    want, wi = O.sample_pdf(mid.cpu(), wc[:, 1:-1].cpu(), 40, u=u, return_inds=True)
    got, gi = ops.sample_pdf(mid, wc[:, 1:-1].contiguous(), 40, u=u.to(DEV), return_inds=True)
    assert int((gi.cpu() != wi).sum()) == 0
    assert float((got.cpu() - want).abs().max()) <= 2e-6
    _, za, _ = ops.sample_pdf_merge(zc, wc, 40, u=u.to(DEV))
    want_all = torch.sort(torch.cat([zc.cpu(), want], -1), -1).values
    assert float((za.cpu() - want_all).abs().max()) <= 2e-6


def test_argmax_gather(lib):
    from nerf_sampling_b200 import ops

    g = torch.Generator().manual_seed(11)
    w = torch.rand(300, 192, generator=g)
    w[5, 10] = w[5, 100] = 2.0      # tie -> first index
    w[6] = 0.0                      # all equal -> index 0
    z = torch.rand(300, 192, generator=g)
    raw = torch.randn(300, 192, 4, generator=g)
    idx, mz, mw, rgb = ops.argmax_gather(w.to(DEV), z.to(DEV), raw.to(DEV))
    top = w.argmax(dim=1, keepdim=True)
    assert torch.equal(idx.cpu(), top)
    assert torch.equal(mz.cpu(), torch.gather(z, 1, top)) and torch.equal(mw.cpu(), torch.gather(w, 1, top))
    want_rgb = torch.gather(torch.sigmoid(raw[..., :3]), 1, top.unsqueeze(-1).expand(-1, 1, 3)).squeeze(1)
    assert float((rgb.cpu() - want_rgb).abs().max()) <= 1e-6


def test_full_nerf_and_max_pts_modes_match_golden(lib, b200_models):
    """render.py -nf / -nm: vanilla 64 + 128 hierarchical render (config #4 shape) against the reference."""
    g = load_golden("g4")
    H = W = int(g["H"])
    rgb, disp, ex = render_b200(b200_models, H, W, 32, use_full_nerf=True)
    assert float((ex["depth_net_z_vals"].reshape(-1, 192).cpu() - torch.from_numpy(g["z_fine"])).abs().max()) <= 1e-4
    assert float((ex["depth_net_weights"].reshape(-1, 192).cpu() - torch.from_numpy(g["weights_fine"])).abs().max()) <= RGB_TOL
    assert float((rgb.reshape(-1, 3).cpu() - torch.from_numpy(g["rgb"])).abs().max()) <= RGB_TOL
    rgb_m, disp_m, ex_m = render_b200(b200_models, H, W, 32, use_nerf_max_pts=True)
    assert float((ex_m["max_z_vals"].reshape(-1, 1).cpu() - torch.from_numpy(g["max_z"])).abs().max()) <= 1e-4
    assert float((rgb_m.reshape(-1, 3).cpu() - torch.from_numpy(g["max_rgb"])).abs().max()) <= RGB_TOL
    assert float(disp_m.abs().max()) == 0.0


def test_training_render_forward_matches_golden(lib, b200_models):
    """render_rays (train mode, perturb=0): arg-max target depth, DepthNet depth and the 1-sample colour."""
    from nerf_sampling_b200.nerf_pytorch import nerf_utils
    from nerf_sampling_b200.trainers import DepthNetTrainer

    g = load_golden("g5")
    b_coarse, b_fine, b_dn = b200_models
    tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=True,
                         white_bkgd=True, device=DEV, n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                         input_dims_embed=3, perturb=0.0)
    kw = dict(network_fn=b_coarse, network_fine=b_fine, depth_network=b_dn, network_query_fn=None, N_samples=64,
              N_importance=128, trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=0.0, lindisp=True, ndc=False, near=2.0,
              far=6.0, use_viewdirs=True, model_mode="train")
    with torch.no_grad():
        rgb, disp, ex = nerf_utils.render(800, 800, O.intrinsics(800, 800), rays=(cu(g["rays_o"]), cu(g["rays_d"])), retraw=True, **kw)
    assert float((ex["depth_net_z_vals"].cpu() - torch.from_numpy(g["z_dn"])).abs().max()) <= 2e-5
    assert float((ex["max_z_vals"].cpu() - torch.from_numpy(g["max_z"])).abs().max()) <= 1e-4
    assert float((rgb.cpu() - torch.from_numpy(g["rgb"])).abs().max()) <= RGB_TOL
    assert float(disp.min()) == 1e10  # S == 1 quirk


# --------------------------------------------------------------------------------------------- training step (config #5)
def _train_setup(b200_models):
    from nerf_sampling_b200.trainers import DepthNetTrainer

    b_coarse, b_fine, b_dn = b200_models
    tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=True,
                         white_bkgd=True, device=DEV, n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                         input_dims_embed=3, perturb=0.0)
    kw = dict(network_fn=b_coarse, network_fine=b_fine, depth_network=b_dn, network_query_fn=None, N_samples=64,
              N_importance=128, trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=0.0, lindisp=True, ndc=False, near=2.0,
              far=6.0, use_viewdirs=True, model_mode="train")
    return tr, kw


def test_training_step_gradients_match_golden(lib, b200_models):
    """core_optimization_loop (Trainer.py:506-544): both losses and the DepthNet gradients of the reference's own
    backward pass on the same 96 rays / targets (fixture g5)."""
    from nerf_sampling_b200.nerf_pytorch import nerf_utils

    g = load_golden("g5")
    _, _, b_dn = b200_models
    tr, kw = _train_setup(b200_models)
    for p in b_dn.parameters():
        p.grad = None
    rgb, disp, ex = nerf_utils.render(800, 800, O.intrinsics(800, 800), rays=(cu(g["rays_o"]), cu(g["rays_d"])), retraw=True, **kw)
    assert rgb.requires_grad and ex["depth_net_z_vals"].requires_grad and not ex["max_z_vals"].requires_grad
    target = cu(g["target"])
    img_loss = torch.mean((rgb - target) ** 2)
    dn_loss = torch.nn.functional.mse_loss(ex["depth_net_z_vals"], ex["max_z_vals"])
    dn_loss.backward(retain_graph=True)
    img_loss.backward()
    assert abs(float(img_loss) - float(g["img_loss"])) <= 1e-5 * max(1.0, float(g["img_loss"]))
    assert abs(float(dn_loss) - float(g["dn_loss"])) <= 1e-4 * max(1.0, float(g["dn_loss"]))
    grads = {k: p.grad for k, p in b_dn.named_parameters()}
    assert all(v is not None and bool(torch.isfinite(v).all()) for v in grads.values())
    gnorm = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    assert abs(gnorm - float(g["grad_norm"])) <= 2e-3 * float(g["grad_norm"]), (gnorm, float(g["grad_norm"]))
    for key, name in (("grad_to_depth_w", "to_depth.0.weight"), ("grad_cat0_b", "cat_layers.0.bias"),
                      ("grad_origin0_b", "origin_layers.0.bias"), ("grad_inter9_b", "intersection_layers.9.bias")):
        want = torch.from_numpy(g[key])
        got = grads[name].cpu()
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 2e-3 * float(want.abs().max()) + 1e-9, name
    for p in b_dn.parameters():
        p.grad = None


@pytest.mark.parametrize("chain,H,W", [("gemm", 9, 13), ("fused", 9, 13), ("fused", 40, 50)])
def test_depthnet_literal_forward_and_backward_vs_torch(lib, oracle_models, chain, H, W, monkeypatch):
    """The training form of DepthNet against the oracle's literal forward differentiated by torch autograd.  "gemm": every layer a
    3xTF32 product (2^-20 per link); "fused": the activated cat layers as ONE launch of the split-precision MLP kernel (bf16 hi + lo
    operands, 2^-17 per link; its saved activations agree with the 3xTF32 ones to 5e-6, tools/chain_diag.py).

    The derivative of a LeakyReLU network jumps where a pre-activation changes sign.  A hidden value within rounding distance of
    zero may take the other slope than torch's; that changes ONE ray's contribution to one row of a weight gradient by ~100 %, i.e.
    the row by ~1 / n_rays.  With 2^-17 links this happens for about 3 of 1000 rays x 2560 hidden values (0 or 1 of the 117 rays of
    the small case: observed 1e-2 of the tensor's largest entry when it does), so the fused route is held to the 2e-3 bound on 2000
    rays and to 2 / n_rays on 117; at the 4096 rays of BASELINE config #5 a flip moves a gradient by 2.4e-4
    (test_training_step_4096_rays_all_gradients_vs_autograd: worst tensor 4.4e-4 with either route)."""
    from nerf_sampling_b200.depth_nets import DepthNet

    monkeypatch.setenv("B200NERF_TRAIN_CHAIN", chain)
    _, _, dn = oracle_models
    m = DepthNet(hidden_sizes=[256] * 10, cat_hidden_sizes=[256] * 10, sphere_radius=2.0)
    m.load_state_dict(dn)
    m.to(DEV)
    _, packed = scene_rays(H, W)
    ro, rd = packed[:, 0:3].contiguous().to(DEV), packed[:, 3:6].contiguous().to(DEV)
    z = m(ro, rd)
    w = torch.linspace(0.5, 1.5, z.numel(), device=DEV).reshape(z.shape)  # one sign: no cancellation in the bias sums
    (z * w).sum().backward()
    ref = {k: v.clone().to(DEV).requires_grad_(True) for k, v in dn.items()}
    zr = O.depthnet_forward(ref, ro, rd)
    (zr * w).sum().backward()
    assert float((z - zr).abs().max()) <= 1e-5
    tol = 2e-3 if (chain == "gemm" or H * W >= 2000) else 2.0 / (H * W)
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        want = ref[k].grad
        worst = max(worst, (k, float((p.grad - want).abs().max()) / (float(want.abs().max()) + 1e-12)), key=lambda t: t[1])
        assert float((p.grad - want).abs().max()) <= tol * float(want.abs().max()) + 1e-7, k  # sums with cancellation
    print(f"DepthNet training form, chain={chain}, {H * W} rays: worst tensor {worst[0]} {worst[1]:.2e} of its largest entry")


@pytest.mark.parametrize("route,side", [("fp32", 8), ("split", 8), ("split", 37)])
def test_nerf_point_jvp_vs_torch(lib, oracle_models, b200_models, route, side):
    """raw and d raw / d z of the frozen NeRF at one sample per ray against torch autograd on the oracle: the grouped 3xTF32
    products over the fp32 tensors ("fp32") and the two-launch route over the packed split-precision model ("split": primal +
    ReLU masks, tangent pass; 37 x 37 rays = ragged tiles on several CTA pairs)."""
    from nerf_sampling_b200 import training

    _, fine, _ = oracle_models
    _, b_fine, _ = b200_models
    _, packed = scene_rays(side, side)
    ro, rd, vd = (packed[:, a:b].contiguous().to(DEV) for a, b in ((0, 3), (3, 6), (8, 11)))
    z = (3.0 + torch.rand(ro.shape[0], 1, generator=torch.Generator().manual_seed(2))).to(DEV).requires_grad_(True)
    raw = training.NerfPointFn.apply(z, ro, rd, vd, b_fine.packed() if route == "split" else None, *training.nerf_params(b_fine))
    fp = O.params_to(fine, DEV)
    z2 = z.detach().clone().requires_grad_(True)
    pts = ro[:, None, :] + rd[:, None, :] * z2[:, :, None]
    want = O.run_network(pts, vd, fp)
    assert float((raw - want).abs().max()) <= (1e-5 if route == "fp32" else 1e-4)   # 1e-4: the split-precision bound of inference
    # The derivative of a ReLU network jumps where a pre-activation crosses zero: a ray with one of its 2176 hidden units within
    # rounding distance of the kink may take the other branch than the fp32 oracle (the raw outputs still agree).  Such rays are
    # counted and bounded, every other ray is held to 1e-3 of the channel's largest derivative.
    worst, kink = 0.0, torch.zeros(ro.shape[0], dtype=torch.bool, device=DEV)
    errs = []
    for c in range(4):
        gw, = torch.autograd.grad(want[..., c].sum(), z2, retain_graph=True)
        gg, = torch.autograd.grad(raw[..., c].sum(), z, retain_graph=True)
        err = ((gg - gw).abs() / float(gw.abs().max())).reshape(-1)
        errs.append(err)
        kink |= err > 1e-3
    n_kink = int(kink.sum())
    for err in errs:
        if n_kink < ro.shape[0]:
            worst = max(worst, float(err[~kink].max()))
    print(f"d raw / d z, route {route}, {side * side} rays: max-abs raw error {float((raw - want).abs().max()):.2e}, "
          f"worst derivative error {worst:.2e} of the channel's largest, rays across a ReLU kink: {n_kink} "
          f"(largest error there {max(float(e.max()) for e in errs):.2e})")
    assert n_kink <= max(1, ro.shape[0] // 200)


def test_adam_matches_torch(lib):
    from nerf_sampling_b200 import training

    torch.manual_seed(0)
    p1 = torch.nn.Parameter(torch.randn(300, 70, device=DEV))
    p2 = torch.nn.Parameter(p1.detach().clone())
    ours, ref = training.Adam([p1], lr=1e-4), torch.optim.Adam([p2], lr=1e-4)
    for step in range(5):
        gr = torch.randn_like(p1) * (10.0 ** (step - 2))
        p1.grad, p2.grad = gr.clone(), gr.clone()
        ours.step()
        ref.step()
    assert float((p1 - p2).abs().max()) <= 1e-6  # one ulp at |p| ~ 2: torch orders the update differently


def test_core_optimization_loop_step(lib, oracle_models):
    """One Trainer.core_optimization_loop step: losses of the reference, DepthNet moves, NeRFs stay frozen."""
    from nerf_sampling_b200 import training
    from nerf_sampling_b200.depth_nets import DepthNet
    from nerf_sampling_b200.nerf_pytorch.run_nerf_helpers import NeRF

    coarse, fine, dn = oracle_models

    def nerf(sd):
        m = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        m.load_state_dict(sd)
        return m.to(DEV)

    d = DepthNet(hidden_sizes=[256] * 10, cat_hidden_sizes=[256] * 10, sphere_radius=2.0)
    d.load_state_dict(dn)
    models = (nerf(coarse), nerf(fine), d.to(DEV))
    tr, kw = _train_setup(models)
    tr.H, tr.W, tr.K, tr.chunk = 800, 800, O.intrinsics(800, 800), 32768
    g = load_golden("g5")
    opt = training.Adam(list(models[2].parameters()), lr=1e-4)
    before = [p.detach().clone() for p in models[2].parameters()]
    fine_before = [p.detach().clone() for p in models[1].parameters()]
    loss, dn_loss, psnr, psnr0 = tr.core_optimization_loop(opt, kw, (cu(g["rays_o"]), cu(g["rays_d"])), 0, cu(g["target"]))
    assert abs(float(loss) - float(g["img_loss"])) <= 1e-5 and abs(float(dn_loss) - float(g["dn_loss"])) <= 1e-4 and psnr0 is None
    moved = [float((a - b).abs().max()) for a, b in zip(before, models[2].parameters())]
    assert all(0.0 < m <= 1.01e-4 for m in moved)   # first Adam step moves every weight by ~lr
    assert all(torch.equal(a, b) for a, b in zip(fine_before, models[1].parameters()))
    loss2, dn_loss2, _, _ = tr.core_optimization_loop(opt, kw, (cu(g["rays_o"]), cu(g["rays_d"])), 1, cu(g["target"]))
    assert float(dn_loss2) < float(dn_loss)  # the depth loss goes down on the batch it was fitted to


def test_render_path_matches_single_views(lib, b200_models, tmp_path):
    """render_path (nerf_utils.py:258-360): pipelined multi-pose render == the same poses rendered one by one;
    PSNR bookkeeping and PNG output."""
    import numpy as np

    from nerf_sampling_b200.nerf_pytorch import nerf_utils
    from nerf_sampling_b200.trainers import DepthNetTrainer

    b_coarse, b_fine, b_dn = b200_models
    tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=True,
                         white_bkgd=True, device=DEV, n_layers=10, layer_width=256, N_importance=128, N_samples=64,
                         input_dims_embed=3, distance=0.1, sampling_mode="uniform", n_depth_samples=16)
    kw = dict(network_fn=b_coarse, network_fine=b_fine, depth_network=b_dn, network_query_fn=None, N_samples=64,
              N_importance=128, trainer=tr, white_bkgd=True, raw_noise_std=0.0, perturb=False, lindisp=True, ndc=False,
              near=2.0, far=6.0, use_viewdirs=True, model_mode="test")
    H = W = 40
    K = O.intrinsics(H, W)
    poses = torch.stack([O.pose_spherical(a, -30.0, 4.0) for a in (0.0, 40.0, 80.0)])
    gt = np.random.default_rng(0).random((3, H, W, 3), dtype=np.float32)
    with torch.no_grad():
        rgbs, disps, psnr = nerf_utils.render_path(poses, [H, W, float(K[0][0])], K, 32768, kw, gt_imgs=gt, savedir=str(tmp_path))
        singles = [nerf_utils.render_test(H, W, K, chunk=32768, c2w=p[:3, :4], **kw) for p in poses]
    assert rgbs.shape == (3, H, W, 3) and disps.shape == (3, H, W)
    for i, (rgb, disp, _) in enumerate(singles):
        assert np.array_equal(rgbs[i], rgb.cpu().numpy()) and np.array_equal(disps[i], disp.cpu().numpy())
    want = np.mean([-10.0 * np.log10(np.mean(np.square(rgbs[i] - gt[i]))) for i in range(3)])
    assert abs(psnr - want) < 1e-9
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".png")) == ["000.png", "001.png", "002.png"]
    assert "Avg of 3 images" in open(tmp_path / "psnr.txt").read()


def test_sample_random_ray_batch_matches_full_ray_grid(lib):
    """Trainer.sample_random_ray_batch (Trainer.py:400-475): rays generated for the selected pixels only are bit-identical
    to indexing the full get_rays grid; the reference RNG option reproduces np.random.choice; centre crop respected."""
    from nerf_sampling_b200 import ops
    from nerf_sampling_b200.trainers import DepthNetTrainer

    H, W = 60, 80
    tr = DepthNetTrainer(dataset_type="blender", basedir="/tmp", expname="x", no_batching=True, datadir="x", half_res=True,
                         white_bkgd=True, device=DEV, N_rand=128, precrop_iters=5, precrop_frac=0.5)
    tr.H, tr.W, tr.K = H, W, O.intrinsics(H, W)
    images = torch.rand(3, H, W, 3, generator=torch.Generator().manual_seed(0))
    poses = torch.stack([O.pose_spherical(a, -30.0, 4.0) for a in (0.0, 40.0, 80.0)])
    tr.reference_rng = True
    np.random.seed(7)
    _, _, batch_rays, target = tr.sample_random_ray_batch(None, 0, [0, 1, 2], images, poses, i=10)
    np.random.seed(7)
    img_i = int(np.random.choice([0, 1, 2]))
    sel = torch.from_numpy(np.random.choice(H * W, size=[128], replace=False))
    ro, rd, _ = ops.get_rays(H, W, tr.K, poses[img_i][:3, :4])
    assert torch.equal(batch_rays[0].cpu(), ro.cpu()[sel]) and torch.equal(batch_rays[1].cpu(), rd.cpu()[sel])
    assert torch.equal(target.cpu(), images[img_i].reshape(-1, 3)[sel])
    tr.reference_rng = False
    _, _, batch_rays, target = tr.sample_random_ray_batch(None, 0, [1], images, poses, i=0)   # inside the pre-crop phase
    ro, rd, _ = ops.get_rays(H, W, tr.K, poses[1][:3, :4])
    crop = torch.zeros(H, W, dtype=torch.bool)
    crop[H // 2 - 15 : H // 2 + 15, W // 2 - 20 : W // 2 + 20] = True
    allowed = {tuple(r) for r in rd.cpu()[crop.reshape(-1)].tolist()}
    assert batch_rays.shape == (2, 128, 3) and all(tuple(r) in allowed for r in batch_rays[1].cpu().tolist())
    assert len({tuple(r) for r in batch_rays[1].cpu().tolist()}) == 128   # without replacement


def test_graphed_training_step_matches_eager(lib, oracle_models):
    """core_optimization_loop captured in one CUDA graph: three replays move DepthNet exactly like three eager steps."""
    from nerf_sampling_b200 import training
    from nerf_sampling_b200.depth_nets import DepthNet
    from nerf_sampling_b200.nerf_pytorch.run_nerf_helpers import NeRF

    coarse, fine, dn = oracle_models
    g = load_golden("g5")
    rays, target = (cu(g["rays_o"]), cu(g["rays_d"])), cu(g["target"])

    def build():
        def nerf(sd):
            m = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
            m.load_state_dict(sd)
            return m.to(DEV)

        d = DepthNet(hidden_sizes=[256] * 10, cat_hidden_sizes=[256] * 10, sphere_radius=2.0)
        d.load_state_dict(dn)
        models = (nerf(coarse), nerf(fine), d.to(DEV))
        tr, kw = _train_setup(models)
        tr.H, tr.W, tr.K, tr.chunk = 800, 800, O.intrinsics(800, 800), 32768
        return models, tr, kw, training.Adam(list(models[2].parameters()), lr=1e-4)

    models_e, tr_e, kw_e, opt_e = build()
    eager = [tr_e.core_optimization_loop(opt_e, kw_e, rays, i, target) for i in range(3)]
    models_g, tr_g, kw_g, opt_g = build()
    step = training.GraphedTrainStep(tr_g, opt_g, kw_g, rays[0].shape[0])
    graphed = []
    for i in range(3):
        out = step(rays, target)
        graphed.append(tuple(float(x) for x in out))
    torch.cuda.synchronize()
    for e, gq in zip(eager, graphed):
        assert abs(float(e[0]) - gq[0]) <= 1e-6 and abs(float(e[1]) - gq[1]) <= 1e-5
    for a, b in zip(models_e[2].parameters(), models_g[2].parameters()):
        assert float((a - b).abs().max()) <= 2e-6   # split-K atomics make the weight gradients order-dependent in the last bits
    assert int(opt_g.state[next(models_g[2].parameters())]["step"]) == 3
