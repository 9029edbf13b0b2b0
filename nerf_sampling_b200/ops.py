"""Operator-level host API: torch CUDA tensors in, C-ABI kernels on the current stream, torch tensors out.

Every function here is a thin marshal around one ``b200nerf_*`` entry point; there is no torch arithmetic on the
data path and no CPU fallback (CPU tensors are rejected).
"""

from __future__ import annotations

import ctypes as C
import functools
from typing import Optional, Tuple

import torch

from . import _lib
from .packing import PREC_FAST, PackedDepthNet, PackedNeRF

PLACE_MODES = {"depth_only": 0, "uniform": 1, "gaussian": 2}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B200NerfError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def get_rays(H: int, W: int, K, c2w, device=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """rays_o, rays_d, viewdirs ([H*W,3] each) for a pinhole view (run_nerf_helpers.py:187-202)."""
    device = torch.device(device if device is not None else "cuda")
    c = torch.as_tensor(c2w, dtype=torch.float32, device="cpu")[:3, :4].contiguous()
    n = H * W
    ro = torch.empty(n, 3, device=device)
    rd = torch.empty(n, 3, device=device)
    vd = torch.empty(n, 3, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().b200nerf_get_rays(H, W, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]),
                                                c.data_ptr(), _p(ro), _p(rd), _p(vd), _stream()))
    return ro, rd, vd


def get_rays_at(H: int, W: int, K, c2w, pix: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """rays_o, rays_d, viewdirs [n,3] of the pixels ``pix`` (flat row-major int64 indices on the device)."""
    if not pix.is_cuda or pix.dtype != torch.int64:
        raise _lib.B200NerfError("pix must be an int64 CUDA tensor")
    pix = pix.contiguous()
    c = torch.as_tensor(c2w, dtype=torch.float32, device="cpu")[:3, :4].contiguous()
    n = pix.numel()
    ro, rd, vd = (torch.empty(n, 3, device=pix.device) for _ in range(3))
    with torch.cuda.device(pix.device):
        _lib.check(_lib.lib().b200nerf_get_rays_at(H, W, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]),
                                                   c.data_ptr(), _p(pix), n, _p(ro), _p(rd), _p(vd), _stream()))
    return ro, rd, vd


def gather_pixels(image: torch.Tensor, pix: torch.Tensor) -> torch.Tensor:
    """image.reshape(-1, C)[pix] for a device image [H,W,C] and int64 pixel indices."""
    image = _dev(image, "image")
    ch = image.shape[-1]
    pix = pix.contiguous()
    out = torch.empty(pix.numel(), ch, device=image.device)
    with torch.cuda.device(image.device):
        _lib.check(_lib.lib().b200nerf_gather_pixels(_p(image), _p(pix), pix.numel(), ch, _p(out), _stream()))
    return out


def normalize_dirs(rays_d: torch.Tensor) -> torch.Tensor:
    rays_d = _dev(rays_d, "rays_d").reshape(-1, 3)
    out = torch.empty_like(rays_d)
    with torch.cuda.device(rays_d.device):
        _lib.check(_lib.lib().b200nerf_normalize_dirs(_p(rays_d), rays_d.shape[0], _p(out), _stream()))
    return out


def depthnet_forward(pk: PackedDepthNet, rays_o, rays_d, radius=2.0, near=2.0, far=6.0) -> torch.Tensor:
    """DepthNet.forward -> [N,1] (depth_nets/depth_net.py:117-169)."""
    rays_o, rays_d = _dev(rays_o, "rays_o"), _dev(rays_d, "rays_d")
    n = rays_o.shape[0]
    out = torch.empty(n, 1, device=rays_o.device)
    with torch.cuda.device(rays_o.device):
        _lib.check(_lib.lib().b200nerf_depthnet_fwd(_p(pk.wpack), _p(pk.aux), pk.n_hidden, pk.prec, _p(rays_o), _p(rays_d), n,
                                                    float(radius), float(near), float(far), _p(out), _stream()))
    return out


@functools.lru_cache(maxsize=64)
def _device_linspace(lo: float, hi: float, steps: int, device: str) -> torch.Tensor:
    """torch.linspace evaluated on the CPU exactly as the reference does (so grids are bit-identical), uploaded once per
    (range, length, device): a pageable host->device copy per call would synchronise the stream every step."""
    return torch.linspace(lo, hi, steps=steps, device="cpu").to(device)


def uniform_grid(std: float, n_samples: int, device) -> torch.Tensor:
    """The S-1 offsets of uniform placement; computed by torch.linspace exactly as the reference does
    (nerf_pytorch/utils.py:232) so that the depths are bit-identical."""
    return _device_linspace(-float(std), float(std), n_samples - 1, str(torch.device(device)))


def place_samples(mean, n_samples: int, mode: str, std: float, noise: Optional[torch.Tensor] = None,
                  clip=(2.0, 6.0)) -> torch.Tensor:
    """z [N,S] of sample_points_around_mean (nerf_pytorch/utils.py:220-244); pts are not materialised."""
    mean = _dev(mean, "mean").reshape(-1)
    n = mean.shape[0]
    m = PLACE_MODES[mode]
    if mode == "depth_only":
        n_samples = 1
        offs = None
    elif mode == "uniform":
        offs = uniform_grid(std, n_samples, mean.device) if n_samples > 1 else None
    else:
        if noise is None:
            noise = torch.randn(n, n_samples - 1, device=mean.device)
        offs = _dev(std * noise, "noise")
    z = torch.empty(n, n_samples, device=mean.device)
    with torch.cuda.device(mean.device):
        _lib.check(_lib.lib().b200nerf_place_samples(_p(mean), _p(offs), n, n_samples, m, float(clip[0]), float(clip[1]),
                                                     _p(z), _stream()))
    return z


def points(rays_o, rays_d, z) -> torch.Tensor:
    """pts [N,S,3] = o + d*z, only when somebody asks for it."""
    rays_o, rays_d, z = _dev(rays_o, "rays_o"), _dev(rays_d, "rays_d"), _dev(z, "z")
    n, s = z.shape
    out = torch.empty(n, s, 3, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().b200nerf_points(_p(rays_o), _p(rays_d), _p(z), n, s, _p(out), _stream()))
    return out


def nerf_mlp(pk: PackedNeRF, viewdirs, *, rays_o=None, rays_d=None, z=None, pts=None) -> torch.Tensor:
    """raw [N,S,4] = NeRF(encode(pts), encode(viewdirs)) with pts = o + d*z or explicit
    (trainers/Trainer.py:789-806 + run_nerf_helpers.py:109-134)."""
    viewdirs = _dev(viewdirs, "viewdirs")
    n = viewdirs.shape[0]
    if pts is not None:
        pts = _dev(pts, "pts")
        s = pts.shape[1]
        ro = rd = zz = None
    else:
        ro, rd, zz = _dev(rays_o, "rays_o"), _dev(rays_d, "rays_d"), _dev(z, "z")
        s = zz.shape[1]
    raw = torch.empty(n, s, 4, device=viewdirs.device)
    ws = torch.empty(n + 4, device=viewdirs.device, dtype=torch.int32) if pk.prec == PREC_FAST else None
    model = pk.c_model()
    with torch.cuda.device(viewdirs.device):
        _lib.check(_lib.lib().b200nerf_nerf_query(C.byref(model), _p(ro), _p(rd), _p(viewdirs), _p(zz), _p(pts), n, s, _p(ws),
                                                  _p(raw), _stream()))
    nerf_mlp.last_guard_ws = ws  # PREC_FAST: ws[0] = number of samples re-evaluated in split precision
    return raw


def composite(raw, z, rays_d, white_bkgd=True, noise=None, want_alphas=True):
    """raw2outputs -> (rgb, disp, acc, depth, weights, alphas) (trainers/sampling_trainer.py:153-230)."""
    raw, z, rays_d = _dev(raw, "raw"), _dev(z, "z"), _dev(rays_d, "rays_d")
    n, s = z.shape
    dev = raw.device
    rgb = torch.empty(n, 3, device=dev)
    disp = torch.empty(n, device=dev)
    acc = torch.empty(n, device=dev)
    depth = torch.empty(n, device=dev)
    sw = s if s > 1 else 0  # S == 1: the reference returns empty per-sample tensors
    weights = torch.empty(n, sw, device=dev)
    alphas = torch.empty(n, sw, device=dev) if want_alphas else None
    nz = None if noise is None else _dev(noise, "noise")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200nerf_composite_fwd(_p(raw), _p(z), _p(rays_d), _p(nz), n, s, int(bool(white_bkgd)), _p(rgb),
                                                     _p(disp), _p(acc), _p(depth), _p(weights) if sw else None,
                                                     _p(alphas) if (sw and want_alphas) else None, _stream()))
    return rgb, disp, acc, depth, weights, alphas


def render_depthnet(dn: PackedDepthNet, nerf: PackedNeRF, rays_o, rays_d, viewdirs, n_samples: int, mode: str, std: float,
                    radius=2.0, near=2.0, far=6.0, noise=None, want_weights=True, tile=False, tile_out=None):
    """DepthNet branch of render_rays_test for rays on the device (nerf_utils.py:834-866), one C call.

    ``tile=True`` (or a preallocated ``tile_out`` [N,4]) makes the composite kernel write one 16-byte (r, g, b, disp) pixel per
    ray -- the layout the multi-GPU render all-gathers and render_path ships to the host -- instead of separate maps."""
    rays_o, rays_d, viewdirs = _dev(rays_o, "rays_o"), _dev(rays_d, "rays_d"), _dev(viewdirs, "viewdirs")
    n = rays_o.shape[0]
    dev = rays_o.device
    if mode == "depth_only":
        n_samples, offs = 1, None
    elif mode == "uniform":
        offs = uniform_grid(std, n_samples, dev) if n_samples > 1 else None
    else:
        if noise is None:
            noise = torch.randn(n, n_samples - 1, device=dev)
        offs = _dev(std * noise, "noise")
    s = n_samples
    mean = torch.empty(n, 1, device=dev)
    z = torch.empty(n, s, device=dev)
    raw = torch.empty(n, s, 4, device=dev)
    acc = torch.empty(n, device=dev)
    depth = torch.empty(n, device=dev)
    sw = s if s > 1 else 0
    weights = torch.empty(n, sw, device=dev) if want_weights else None
    ws = torch.empty(n + 4, device=dev, dtype=torch.int32) if nerf.prec == PREC_FAST else None
    model = nerf.c_model()
    common = (_p(dn.wpack), _p(dn.aux), dn.n_hidden, dn.prec, C.byref(model), _p(rays_o), _p(rays_d), _p(viewdirs),
              n, s, PLACE_MODES[mode], _p(offs), float(radius), float(near), float(far), _p(mean), _p(z), _p(raw), _p(ws))
    tail = (_p(acc), _p(depth), _p(weights) if (want_weights and sw) else None, _stream())
    out = dict(acc=acc, depth=depth, weights=weights, z=z, raw=raw, z_mean=mean, guard_ws=ws)
    with torch.cuda.device(dev):
        if tile or tile_out is not None:
            if tile_out is None:
                tile_out = torch.empty(n, 4, device=dev)
            elif tuple(tile_out.shape) != (n, 4) or not tile_out.is_contiguous() or tile_out.dtype != torch.float32 or tile_out.device != dev:
                raise _lib.B200NerfError("tile_out must be a contiguous fp32 [n_rays, 4] tensor on the rays' device")
            _lib.check(_lib.lib().b200nerf_render_depthnet_tile(*common, _p(tile_out), *tail))
            out.update(rgbd=tile_out, rgb=tile_out[:, :3], disp=tile_out[:, 3])
        else:
            rgb = torch.empty(n, 3, device=dev)
            disp = torch.empty(n, device=dev)
            _lib.check(_lib.lib().b200nerf_render_depthnet(*common, _p(rgb), _p(disp), *tail))
            out.update(rgb=rgb, disp=disp)
    return out


def coarse_z(near, far, n_rays: int, n_samples: int, lindisp: bool, t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Stratified coarse depths [N,S] (trainers/Trainer.py:603-627)."""
    near, far = _dev(near, "near").reshape(-1), _dev(far, "far").reshape(-1)
    dev = near.device
    t = _device_linspace(0.0, 1.0, n_samples, str(dev))  # the reference's own t grid
    tr = None if t_rand is None else _dev(t_rand, "t_rand")
    z = torch.empty(n_rays, n_samples, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200nerf_coarse_depths(_p(near), _p(far), _p(t), n_rays, n_samples, int(bool(lindisp)), _p(tr),
                                                     _p(z), _stream()))
    return z


def _u_grid(u, n_samples, dev):
    if u is None:  # det=True: u = linspace(0, 1, N_samples) shared by all rays (run_nerf_helpers.py:258-260)
        return _device_linspace(0.0, 1.0, n_samples, str(torch.device(dev))), 0
    return _dev(u, "u"), 1


def sample_pdf(bins, weights, n_samples: int, u: Optional[torch.Tensor] = None, return_inds: bool = False):
    """Inverse-CDF samples [N,n_samples] (run_nerf_helpers.py:250-293); ``u`` [N,n_samples] overrides the det grid."""
    bins, weights = _dev(bins, "bins"), _dev(weights, "weights")
    n, b = bins.shape
    uu, per_ray = _u_grid(u, n_samples, bins.device)
    out = torch.empty(n, n_samples, device=bins.device)
    inds = torch.empty(n, n_samples, device=bins.device, dtype=torch.int64) if return_inds else None
    with torch.cuda.device(bins.device):
        _lib.check(_lib.lib().b200nerf_sample_pdf(_p(bins), _p(weights), _p(uu), per_ray, n, b, n_samples, _p(out), _p(inds), _stream()))
    return (out, inds) if return_inds else out


def sample_pdf_merge(z_coarse, weights, n_importance: int, u: Optional[torch.Tensor] = None, return_inds: bool = False):
    """(z_samples [N,Nf], z_all [N,Nc+Nf] sorted, inds or None) from the coarse pass (trainers/Trainer.py:668-686)."""
    z_coarse, weights = _dev(z_coarse, "z_coarse"), _dev(weights, "weights")
    n, sc = z_coarse.shape
    dev = z_coarse.device
    uu, per_ray = _u_grid(u, n_importance, dev)
    zs = torch.empty(n, n_importance, device=dev)
    za = torch.empty(n, sc + n_importance, device=dev)
    inds = torch.empty(n, n_importance, device=dev, dtype=torch.int64) if return_inds else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200nerf_sample_pdf_merge(_p(z_coarse), _p(weights), _p(uu), per_ray, n, sc, n_importance, _p(zs),
                                                        _p(inds), _p(za), _stream()))
    return zs, za, inds


def argmax_gather(weights, z, raw=None):
    """(top [N,1] int64, z[top] [N,1], w[top] [N,1], sigmoid(rgb[top]) [N,3]) -- nerf_utils.py:689-690, :806-812."""
    weights, z = _dev(weights, "weights"), _dev(z, "z")
    n, s = weights.shape
    dev = weights.device
    raw = None if raw is None else _dev(raw, "raw")
    idx = torch.empty(n, 1, device=dev, dtype=torch.int64)
    mz = torch.empty(n, 1, device=dev)
    mw = torch.empty(n, 1, device=dev)
    rgb = torch.empty(n, 3, device=dev) if raw is not None else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200nerf_argmax_gather(_p(weights), _p(z), _p(raw), n, s, _p(idx), _p(mz), _p(mw), _p(rgb), _stream()))
    return idx, mz, mw, rgb


def umma_selftest(a_bf16: torch.Tensor, b_bf16: torch.Tensor) -> torch.Tensor:
    """D = A @ B^T through the MLP kernels' operand layout (A [128,K], B [N,K], bf16)."""
    assert a_bf16.dtype == torch.bfloat16 and b_bf16.dtype == torch.bfloat16 and a_bf16.is_cuda
    a, b = a_bf16.contiguous(), b_bf16.contiguous()
    k, n = a.shape[1], b.shape[0]
    d = torch.empty(128, n, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().b200nerf_umma_selftest(a.data_ptr(), b.data_ptr(), d.data_ptr(), k, n, _stream()))
    return d
