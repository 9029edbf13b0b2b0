"""Mirror of the render driver ``nerf_sampling/nerf_pytorch/nerf_utils.py`` (:27-255, :393-494, :736-876)."""

from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F

from .. import ops
from . import run_nerf_helpers, utils
from .utils import sample_points_around_mean  # noqa: F401  (re-exported like the reference)


def raw2alpha(raw, dists):
    """alpha = 1 - exp(-relu(sigma) * delta) (nerf_utils.py:27-42); fused in the composite kernel on the render path."""
    return 1.0 - torch.exp(-F.relu(raw) * dists)


def batchify(fn, chunk):
    """nerf_utils.py:45-55."""
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[i : i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)


def prepare_rays(c2w, c2w_staticcam, use_viewdirs, ndc, H, W, K, near, far, rays):
    """Ray batch rows [o(3), d(3), near, far, viewdir(3)] (nerf_utils.py:156-188)."""
    if ndc:
        raise NotImplementedError("NDC rays (LLFF forward-facing scenes) are outside the Blender/DepthNet path")
    if c2w is not None:
        dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else "cuda"
        rays_o, rays_d, viewdirs = ops.get_rays(H, W, K, c2w, dev)
        sh = (H, W, 3)
        if c2w_staticcam is not None:
            rays_o, rays_d, _ = ops.get_rays(H, W, K, c2w_staticcam, dev)
    else:
        rays_o, rays_d = rays
        sh = tuple(rays_d.shape)
        rays_o = rays_o.reshape(-1, 3).float().contiguous()
        rays_d = rays_d.reshape(-1, 3).float().contiguous()
        viewdirs = ops.normalize_dirs(rays_d)
    near_t = torch.full_like(rays_d[..., :1], float(near))
    far_t = torch.full_like(rays_d[..., :1], float(far))
    packed = torch.cat([rays_o, rays_d, near_t, far_t], -1)
    if use_viewdirs:
        packed = torch.cat([packed, viewdirs], -1)
    return packed, rays_o, rays_d, sh


# The reference bounds memory with a 32,768-ray chunk loop (nerf_utils.py:58-85).  On a 180 GB part a whole 800x800 view is
# one pass (raw for 640,000 rays x 192 samples is 2 GB) and the result is chunk-invariant (tested), so consecutive chunks are
# coalesced up to this many rays per pass: fewer launches, no tail effects.  Set to 0 to honour `chunk` literally.
COALESCE_RAYS = 1 << 20


def _batchify(fn, rays_flat, chunk, **kwargs):
    out = {}
    if COALESCE_RAYS and chunk < rays_flat.shape[0]:
        chunk = max(chunk, min(rays_flat.shape[0], COALESCE_RAYS))
    for i in range(0, rays_flat.shape[0], chunk):
        for k, v in fn(rays_flat[i : i + chunk], **kwargs).items():
            out.setdefault(k, []).append(v)
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in out.items()}


def batchify_rays_test(rays_flat, chunk=1024 * 32, **kwargs):
    """Chunk loop of nerf_utils.py:73-85 (no allocator flush, no sync per chunk)."""
    return _batchify(render_rays_test, rays_flat, chunk, **kwargs)


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    return _batchify(render_rays, rays_flat, chunk, **kwargs)


def _finish(all_ret, sh, rays_o, rays_d):
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    head = ["depth_net_rgb_map", "depth_net_disp_map"]
    extras = {k: v for k, v in all_ret.items() if k not in head}
    extras["rays_o"], extras["rays_d"] = rays_o, rays_d
    return [all_ret[k] for k in head] + [extras]


def render_test(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0.0, far=1.0, use_viewdirs=False,
                c2w_staticcam=None, **kwargs):
    """[rgb, disp, extras] for a full view or a ray batch (nerf_utils.py:191-255)."""
    packed, rays_o, rays_d, sh = prepare_rays(c2w, c2w_staticcam, use_viewdirs, ndc, H, W, K, near, far, rays)
    return _finish(batchify_rays_test(packed, chunk, **kwargs), sh, rays_o, rays_d)


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0.0, far=1.0, use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """Training-time render (nerf_utils.py:88-153)."""
    packed, rays_o, rays_d, sh = prepare_rays(c2w, c2w_staticcam, use_viewdirs, ndc, H, W, K, near, far, rays)
    return _finish(batchify_rays(packed, chunk, **kwargs), sh, rays_o, rays_d)


def _write_png(path: str, rgb8) -> None:
    """Minimal 8-bit RGB PNG encoder (zlib from the standard library; the reference uses imageio.imwrite)."""
    import struct
    import zlib

    import numpy as np

    h, w, c = rgb8.shape
    assert c == 3 and rgb8.dtype == np.uint8
    rows = np.concatenate([np.zeros((h, 1), np.uint8), rgb8.reshape(h, w * 3)], axis=1)  # filter type 0 per scanline

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(rows.tobytes(), 3)) + chunk(b"IEND", b""))


def write_video(path: str, rgbs) -> None:
    """``imageio.mimwrite(video.mp4)`` of Trainer.py:224-229 when imageio (+ffmpeg) exists; the PNG frames are always written."""
    try:
        import imageio

        imageio.mimwrite(path, run_nerf_helpers.to8b(rgbs), fps=30, quality=8)
    except Exception as e:  # no encoder in this environment
        print(f"[b200nerf] video not written ({type(e).__name__}); frames are in {os.path.dirname(path)}")


class _HostRing:
    """Page-locked staging slots for finished views (cudaHostAlloc is slow, so the ring is cached per shape and reused by
    later calls); a slot is recycled once the worker that unloads it into the result arrays has finished."""

    _cache = {}

    def __init__(self, shape, slots):
        self.bufs = [torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in range(slots)]
        self.busy = [None] * slots   # the future that still reads the slot

    @classmethod
    def get(cls, shape, slots=4):
        key = (tuple(shape), slots)
        ring = cls._cache.get(key)
        if ring is None:
            ring = cls._cache[key] = cls(shape, slots)
        return ring

    def acquire(self, i):
        k = i % len(self.bufs)
        if self.busy[k] is not None:
            self.busy[k].result()
            self.busy[k] = None
        return k, self.bufs[k]


def _plain_depthnet_mode(trainer, render_kwargs):
    return not (trainer.compare_nerf or trainer.use_nerf_max_pts or trainer.use_full_nerf
                or getattr(trainer, "save_scene_data", False)) and render_kwargs.get("use_viewdirs", False)


def _render_tile(H, W, K, c2w, lo, hi, chunk, render_kwargs, want_extras, out=None):
    """rgb|disp tile [hi-lo, 4] of rays [lo, hi) of one view (+ extras dict of the slice when asked).

    Plain DepthNet mode is ONE C call whose composite kernel writes the 16-byte pixels itself
    (``b200nerf_render_depthnet_tile``); the NeRF-comparison modes go through ``render_test`` on the ray slice."""
    trainer = render_kwargs["trainer"]
    dev = torch.device("cuda", torch.cuda.current_device())
    rays_o, rays_d, viewdirs = ops.get_rays(H, W, K, c2w, dev)
    if (lo, hi) != (0, H * W):
        rays_o, rays_d, viewdirs = rays_o[lo:hi], rays_d[lo:hi], viewdirs[lo:hi]
    if _plain_depthnet_mode(trainer, render_kwargs) and not want_extras:
        dn = render_kwargs["depth_network"]
        net = render_kwargs["network_fine"] if render_kwargs.get("network_fine") is not None else render_kwargs["network_fn"]
        res = ops.render_depthnet(dn.packed(), net.packed(), rays_o, rays_d, viewdirs, trainer.n_depth_samples,
                                  trainer.sampling_mode, trainer.distance, radius=float(dn.sphere_radius), near=float(dn.near),
                                  far=float(dn.far), want_weights=False, tile=True, tile_out=out)
        return res["rgbd"], {}
    kw = dict(render_kwargs)
    near, far = kw.pop("near"), kw.pop("far")
    ndc, use_viewdirs = kw.pop("ndc", False), kw.pop("use_viewdirs", False)
    n = hi - lo
    packed = torch.cat([rays_o, rays_d, torch.full((n, 1), float(near), device=dev), torch.full((n, 1), float(far), device=dev)]
                       + ([viewdirs] if use_viewdirs else []), -1)
    if ndc:
        raise NotImplementedError("NDC rays (LLFF forward-facing scenes) are outside the Blender/DepthNet path")
    ret = batchify_rays_test(packed, chunk, **kw)
    tile = out if out is not None else torch.empty(n, 4, device=dev)
    tile[:, :3] = ret["depth_net_rgb_map"]
    tile[:, 3] = ret["depth_net_disp_map"]
    return tile, ret


def render_path(render_poses, hwf, K, chunk, render_kwargs, step=0, wandb_log=False, save_scene_data=False, gt_imgs=None,
                savedir=None, render_factor=0, group=None, shard=None, dst=None):
    """Render a list of camera poses (nerf_utils.py:258-360) -> (rgbs [n,H,W,3], disps [n,H,W], mean PSNR) as numpy.

    Same results as the reference's loop, different schedule: view k+1 is rendered while view k's pixels travel to pinned
    host memory on a side stream, and unloading / PNG encoding (``savedir``) runs in worker threads, so the GPU never waits
    for ``.cpu().numpy()``, PSNR arithmetic or file I/O.

    Multi-GPU (one process per GPU, ``torch.distributed`` initialised): ``shard="views"`` gives view i to rank
    ``i % world`` (throughput: BASELINE config #3), ``shard="rays"`` gives every rank the contiguous ray slice
    ``parallel.shard_bounds(H*W, world, rank)`` of EVERY view (latency / strong scaling).  Either way a rank writes its
    finished pixels as one [n_local, 4] rgb|disp tile, the tiles are all-gathered with NCCL on a side stream while the next
    view renders (double-buffered), and every rank returns the full image stack.  Rays are independent, so the sharded images
    are bit-identical to the single-GPU ones.  ``shard=None`` (default) renders every pose on the calling rank.  ``dst=r`` keeps the
    host copy (and the PNG / PSNR work) on rank ``r`` only -- the other ranks return empty image stacks -- which is what a driver
    script that saves from rank 0 wants; ``dst=None`` returns the full stack on every rank.

    ``wandb_log`` switches ``trainer.compare_nerf`` on like the reference (:294-295; the ray plots themselves are the host
    application's); ``save_scene_data`` collects ``depth_net_pts`` / ``depth_net_weights`` for this call only."""
    import concurrent.futures

    import numpy as np
    import torch.distributed as dist

    from .. import parallel

    H, W, focal = hwf
    H, W = int(H), int(W)
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    poses = torch.as_tensor(np.asarray(render_poses.detach().cpu()) if isinstance(render_poses, torch.Tensor) else np.asarray(render_poses),
                            dtype=torch.float32)   # one D2H for all poses: the ray kernel takes the matrix by value
    n = poses.shape[0]
    n_rays = H * W
    trainer = render_kwargs["trainer"]
    if wandb_log:
        trainer.compare_nerf = True
    prev_ssd = getattr(trainer, "save_scene_data", False)
    if save_scene_data:
        trainer.save_scene_data = True
    world, rank = 1, 0
    if shard is not None:
        if shard not in ("views", "rays"):
            raise ValueError("shard must be None, 'views' or 'rays'")
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
    world = max(world, 1)
    sharded = world > 1
    if sharded and (save_scene_data or trainer.compare_nerf):
        raise NotImplementedError("per-sample extras (save_scene_data / compare_nerf) are not gathered across ranks")

    dev = torch.device("cuda", torch.cuda.current_device())
    main = torch.cuda.current_stream()
    side = _side_stream(dev)
    to_host = (not sharded) or dst is None or dst == rank
    rgbs = np.empty((n if to_host else 0, H, W, 3), np.float32)
    disps = np.empty((n if to_host else 0, H, W), np.float32)
    pool = concurrent.futures.ThreadPoolExecutor(max_workers=4)
    all_pts, all_weights, mses = [], [], []
    if savedir is not None:
        os.makedirs(savedir, exist_ok=True)

    def unload(views, host, ev):
        """worker: wait for the D2H, copy the planar rgb / disp slabs into the result arrays, encode PNGs."""
        ev.synchronize()
        h_rgb, h_disp = _planar(host)
        h_rgb, h_disp = h_rgb.numpy(), h_disp.numpy()
        for j, i in enumerate(views):
            if i >= n:
                continue
            rgbs[i] = h_rgb[j].reshape(H, W, 3)
            disps[i] = h_disp[j].reshape(H, W)
            if savedir is not None:
                _write_png(os.path.join(savedir, "{:03d}.png".format(i)), run_nerf_helpers.to8b(rgbs[i]))

    try:
        if not sharded:
            ring = _HostRing.get((1, n_rays, 4))
            for i in range(n):
                want_extras = trainer.compare_nerf or save_scene_data
                tile, extras = _render_tile(H, W, K, poses[i, :3, :4], 0, n_rays, chunk, render_kwargs, want_extras)
                k, host = ring.acquire(i)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    _to_host_planar(host, tile.view(1, n_rays, 4))
                    ev = torch.cuda.Event()
                    ev.record(side)
                tile.record_stream(side)
                ring.busy[k] = pool.submit(unload, [i], host, ev)
                if trainer.compare_nerf and extras.get("max_z_vals") is not None:
                    # F.mse_loss(max_z [H,W,1], z_vals [H,W,S]) broadcasts over the samples in the reference (:311-316)
                    mses.append(torch.mean((extras["max_z_vals"] - extras["depth_net_z_vals"]) ** 2))
                if save_scene_data and savedir is not None:
                    all_pts.append(extras["depth_net_pts"].reshape(-1, 3))
                    all_weights.append(extras["depth_net_weights"].reshape(-1))
        else:
            rounds = (n + world - 1) // world if shard == "views" else n
            lo, hi = (0, n_rays) if shard == "views" else parallel.shard_bounds(n_rays, world, rank)
            per = n_rays if shard == "views" else (n_rays + world - 1) // world   # padded tile rows (ragged ray shards)
            ring = _HostRing.get((world, per, 4), slots=2)
            tiles = [torch.zeros(per, 4, device=dev) for _ in range(2)]
            gathered = [torch.empty(world, per, 4, device=dev) for _ in range(2)]
            reusable = [None, None]   # event: the D2H that read gathered[b] (and the gather that read tiles[b]) has finished
            for r in range(rounds):
                b = r & 1
                view = r * world + rank if shard == "views" else r
                if reusable[b] is not None:
                    main.wait_event(reusable[b])
                if view < n:
                    _render_tile(H, W, K, poses[view, :3, :4], lo, hi, chunk, render_kwargs, False, out=tiles[b][: hi - lo])
                k, host = ring.acquire(r)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    dist.all_gather_into_tensor(gathered[b].view(world * per, 4), tiles[b], group=group)
                    if to_host:
                        _to_host_planar(host, gathered[b])
                    ev = torch.cuda.Event()
                    ev.record(side)
                reusable[b] = ev
                if not to_host:
                    continue
                if shard == "views":
                    ring.busy[k] = pool.submit(unload, [r * world + q for q in range(world)], host, ev)
                else:
                    ring.busy[k] = pool.submit(_unload_ray_shards, host, ev, world, n_rays, H, W, r, rgbs, disps, savedir)
        side.synchronize()
        for fut in ring.busy:
            if fut is not None:
                fut.result()
        ring.busy = [None] * len(ring.busy)
    finally:
        pool.shutdown(wait=True)
        trainer.save_scene_data = prev_ssd

    total_psnr = 0.0
    if gt_imgs is not None and render_factor == 0 and to_host:
        lines = []
        for i in range(n):
            psnr = -10.0 * np.log10(np.mean(np.square(rgbs[i] - np.asarray(gt_imgs[i])[..., :3])))
            total_psnr += psnr
            lines.append(f"{i:03d}.png, PSNR: {psnr}" + (f", MSE: {float(mses[i])}" if i < len(mses) else ""))
        if savedir is not None:
            with open(os.path.join(savedir, "psnr.txt"), "a") as f:
                f.write("\n".join(lines) + f"\nAvg of {n} images:\nPSNR: {total_psnr / n}\n")
                if mses:
                    f.write(f"MSE: {float(sum(mses)) / n}")
    if save_scene_data and savedir is not None:
        torch.save({"all_pts": torch.cat(all_pts), "all_weights": torch.cat(all_weights)}, os.path.join(savedir, "scene_data.pt"))
    return rgbs, disps, total_psnr / max(n, 1)


def _planar(host):
    """The two planes of a pinned slot [V, per, 4 floats]: rgb [V, per, 3] followed by disp [V, per].  The device-side rgb|disp
    pixels are de-interleaved by the copy itself (a strided device read on the copy stream), so the host only memcpy's."""
    v, per, _ = host.shape
    flat = host.view(-1)
    return flat[: v * per * 3].view(v, per, 3), flat[v * per * 3:].view(v, per)


def _to_host_planar(host, px):
    """px [V, per, 4] on the device -> the slot's planes, asynchronously on the current stream."""
    h_rgb, h_disp = _planar(host)
    h_rgb.copy_(px[..., :3], non_blocking=True)
    h_disp.copy_(px[..., 3], non_blocking=True)


def _unload_ray_shards(host, ev, world, n_rays, H, W, view, rgbs, disps, savedir):
    """worker for shard="rays": stitch the ranks' ray slices (padded to equal length for the all-gather) into view ``view``."""
    from .. import parallel

    ev.synchronize()
    h_rgb, h_disp = _planar(host)
    h_rgb, h_disp = h_rgb.numpy(), h_disp.numpy()
    out_rgb, out_disp = rgbs[view].reshape(n_rays, 3), disps[view].reshape(n_rays)
    for q in range(world):
        a, b = parallel.shard_bounds(n_rays, world, q)
        out_rgb[a:b] = h_rgb[q, : b - a]
        out_disp[a:b] = h_disp[q, : b - a]
    if savedir is not None:
        _write_png(os.path.join(savedir, "{:03d}.png".format(view)), run_nerf_helpers.to8b(rgbs[view]))


_SIDE_STREAMS = {}


def _side_stream(dev):
    """One persistent copy / collective stream per device (stream creation per call churns the driver)."""
    s = _SIDE_STREAMS.get(dev)
    if s is None:
        s = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


def sample_as_in_NeRF(ray_batch, network_fn, network_fine, network_query_fn, N_samples, trainer, perturb, raw_noise_std,
                      lindisp, white_bkgd, kwargs, pytest):
    """Vanilla coarse + fine pass (nerf_utils.py:497-611)."""
    n_rays = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    c = trainer.sample_coarse_points(near=near, far=far, perturb=perturb, N_rays=n_rays, N_samples=N_samples,
                                     viewdirs=viewdirs, network_fn=network_fn, network_query_fn=network_query_fn,
                                     rays_o=rays_o, rays_d=rays_d, raw_noise_std=raw_noise_std, white_bkgd=white_bkgd,
                                     pytest=pytest, lindisp=lindisp, kwargs=kwargs)
    f = trainer.sample_fine_points(z_vals=c[5], weights=c[6], perturb=perturb, pytest=pytest, rays_d=rays_d, rays_o=rays_o,
                                   rgb_map=c[0], disp_map=c[1], acc_map=c[2], network_fn=network_fn,
                                   network_fine=network_fine, network_query_fn=network_query_fn, viewdirs=viewdirs,
                                   raw_noise_std=raw_noise_std, white_bkgd=white_bkgd)
    (_, _, _, fine_rgb, fine_disp, _, fine_raw, fine_z, fine_pts, fine_density, fine_alphas, fine_weights) = f
    return fine_density, fine_z, fine_pts, fine_rgb, fine_weights, fine_alphas, fine_disp, fine_raw


def render_rays_test(ray_batch, network_fn, network_query_fn, N_samples, trainer, retraw=True, lindisp=False, perturb=0.0,
                     N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.0, verbose=False, pytest=False,
                     **kwargs):
    """Inference render of one ray chunk (nerf_utils.py:736-876).

    Differences from the reference, both documented in INTEGRATION.md: per-sample extras stay on the device
    (the reference ``.cpu()``s ~0.8 GB per 800x800 view), and ``depth_net_pts`` is only materialised when the
    trainer asks for scene data."""
    rays_o, rays_d = ray_batch[:, 0:3].contiguous(), ray_batch[:, 3:6].contiguous()
    viewdirs = ray_batch[:, -3:].contiguous() if ray_batch.shape[-1] > 8 else None
    ret = {}
    nerf_modes = trainer.compare_nerf or trainer.use_nerf_max_pts or trainer.use_full_nerf
    if nerf_modes:
        (fine_density, fine_z, fine_pts, fine_rgb, fine_weights, fine_alphas, fine_disp, fine_raw) = sample_as_in_NeRF(
            ray_batch=ray_batch, N_samples=N_samples, network_fn=network_fn, network_fine=network_fine,
            network_query_fn=network_query_fn, trainer=trainer, perturb=perturb, raw_noise_std=raw_noise_std,
            lindisp=lindisp, white_bkgd=white_bkgd, pytest=pytest, kwargs=kwargs)
        top, max_z, max_w, max_rgb = ops.argmax_gather(fine_weights, fine_z, fine_raw)
        max_pts = ops.points(rays_o, rays_d, max_z)
        ret["max_z_vals"], ret["max_pts"], ret["max_weights"] = max_z, max_pts, max_w

    if trainer.use_nerf_max_pts:
        rgb_map, disp, weights, pts, z = max_rgb, torch.zeros_like(max_rgb), max_w, max_pts, max_z
    elif trainer.use_full_nerf:
        rgb_map, disp, weights, pts, z = fine_rgb, fine_disp, fine_weights, fine_pts, fine_z
    else:
        net = network_fine if network_fine is not None else network_fn
        out = ops.render_depthnet(kwargs["depth_network"].packed(), net.packed(), rays_o, rays_d, viewdirs,
                                  trainer.n_depth_samples, trainer.sampling_mode, trainer.distance,
                                  radius=float(kwargs["depth_network"].sphere_radius),
                                  near=float(kwargs["depth_network"].near), far=float(kwargs["depth_network"].far))
        rgb_map, disp, weights, z = out["rgb"], out["disp"], out["weights"], out["z"]
        pts = ops.points(rays_o, rays_d, z) if getattr(trainer, "save_scene_data", False) else None
        if retraw:
            ret["raw"] = out["raw"]
    ret["depth_net_rgb_map"] = rgb_map
    ret["depth_net_weights"] = weights
    ret["depth_net_disp_map"] = disp
    ret["depth_net_z_vals"] = z
    if pts is not None:
        ret["depth_net_pts"] = pts
    return ret


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, trainer, retraw=True, lindisp=False, perturb=0.0,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.0, verbose=False, pytest=False, **kwargs):
    """Training render (nerf_utils.py:614-733): hierarchical target depth from the frozen NeRFs (tensor-core kernels, no
    gradient) + DepthNet with one sample per ray.  With gradients enabled ``depth_net_rgb_map`` and ``depth_net_z_vals``
    are attached to a graph that reaches DepthNet's parameters (training.py)."""
    rays_o, rays_d = ray_batch[:, 0:3].contiguous(), ray_batch[:, 3:6].contiguous()
    viewdirs = ray_batch[:, -3:].contiguous() if ray_batch.shape[-1] > 8 else None
    with torch.no_grad():
        (_, fine_z, _, _, fine_weights, _, _, fine_raw) = sample_as_in_NeRF(
            ray_batch=ray_batch, N_samples=N_samples, network_fn=network_fn, network_fine=network_fine,
            network_query_fn=network_query_fn, trainer=trainer, perturb=perturb, raw_noise_std=raw_noise_std, lindisp=lindisp,
            white_bkgd=white_bkgd, pytest=pytest, kwargs=kwargs)
        _, max_z, _, _ = ops.argmax_gather(fine_weights, fine_z, fine_raw)
    z_dn = kwargs["depth_network"](rays_o, rays_d)
    net = network_fine if network_fine is not None else network_fn
    raw = net.query(viewdirs, rays_o=rays_o, rays_d=rays_d, z=z_dn)
    if raw.requires_grad:
        from .. import training

        rgb, disp = training.CompositeSingleFn.apply(raw, z_dn, rays_d)
    else:
        rgb, disp, *_ = ops.composite(raw, z_dn, rays_d, white_bkgd=True)
    ret = {"depth_net_rgb_map": rgb, "depth_net_disp_map": disp, "depth_net_z_vals": z_dn, "max_z_vals": max_z,
           "depth_net_pts": ops.points(rays_o, rays_d, z_dn), "max_pts": ops.points(rays_o, rays_d, max_z)}
    if retraw:
        ret["raw"] = raw
    return ret


def create_nerf(args, model):
    """Coarse/fine NeRF + optimizer + render kwargs (nerf_utils.py:393-494)."""
    embed_fn, input_ch = run_nerf_helpers.get_embedder(args.multires, args.i_embed, args.input_dims_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = run_nerf_helpers.get_embedder(args.multires_views, args.i_embed, args.input_dims_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    mk = lambda d, w: model(D=d, W=w, input_ch=input_ch, output_ch=output_ch, skips=skips,  # noqa: E731
                            input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(args.device)
    model_nerf = mk(args.netdepth, args.netwidth)
    grad_vars = list(model_nerf.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = mk(args.netdepth_fine, args.netwidth_fine)
        grad_vars += list(model_fine.parameters())

    def network_query_fn(inputs, viewdirs, network_fn):
        return args.run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=args.netchunk)

    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    start = 0
    if args.ft_path is not None and args.ft_path != "None":
        ckpts = [args.ft_path]
    else:
        d = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if "tar" in f] if os.path.isdir(d) else []
    if len(ckpts) > 0 and not args.no_reload:
        ckpt = torch.load(ckpts[-1], map_location=args.device, weights_only=False)
        start = ckpt["global_step"]
        utils.load_nerf(model_nerf, model_fine, optimizer, ckpt)

    render_kwargs_train = {
        "network_query_fn": network_query_fn, "perturb": args.perturb, "N_importance": args.N_importance,
        "network_fine": model_fine, "N_samples": args.N_samples, "network_fn": model_nerf,
        "use_viewdirs": args.use_viewdirs, "white_bkgd": args.white_bkgd, "raw_noise_std": args.raw_noise_std,
        "trainer": args,
    }
    if args.dataset_type != "llff" or getattr(args, "no_ndc", False):
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test["perturb"] = False
    render_kwargs_test["raw_noise_std"] = 0.0
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
