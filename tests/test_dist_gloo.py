"""World-size-2 gloo test (CPU) of the ray-sharding host logic used for multi-GPU renders."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_sampling_b200.parallel import shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(lo, hi):
    r = torch.arange(lo, hi, dtype=torch.float32)
    return torch.stack([r, r * 2, r * 3, -r], -1)  # "rgb + disp" tile that encodes the ray index


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_sampling_b200.parallel import render_sharded

    full = render_sharded(_fake_render, n_rays)
    ok = torch.equal(full, _fake_render(0, n_rays))
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 640000, 640001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_render_gathers_identical_image_world2():
    ctx = mp.get_context("spawn")
    for n_rays in (1000, 1001):  # even and ragged split
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(ok for _, ok, _ in res) and all(shape == (n_rays, 4) for _, _, shape in res)


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_sampling_b200.parallel import allreduce_gradients

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[2].grad = None  # a parameter without gradient is skipped
    for i, p in enumerate(params[:2]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    scale = allreduce_gradients(params)
    ok = (scale == 0.5 and torch.equal(params[0].grad, torch.full((3, 5), 3.0)) and torch.equal(params[1].grad, torch.full((7,), 6.0))
          and params[2].grad is None)
    # gradients laid out back to back in one buffer (what training.DepthNetTrainFn.backward produces): reduced in place
    flat = torch.arange(22, dtype=torch.float32) * (rank + 1)
    params[0].grad, params[1].grad = flat[:15].view(3, 5), flat[15:22]
    scale = allreduce_gradients(params)
    ok = ok and scale == 0.5 and torch.equal(flat, torch.arange(22, dtype=torch.float32) * 3) and params[0].grad.data_ptr() == flat.data_ptr()
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    """Flat-buffer SUM all-reduce of DepthNet gradients + the 1/world scale handed to the optimizer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_ray_shard_stitching_ragged():
    """render_path(shard="rays"): the ranks' ray slices travel padded to equal length (all_gather needs equal tiles); the unload
    worker stitches them back by shard_bounds.  Pure host logic: checked here with a fabricated gathered buffer."""
    import numpy as np

    from nerf_sampling_b200.nerf_pytorch import nerf_utils

    class Done:
        def synchronize(self):
            pass

    H, W = 7, 9                      # 63 rays over 4 ranks: 16 + 16 + 16 + 15
    n, world = H * W, 4
    full = np.random.default_rng(0).random((n, 4), dtype=np.float32)
    per = (n + world - 1) // world
    host = torch.full((world, per, 4), -1.0)
    h_rgb, h_disp = nerf_utils._planar(host)        # the slot's layout: rgb plane [world, per, 3] then disp plane [world, per]
    for r in range(world):
        a, b = shard_bounds(n, world, r)
        h_rgb[r, : b - a] = torch.from_numpy(full[a:b, :3])
        h_disp[r, : b - a] = torch.from_numpy(full[a:b, 3])
    rgbs = np.zeros((2, H, W, 3), np.float32)
    disps = np.zeros((2, H, W), np.float32)
    nerf_utils._unload_ray_shards(host, Done(), world, n, H, W, 1, rgbs, disps, None)
    assert np.array_equal(rgbs[1], full[:, :3].reshape(H, W, 3)) and np.array_equal(disps[1], full[:, 3].reshape(H, W))
    assert not rgbs[0].any()

