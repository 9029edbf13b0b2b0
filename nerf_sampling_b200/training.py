"""Autograd plumbing of the training render (config #5): torch.autograd.Function shells around the C-ABI training
entry points of ``csrc/train.cu``.  No arithmetic happens here; tensors are marshalled, workspaces allocated.

Reference: ``Trainer.core_optimization_loop`` (nerf_pytorch/trainers/Trainer.py:506-544) back-propagates
``mse(z_dn, max_z)`` and ``mse(rgb, target)`` into DepthNet only: the NeRFs are frozen, so the colour loss reaches
DepthNet through ``rgb = sigmoid(raw_rgb(o + d * z_dn))`` (one sample per ray, nerf_utils.py:692-715).
"""

from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptrs(tensors: List[torch.Tensor]):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _ints(vals: List[int]):
    return (C.c_int * len(vals))(*vals)


def depthnet_arch(module):
    """(branch widths, cat widths) of a DepthNet shell, from its Linear layers."""
    hidden = [m.out_features for m in module.origin_layers]
    cat = [m.out_features for m in module.cat_layers if isinstance(m, torch.nn.Linear)]
    return hidden, cat


def depthnet_params(module) -> List[torch.nn.Parameter]:
    """Parameters in the order ``b200nerf_depthnet_train_fwd`` expects (= state_dict order)."""
    out = []
    for seq in (module.origin_layers, module.direction_layers, module.intersection_layers, module.cat_layers, module.to_depth):
        for m in seq:
            if isinstance(m, torch.nn.Linear):
                out += [m.weight, m.bias]
    return out


class DepthNetTrainFn(torch.autograd.Function):
    """z [N,1] = DepthNet(rays_o, rays_d) in literal per-layer fp32 form, differentiable w.r.t. the parameters."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, arch, radius, near, far, *params):
        hidden, cat = arch
        L = _lib.lib()
        n = rays_o.shape[0]
        dev = rays_o.device
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise _lib.B200NerfError("DepthNet parameters must be contiguous fp32 CUDA tensors")
        ws = torch.empty(L.b200nerf_depthnet_train_ws_floats(n, len(hidden), _ints(hidden), len(cat), _ints(cat)), device=dev)
        z = torch.empty(n, 1, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.b200nerf_depthnet_train_fwd(_ptrs(list(params)), len(hidden), _ints(hidden), len(cat), _ints(cat),
                                                     rays_o.data_ptr(), rays_d.data_ptr(), n, float(radius), float(near),
                                                     float(far), ws.data_ptr(), z.data_ptr(), _stream()))
        ctx.ws, ctx.arch, ctx.nf, ctx.n = ws, arch, (float(near), float(far)), n
        ctx.save_for_backward(*params)
        return z

    @staticmethod
    def backward(ctx, dz):
        params = ctx.saved_tensors
        hidden, cat = ctx.arch
        L = _lib.lib()
        dz = dz.contiguous().float()
        grads = [torch.empty_like(p) for p in params]
        with torch.cuda.device(dz.device):
            _lib.check(L.b200nerf_depthnet_train_bwd(_ptrs(list(params)), len(hidden), _ints(hidden), len(cat), _ints(cat), ctx.n,
                                                     ctx.nf[0], ctx.nf[1], ctx.ws.data_ptr(), dz.data_ptr(), _ptrs(grads), _stream()))
        # the activations in ctx.ws stay valid: the reference back-propagates twice (retain_graph=True, Trainer.py:537-538)
        return (None, None, None, None, None, None) + tuple(grads)


class NerfPointFn(torch.autograd.Function):
    """raw [N,1,4] of the frozen NeRF at p = o + d z (one sample per ray, fp32), differentiable w.r.t. z."""

    @staticmethod
    def forward(ctx, z, rays_o, rays_d, viewdirs, *params):
        L = _lib.lib()
        n = z.shape[0]
        dev = z.device
        ws = torch.empty(L.b200nerf_nerf_point_ws_floats(n), device=dev)
        raw = torch.empty(n, 1, 4, device=dev)
        draw = torch.empty(n, 4, device=dev)
        zz = z.detach().reshape(-1).contiguous().float()
        with torch.cuda.device(dev):
            _lib.check(L.b200nerf_nerf_point_jvp(_ptrs(list(params)), rays_o.data_ptr(), rays_d.data_ptr(), viewdirs.data_ptr(),
                                                 zz.data_ptr(), n, ws.data_ptr(), raw.data_ptr(), draw.data_ptr(), _stream()))
        ctx.save_for_backward(draw)
        ctx.zshape = z.shape
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        (draw,) = ctx.saved_tensors
        gz = (g_raw.reshape(-1, 4) * draw).sum(-1).reshape(ctx.zshape)  # chain rule over the four outputs
        return (gz, None, None, None) + (None,) * 24


class CompositeSingleFn(torch.autograd.Function):
    """raw2outputs with S == 1 (the reference's empty-interval quirk, sampling_trainer.py:178-180,220-221):
    rgb = sigmoid(raw rgb), disp = 1e10, differentiable w.r.t. raw."""

    @staticmethod
    def forward(ctx, raw, z, rays_d):
        from . import ops

        rgb, disp, *_ = ops.composite(raw.detach(), z.detach(), rays_d, white_bkgd=True)
        ctx.save_for_backward(rgb)
        return rgb, disp

    @staticmethod
    def backward(ctx, g_rgb, g_disp):
        (rgb,) = ctx.saved_tensors
        g = torch.zeros(rgb.shape[0], 1, 4, device=rgb.device)
        g[:, 0, :3] = g_rgb * rgb * (1.0 - rgb)
        return g, None, None


def nerf_params(module) -> List[torch.nn.Parameter]:
    """The 24 NeRF tensors in the order of ``b200nerf_nerf_pack``."""
    from .packing import NERF_KEYS

    sd = dict(module.named_parameters())
    return [sd[k] for k in NERF_KEYS]


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) on the library's fused kernel, one launch per tensor.
    ``grad_scale`` multiplies every gradient first (1 / world_size after a sum all-reduce)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        L = _lib.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            rows, steps, dev = [], set(), None
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                steps.add(int(st["step"]))
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                rows.append((p, g, st))
                dev = p.device
            if not rows:
                continue
            with torch.cuda.device(dev):
                if len(steps) == 1:
                    # one launch for the whole group: a small pointer table travels to the device every step (the
                    # gradient tensors are new objects after every backward)
                    # pinned staging buffers, two in rotation: a buffer is rewritten only after its previous upload ran
                    ring = self.__dict__.setdefault("_table_ring", [])
                    if len(ring) < 2 or ring[0][0].shape[0] != len(rows):
                        ring[:] = [[torch.empty(len(rows), 5, dtype=torch.int64).pin_memory(), None] for _ in range(2)]
                    self._table_turn = (getattr(self, "_table_turn", 0) + 1) % 2
                    host, ev = ring[self._table_turn]
                    if ev is not None:
                        ev.synchronize()
                    host.view(-1).numpy()[:] = [x for p, g, st in rows for x in
                                                (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())]
                    table = host.to(dev, non_blocking=True)
                    ring[self._table_turn][1] = torch.cuda.Event()
                    ring[self._table_turn][1].record()
                    _lib.check(L.b200nerf_adam_step_multi(table.data_ptr(), len(rows), float(group["lr"]), float(b1), float(b2),
                                                          float(group["eps"]), steps.pop(), float(grad_scale), _stream()))
                    self._keep = (table, [g for _, g, _ in rows])  # alive until the next step: the launch is asynchronous
                else:
                    for p, g, st in rows:
                        _lib.check(L.b200nerf_adam_step(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                                        p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                        int(st["step"]), float(grad_scale), _stream()))
        return None
