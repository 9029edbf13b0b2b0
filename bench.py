#!/usr/bin/env python
"""Headline benchmark of the render_rays hot path: rays/sec for 800x800 DepthNet renders (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # ours (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)

One step = one 800x800 synthetic Blender-style view per GPU (640,000 rays): DepthNet -> 64 uniform samples around the
predicted depth -> positional encoding + 8x256 skip@4 NeRF MLP -> raw2outputs, random-init weights (seed 42), default
precision (fp16 single pass + split-precision guard band: meets the 1e-3 max-abs contract; --prec selects the others).  With N > 1 every rank renders its own contiguous 640,000-ray slice of an N-view batch
(weak scaling) and the image tiles are all-gathered with NCCL inside the timed region.  Prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 800
S = 64
DISTANCE = 0.1
NERF_FLOP_PER_POINT = 1_186_816          # 2 * 593,408 MAC, literal network
DEPTHNET_FLOP_PER_RAY = 6_660_608        # literal network (the folded inference form executes 1,309,184)
COMPOSITE_BYTES_PER_RAY = 24 * S + 36
# dram__bytes_read.sum + dram__bytes_write.sum of the NeRF MLP kernel, one 800x800x64 launch, from the committed
# ncu --set full captures (profiles/r1c_ncu_summary.md: fast kernel 195.1 MB read + 604.0 MB written, guard-band launch
# 30.9 + 1.1 MB; profiles/r1a_ncu_summary.md for the split mode); algorithmic I/O is 819 MB (z in, raw out)
NCU_DRAM_BYTES_PER_LAUNCH = {"fast": 831.2e6, "fp16": 799.2e6, "split": 816.3e6}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_sustained=float(p["bf16_tflops_sustained"]), tf_burst=float(p["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tf_sustained=1400.0, tf_burst=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def pose_spherical(theta: float, phi: float, radius: float) -> torch.Tensor:
    """Camera-to-world matrix of the Blender test orbit (the reference's load_blender.py:8-43): translate along z,
    rotate by phi about x, by theta about y, then the Blender axis flip."""
    import numpy as np

    ph, th = phi / 180.0 * np.pi, theta / 180.0 * np.pi
    t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=np.float32)
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], dtype=np.float32)
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return torch.from_numpy(flip @ (rt @ (rp @ t)))


def pose_for_step(i: int):
    """Pose i of the 200-view test orbit (theta = -180 + 1.8 i, phi = -30, r = 4.031), 3x4."""
    return pose_spherical(-180.0 + 1.8 * (i % 200), -30.0, 4.031)[:3, :4].contiguous()


def intrinsics():
    import numpy as np

    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])


def build_models(device, prec):
    """Random-init weights, seed 42, in the reference's construction order (coarse NeRF, fine NeRF, DepthNet)."""
    from nerf_sampling_b200.depth_nets import DepthNet
    from nerf_sampling_b200.nerf_pytorch.run_nerf_helpers import NeRF

    torch.manual_seed(42)
    mk = lambda: NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)  # noqa: E731
    coarse, fine = mk(), mk()
    dn = DepthNet(hidden_sizes=[256] * 10, cat_hidden_sizes=[256] * 10, sphere_radius=2.0)
    for m in (coarse, fine, dn):
        m.precision = prec
        m.to(device)
    return coarse, fine, dn


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own algorithm (oracle/nerf_oracle.py, pinned bit-exactly to the reference) on the host CPU."""
    if rank != 0:
        return
    from oracle import nerf_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    coarse, fine, dn = O.init_models(42)
    sample = args.ref_rays
    K = intrinsics()
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            packed, *_ = O.prepare_rays(H, W, K, c2w=pose_for_step(i))
            sel = packed[(H // 2) * W : (H // 2) * W + sample]  # contiguous slice from the image centre
            t0 = time.perf_counter()
            O.render_rays_test(sel, coarse, fine, dn, n_depth_samples=S, sampling_mode="uniform", distance=DISTANCE)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    val = sample * len(times) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "rays_per_sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"800x800 view, DepthNet + {S} uniform samples/ray (BASELINE config #2), host CPU", "rays_per_step": sample,
                   "note": "each step is a bounded sample of the view (contiguous rays from the image centre)"},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} rays x {S} samples per step, {len(times)} steps"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
def cpu_baseline_sample(seconds_target=15.0):
    from oracle import nerf_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    coarse, fine, dn = O.init_models(42)
    K = intrinsics()
    with torch.no_grad():
        packed, *_ = O.prepare_rays(H, W, K, c2w=pose_for_step(0))
        base = (H // 2) * W
        t0 = time.perf_counter()
        O.render_rays_test(packed[base : base + 1024], coarse, fine, dn, n_depth_samples=S, sampling_mode="uniform", distance=DISTANCE)
        rate = 1024 / (time.perf_counter() - t0)
        n = int(min(32768, max(2048, rate * seconds_target)))
        t0 = time.perf_counter()
        O.render_rays_test(packed[base : base + n], coarse, fine, dn, n_depth_samples=S, sampling_mode="uniform", distance=DISTANCE)
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} contiguous rays of the 800x800 view x {S} samples, one pass after a 1024-ray warm-up ({dt:.1f} s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--prec", default="fast", choices=["fast", "fp16", "split", "bf16"],
                    help="NeRF MLP precision: fast = fp16 single-pass + split-precision guard band (meets the 1e-3 contract), "
                         "fp16 = single-pass only, split = bf16 hi+lo everywhere, bf16 = legacy single-pass on the exact kernel")
    ap.add_argument("--ref-rays", type=int, default=4096, help="rays per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    import nerf_sampling_b200 as pkg
    from nerf_sampling_b200 import _lib, ops
    import ctypes as C

    from nerf_sampling_b200.packing import PREC_BF16, PREC_FAST, PREC_FP16, PREC_SPLIT

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    pkg.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prec = {"fast": PREC_FAST, "fp16": PREC_FP16, "split": PREC_SPLIT, "bf16": PREC_BF16}[args.prec]
    dn_prec = PREC_BF16 if prec == PREC_BF16 else PREC_SPLIT   # z feeds the 2^9 octave: DepthNet always runs split
    L = _lib.lib()
    coarse, fine, dn = build_models(dev, prec)
    dn.precision = dn_prec
    pk_dn, pk_nerf = dn.packed(), fine.packed()
    model = pk_nerf.c_model()
    K = intrinsics()
    n_rays = H * W
    total_steps = args.warmup + args.steps

    # inputs resident in HBM before the timed region: this rank's ray slice of every step's view batch
    rays = [ops.get_rays(H, W, K, pose_for_step(i * world + rank), dev) for i in range(total_steps)]
    grid = ops.uniform_grid(DISTANCE, S, dev)
    mean = torch.empty(n_rays, 1, device=dev)
    z = torch.empty(n_rays, S, device=dev)
    raw = torch.empty(n_rays, S, 4, device=dev)
    rgb = torch.empty(n_rays, 3, device=dev)
    disp = torch.empty(n_rays, device=dev)
    acc = torch.empty(n_rays, device=dev)
    depth = torch.empty(n_rays, device=dev)
    weights = torch.empty(n_rays, S, device=dev)
    guard = torch.zeros(n_rays + 4, dtype=torch.int32, device=dev)
    tile = torch.empty(n_rays, 4, device=dev)
    gathered = torch.empty(world * n_rays, 4, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    st = torch.cuda.current_stream().cuda_stream
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = []

    def step(i, timed):
        ro, rd, vd = rays[i]
        e = [ev() for _ in range(5)] if timed else None
        if timed:
            e[0].record()
        _lib.check(L.b200nerf_depthnet_fwd(pk_dn.wpack.data_ptr(), pk_dn.aux.data_ptr(), pk_dn.n_hidden, dn_prec, ro.data_ptr(), rd.data_ptr(),
                                           n_rays, 2.0, 2.0, 6.0, mean.data_ptr(), st))
        if timed:
            e[1].record()
        _lib.check(L.b200nerf_place_samples(mean.data_ptr(), grid.data_ptr(), n_rays, S, 1, 2.0, 6.0, z.data_ptr(), st))
        if timed:
            e[2].record()
        _lib.check(L.b200nerf_nerf_query(C.byref(model), ro.data_ptr(), rd.data_ptr(), vd.data_ptr(), z.data_ptr(), None, n_rays, S,
                                         guard.data_ptr(), raw.data_ptr(), st))
        if timed:
            e[3].record()
        _lib.check(L.b200nerf_composite_fwd(raw.data_ptr(), z.data_ptr(), rd.data_ptr(), None, n_rays, S, 1, rgb.data_ptr(), disp.data_ptr(),
                                            acc.data_ptr(), depth.data_ptr(), weights.data_ptr(), None, st))
        if timed:
            e[4].record()
            marks.append(e)
        if world > 1:  # image tiles (rgb + disp) to every rank over NVLink
            tile[:, :3].copy_(rgb)
            tile[:, 3].copy_(disp)
            dist.all_gather_into_tensor(gathered, tile)
        flush.zero_()  # evict L2 between iterations (0.04 ms, inside the timed loop, stated in config)

    for i in range(args.warmup):
        step(i, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    t_start, t_end = ev(), ev()
    t_start.record()
    for i in range(args.warmup, total_steps):
        step(i, True)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    ms = torch.tensor([t_start.elapsed_time(t_end)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = world * n_rays * args.steps / (ms_total * 1e-3)

    k_ms = [statistics.mean(m[j].elapsed_time(m[j + 1]) for m in marks) for j in range(4)]  # depthnet, place, mlp, composite
    pk = peaks()
    mlp_tflops = NERF_FLOP_PER_POINT * n_rays * S / (k_ms[2] * 1e-3) / 1e12
    comp_gbs = COMPOSITE_BYTES_PER_RAY * n_rays / (k_ms[3] * 1e-3) / 1e9

    # end-to-end: the C-ABI call a host application makes, pinned host rays in, host image out
    h_ro = torch.empty(n_rays, 3).pin_memory()
    h_rd = torch.empty(n_rays, 3).pin_memory()
    h_rgb = torch.empty(n_rays, 3).pin_memory()
    h_disp = torch.empty(n_rays).pin_memory()
    h_ro.copy_(rays[0][0])
    h_rd.copy_(rays[0][1])
    ws = torch.empty(L.b200nerf_render_host_ws_bytes(n_rays, S), dtype=torch.uint8, device=dev)

    def e2e_step():
        _lib.check(L.b200nerf_render_depthnet_host(pk_dn.wpack.data_ptr(), pk_dn.aux.data_ptr(), pk_dn.n_hidden, dn_prec, C.byref(model),
                                                   h_ro.data_ptr(), h_rd.data_ptr(), n_rays, S, 1, grid.data_ptr(),
                                                   2.0, 2.0, 6.0, ws.data_ptr(), h_rgb.data_ptr(), h_disp.data_ptr(), st))

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()  # synchronises the stream itself: the host image is valid on return
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = world * n_rays * n_e2e / float(e2e_s)

    if rank == 0:
        cpu = None if args.no_cpu_baseline or world > 1 else cpu_baseline_sample()
        line = {
            "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {PREC_FAST: "fp16 operands, fp32 accumulate (+ bf16 hi/lo split re-evaluation of the guard band)",
                      PREC_FP16: "fp16", PREC_SPLIT: "bf16x2-split (bf16 hi+lo operands, fp32 accumulate)", PREC_BF16: "bf16"}[prec],
            "data": "synthetic",
            "config": {"workload": f"800x800 lego-shaped view per GPU, DepthNet + {S} uniform samples/ray, 8x256 skip@4 NeRF "
                                   "(BASELINE config #2; #3 for N>1: one 640,000-ray slice of the view batch per rank)",
                       "rays_per_step_per_gpu": n_rays, "samples_per_ray": S, "precision": args.prec, "weights": "random-init seed 42",
                       "l2": "256 MiB flush between iterations + per-step working set 840 MB > 126 MB L2",
                       "parallelism": f"ray-sharded x{world}, all_gather of rgb+disp tiles" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": 2 * n_rays * 12, "d2h_bytes_per_step": n_rays * 16,
                    "api": "b200nerf_render_depthnet_host (pinned host rays -> host rgb+disp)", "steps": n_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor",
                         "kernel": ("fast::nerf_fast_kernel<fp16> (+ exact::mlp_exact_kernel over the guard band) via b200nerf_nerf_query"
                                    if prec in (PREC_FAST, PREC_FP16) else
                                    ("exact::mlp_exact_kernel<NERF> via b200nerf_nerf_query" if prec == PREC_SPLIT
                                     else "mlp_chain_kernel<BF16,NERF> via b200nerf_nerf_query")),
                         "achieved": mlp_tflops,
                         "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": mlp_tflops / pk["tf_sustained"],
                         "frac_executed": mlp_tflops * (1_187_840 if prec == PREC_BF16 else 1_056_768) / 1_186_816 / pk["tf_sustained"],
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.prec),
                         "peak_source": pk["src"] + ", sustained bf16", "ms_per_launch": k_ms[2],
                         "algorithmic_flop_per_point": NERF_FLOP_PER_POINT,
                         # executed tensor-core work: padded K (1,187,840 FLOP/point); the fast kernel folds feature_linear into the view
                         # layer (-131,072 FLOP/point) and so does the pipelined split kernel, which issues 3 MMAs per MAC
                         "executed_mma_tflops": mlp_tflops * ((3 * 1_056_768) if prec == PREC_SPLIT else
                                                              (1_187_840 if prec == PREC_BF16 else 1_056_768)) / 1_186_816,
                         "guard_band_points_last_step": int(guard[0]) if prec == PREC_FAST else None,
                         "note": ("achieved counts the reference network's 1,186,816 FLOP/point; the fast kernel folds the activation-free "
                                  "feature_linear into views_linears.0 when packing (1,056,768 FLOP/point executed), so frac can "
                                  "exceed 1 -- frac_executed = executed_mma_tflops / peak is the tensor-pipe load")
                         if prec != PREC_BF16 else None},
            "roofline_composite": {"bound": "hbm", "kernel": "comp::composite_tma_kernel<16> (TMA-staged, persistent)", "achieved": comp_gbs, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": comp_gbs / pk["hbm"], "ms_per_launch": k_ms[3], "bytes_per_ray": COMPOSITE_BYTES_PER_RAY},
            "kernel_ms": {"depthnet": k_ms[0], "place": k_ms[1], "nerf_mlp": k_ms[2], "composite": k_ms[3]},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
