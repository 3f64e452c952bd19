#!/usr/bin/env python
"""bench.py — Mrays/s and frames/s of the raytracing_rb hot path on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

A "step" is one frame of the workload BASELINE.json's metric is quoted on (configs[1]: 1920x1080,
ground plane + 16 spheres, hard shadows, 1 spp, no recursion).  Metric: Mrays/s where rays =
traced rays (work-stack items passing the cut at ray_tracer.rb:52) + shadow queries (lit_area calls
from local_lights, world.rb:75), SURVEY.md 8d.

A step is a batch of --frames-per-step frames PER GPU (a camera dolly).  For N > 1 (torchrun, one
rank per GPU) the batch's frames are the dealing unit: config 2 is a 0.12 ms frame, far too small to
cut into tiles per GPU, so rank r renders frames [r*B, (r+1)*B) whole and they are gathered into rank
0's frame slots through a CUDA-IPC peer mapping over NVLink - by direct peer stores from the trace
kernel (--gather store) or by the copy engine behind the kernel (--gather copy).  No collective on the
data path; per-GPU work is fixed as N grows (weak scaling).  Tile partitioning of ONE heavy frame
(config 5) is `--tile-split` (strong scaling of a single frame).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# SURVEY.md 8d algorithmic FLOP constants (add/mul = 1, FMA = 2, div/sqrt = 1, transcendental = 1)
FLOP = dict(ray=7, sphere_reject=20, sphere_accept=44, plane_reject=14, plane_accept=29, hit=82 + 2,
            lambert_per_light=24, lambert_base=15, uv=34, shadow=10, cover_sphere_none=24, cover_sphere_full=56,
            cover_sphere_pen=83 + 4, cover_plane_reject=14, cover_plane_accept=37, primary=60 + 2)


def algorithmic_flops(s):
    return (s["rays"] * FLOP["ray"]
            + (s["sphere_tests"] - s["sphere_accepts"]) * FLOP["sphere_reject"] + s["sphere_accepts"] * FLOP["sphere_accept"]
            + (s["plane_tests"] - s["plane_accepts"]) * FLOP["plane_reject"] + s["plane_accepts"] * FLOP["plane_accept"]
            + s["hits"] * FLOP["hit"] + s["lit_lights"] * FLOP["lambert_per_light"] + s["local_shaded"] * FLOP["lambert_base"]
            + s["texel_fetches"] * FLOP["uv"] + s["shadow_queries"] * FLOP["shadow"]
            + (s["cover_sphere"] - s["cover_sphere_full"] - s["cover_sphere_penumbra"]) * FLOP["cover_sphere_none"]
            + s["cover_sphere_full"] * FLOP["cover_sphere_full"] + s["cover_sphere_penumbra"] * FLOP["cover_sphere_pen"]
            + (s["cover_plane"] - s["cover_plane_accepts"]) * FLOP["cover_plane_reject"]
            + s["cover_plane_accepts"] * FLOP["cover_plane_accept"] + s["samples"] * FLOP["primary"])


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons DURING the timed region through NVML (nvidia-ml-py), every
    ~2 ms; falls back to polling nvidia-smi when NVML is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            nv = self.nv
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
        else:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown")
            out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                           "--format=csv,noheader,nounits"], timeout=5).decode()
            r = [x.strip() for x in out.strip().split(",")]
            self.sm.append(float(r[0]))
            self.max_mhz = float(r[1])
            for (name, _), v in zip(self.REASONS, r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._halt.wait(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def workload(config_id, small):
    from raytracing_rb_b200 import Camera, World, scenes
    kw = {}
    if small:
        kw = dict(width=480, height=270)
    wdoc, cdoc = scenes.build(config_id, **kw)
    return World(wdoc), cdoc, scenes.NAMES[config_id]


def run_reference(args, rank, world_size):
    """The reference's own CPU implementation of the path: no Ruby interpreter exists in this image
    (probed below), so this is the FP64 C++ restatement (oracle/, `kind: port`) run as column strips
    over every host core exactly like render_fork (camera.rb:53-65)."""
    if rank != 0:
        return
    from raytracing_rb_b200 import Camera, make_opts
    from oracle import oracle
    world, cdoc, name = workload(args.config, args.small)
    cd = Camera(world, cdoc).camera_desc()
    cores = os.cpu_count() or 1
    sc = oracle.OracleScene(world.to_scene_desc())
    # bounded sample: a centred window sized so one step stays within seconds on the host cores
    frac = args.cpu_fraction
    W, H = cd.width, cd.height
    ww = max(cores, int(W * frac))
    win = ((W - ww) // 2, 0, (W - ww) // 2 + ww, H)
    opts = make_opts(seed=1, window=win)
    for _ in range(args.warmup):
        sc.render(cd, opts, threads=cores, want_rgb=False, want_hit=False)
    t0 = time.perf_counter()
    rays = 0
    for _ in range(args.steps):
        f = sc.render(cd, opts, threads=cores, want_rgb=False, want_hit=False)
        rays += f.stats["rays"] + f.stats["shadow_queries"]
    dt = time.perf_counter() - t0
    value = rays / dt / 1e6
    sample = "window x in [%d,%d) of %dx%d (%.0f%% of the frame) per step, %d column strips" % (
        win[0], win[2], W, H, 100.0 * ww / W, cores)
    ruby = subprocess.call("command -v ruby", shell=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name + "; one step = a batch of %d frames per GPU (camera dolly)" % args.frames_per_step,
                   "frames_per_step": args.frames_per_step, "host": "CPU only (one frame of the batch per step)",
                   "ruby_present": ruby},
        "frames_per_s_equiv": args.steps / dt * (ww / W),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def batch_cameras(world, cdoc, n, first=0):
    """Frames [first, first+n) of the fly-by: the workload's camera dollying sideways, 0.02 per frame
    (wrapping every 64 frames so every rank's share of a large batch shows the same scene content)."""
    from raytracing_rb_b200 import Camera
    cams = []
    for f in range(first, first + n):
        c = Camera(world, cdoc).camera_desc()
        c.position[1] = c.position[1] + 0.02 * (f % 64)
        cams.append(c)
    return cams


def bind_to_gpu_numa_node(index):
    """Pins this process to the CPU cores NVML reports as local to GPU `index`, BEFORE any pinned host buffer
    is allocated: on a two-socket 8-GPU box a rank whose frame buffers sit on the other socket sends every
    device-to-host frame across the inter-socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:
        return False


def run_ours(args, rank, local_rank, world_size):
    import torch
    from raytracing_rb_b200 import (Renderer, _abi, deal_frames, ipc_open, make_opts, measure_fma_peak, PREC_FAST64, PREC_STRICT)
    from raytracing_rb_b200._lib import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:  # (N = 1 keeps every host core: the CPU baseline leg runs in this process)
        bind_to_gpu_numa_node(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL announces its version on STDOUT when the first communicator is built; stdout must carry
        # exactly one JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    precision = PREC_STRICT if args.precision == "strict" else PREC_FAST64
    fmt = _abi.FMT_RGB8 if args.pixel_format == "rgb8" else _abi.FMT_RGBA8
    bpp = 3 if fmt == _abi.FMT_RGB8 else 4
    world, cdoc, name = workload(args.config, args.small)
    if args.spp:
        cdoc = dict(cdoc, pre_sample_times=args.spp, max_sample_times=args.spp)
    B = max(1, args.frames_per_step)
    tile_split = bool(args.tile_split) and world_size > 1
    # frames this rank renders per step, and the slot each lands in inside rank 0's framebuffer
    if tile_split:
        cams = batch_cameras(world, cdoc, B)            # every rank: its tiles of the same B frames
        slots = list(range(B))
        n_slots = B
    else:
        dealt = deal_frames(rank, world_size, B)        # rank r: frames [r*B, (r+1)*B), whole
        cams = batch_cameras(world, cdoc, B, dealt[0][0] if dealt else 0)
        slots = [slot for _, slot in dealt]
        n_slots = B * world_size
    W, H = cams[0].width, cams[0].height
    slot_bytes = W * H * 4           # slots are sized for RGBA8 whatever the format
    frame_bytes = W * H * bpp
    r = Renderer(world.to_scene_desc(), local_rank)

    # ---- where the pixels go: frame slots in rank 0's framebuffer (a peer mapping for the others) ----
    if rank == 0:
        base_ptr = r.framebuffer_ptr(W, H * n_slots)
    if world_size > 1:
        hbuf = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            hbuf.copy_(torch.frombuffer(bytearray(r.framebuffer_ipc_export(W, H * n_slots)), dtype=torch.uint8))
        dist.broadcast(hbuf, 0)
        if rank != 0:
            base_ptr = ipc_open(local_rank, bytes(hbuf.cpu().numpy().tobytes()))
    use_copy = (not tile_split) and world_size > 1 and rank != 0 and args.gather == "copy"
    local_ptr = r.framebuffer_ptr(W, H * B) if use_copy else None
    stream = torch.cuda.Stream()  # a real (non-NULL) stream: NULL means "the renderer's own stream" in the C ABI
    torch.cuda.set_stream(stream)

    def opts(f, detail=False, prec=None):
        out = (local_ptr + f * slot_bytes) if use_copy else (base_ptr + slots[f] * slot_bytes)
        return make_opts(seed=1, precision=precision if prec is None else prec,
                         tile_rank=rank if tile_split else 0, tile_world=world_size if tile_split else 1,
                         count_detail=detail, stream=stream.cuda_stream, rgba_device_out=out,
                         skip_outputs=_abi.SKIP_RGB | _abi.SKIP_HIT, pixel_format=fmt)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- untimed: this rank's work per batch (counters) and the FMA issue peaks ----
    # The algorithmic (brute-force) operation counts come from the STRICT kernel's detailed counters:
    # FAST64 produces the same frames with fewer executed tests, which must not shrink the numerator.
    flops_batch, rays_batch = 0, 0
    n_count = B if not args.count_one else 1
    for f in range(n_count):
        st, _ = r.render_device(cams[f], opts(f, True, PREC_STRICT if not args.count_fast else None))
        flops_batch += algorithmic_flops(st)
        rays_batch += st["rays"] + st["shadow_queries"]
    if n_count != B:  # heavy frames (config 5): one frame counted, the batch is B copies of its cost
        flops_batch, rays_batch = flops_batch * B, rays_batch * B
    peak64 = measure_fma_peak(local_rank, True) if rank == 0 else 0.0
    peak32 = measure_fma_peak(local_rank, False) if rank == 0 else 0.0

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    step_opts = [opts(f) for f in range(B)]

    def step():
        for f in range(B):
            r.render_device(cams[f], step_opts[f], want_stats=False)
            if use_copy:  # copy engine carries the finished frame to rank 0 while the next one renders
                r.peer_push(local_ptr + f * slot_bytes, base_ptr + slots[f] * slot_bytes, frame_bytes, stream.cuda_stream)
        if use_copy:
            r.peer_push_join(stream.cuda_stream)  # the step ends when its last frame has landed
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed: EXACTLY K steps, CUDA events on the launch stream, L2 flushed between steps ----
    launches0 = lib().rtrb_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    launches = lib().rtrb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(rays_batch), float(flops_batch), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)  # a step is done when its slowest rank is
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms = float(step_ms.sum().item())
    rays_total, flops_total, launches_total = (float(x) for x in tot.tolist())
    frames_total = n_slots  # frames finished per step by the whole job
    value = rays_total * args.steps / (total_ms * 1e-3) / 1e6

    # ---- dominant kernel alone (trace over the pre samples): library-side CUDA events, cold L2 ----
    tr = []
    for k in range(min(args.steps, 10)):
        for f in range(B if not args.count_one else 1):
            flush.zero_()
            st, _ = r.render_device(cams[f], opts(f))
            tr.append(st["trace_ms"])
    trace_ms = float(np.mean(tr))

    # ---- e2e: the public frame call with HOST buffers (pinned), copies inside the timed region ----
    # Every rank delivers the frames it rendered to its own pinned host buffers over its own PCIe link
    # through the pipelined frame API (frame i+1 renders while frame i crosses PCIe).
    cam_bytes = C.sizeof(_abi.CameraDesc) + C.sizeof(_abi.RenderOpts)
    e2e_opts = make_opts(seed=1, precision=precision, pixel_format=fmt,
                         tile_rank=rank if tile_split else 0, tile_world=world_size if tile_split else 1)
    DEPTH = 3  # frames in flight (the library allows 4)
    bufs = [torch.empty((H, W, bpp), dtype=torch.uint8).pin_memory().numpy() for _ in range(DEPTH)]

    def e2e_run(n_steps):
        pending, i = [], 0
        for _ in range(n_steps):
            for f in range(B):
                if len(pending) == DEPTH:
                    r.wait(pending.pop(0))
                pending.append(r.submit(cams[f], bufs[i % DEPTH], e2e_opts))
                i += 1
        for t in pending:
            r.wait(t)
    e2e_run(1)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_dt.item())
    e2e_value = rays_total * args.steps / e2e_s / 1e6

    if rank == 0:
        if tile_split:
            how = "32x32 px super-tiles of every frame dealt round-robin over ranks, peer stores into rank 0's framebuffer"
        elif world_size > 1:
            how = ("whole frames dealt to ranks (rank r renders frames [r*B, (r+1)*B)), gathered into rank 0's frame slots by " +
                   ("the copy engine over NVLink (rtrb_peer_push)" if args.gather == "copy" else "direct peer stores from the trace kernel"))
        else:
            how = "single GPU"
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if tile_split else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name + "; one step = a batch of %d frames per GPU (camera dolly)" % B,
                       "frames_per_step": frames_total, "frames_per_step_per_gpu": B, "precision_mode": args.precision,
                       "pixel_format": args.pixel_format, "rays_per_step": rays_total,
                       "l2": "flushed between timed steps (256 MiB write)",
                       "multi_gpu": how, "rng": "philox4x32-10 counter, seed 1"},
            "frames_per_s": args.steps * frames_total / (total_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": cam_bytes * frames_total,
                    "d2h_bytes_per_step": frame_bytes * frames_total if not tile_split else frame_bytes * B * world_size,
                    "frames_per_s": args.steps * frames_total / e2e_s,
                    "api": "rtrb_submit/rtrb_wait per rank (%d frames in flight, pinned host buffers, %s)" % (DEPTH, args.pixel_format)},
            "gpu_launches": int(launches_total),
            "clocks": clocks,
            "wall_s_timed_region": wall,
        }
        flops_frame = flops_batch / B
        achieved = flops_frame / (trace_ms * 1e-3) / 1e12
        line["roofline"] = {
            "bound": "fp64", "kernel": "trace_pre_fast_kernel", "achieved": achieved, "peak": peak64, "unit": "TFLOP/s",
            "frac": achieved / peak64 if peak64 else None, "traffic": ncu_traffic_bytes() if args.config == 2 and not args.small else None,
            "peak_source": "measured in this job: dependent-free DFMA microbenchmark (rtrb_measure_fma_peak)",
            "peak_fp32": peak32, "algorithmic_flops_per_launch": flops_frame, "kernel_ms": trace_ms,
            "hbm": {"algorithmic_bytes_per_launch": frame_bytes, "peak_gbs": measured_hbm_gbs(),
                    "achieved_gbs": frame_bytes / (trace_ms * 1e-3) / 1e9},
            "note": "FP-issue bound path (SURVEY.md 8d): HBM traffic is the framebuffer write only; rank 0's kernel and counters",
        }
        if world_size == 1:
            line["cpu_baseline"] = cpu_baseline(args, world, cams[0], name)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the trace kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/r1_traffic.json); None when absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_hbm_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"]
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


def cpu_baseline(args, world, cd, name):
    """The oracle (kind 'port') on the GPU box's host cores, bounded sample of the same workload."""
    from raytracing_rb_b200 import make_opts
    from oracle import oracle
    cores = os.cpu_count() or 1
    sc = oracle.OracleScene(world.to_scene_desc())
    W, H = cd.width, cd.height
    ww = max(cores, int(W * args.cpu_fraction))
    win = ((W - ww) // 2, 0, (W - ww) // 2 + ww, H)
    o = make_opts(seed=1, window=win)
    sc.render(cd, o, threads=cores, want_rgb=False, want_hit=False)
    t0 = time.perf_counter()
    n, rays = 0, 0
    while True:
        f = sc.render(cd, o, threads=cores, want_rgb=False, want_hit=False)
        rays += f.stats["rays"] + f.stats["shadow_queries"]
        n += 1
        if time.perf_counter() - t0 > 10.0 or n >= 50:
            break
    dt = time.perf_counter() - t0
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": "%d passes over window x in [%d,%d) of %dx%d, %d column strips (render_fork shape)" % (
                n, win[0], win[2], W, H, cores)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames rendered per timed step")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--precision", default="fast64", choices=["fast64", "strict"])
    ap.add_argument("--small", action="store_true", help="480x270 variant for quick checks (not a bench value)")
    ap.add_argument("--pixel-format", default="rgb8", choices=["rgb8", "rgba8"],
                    help="8-bit frame layout delivered (rgb8: alpha is the constant 255 and stays off the wire)")
    ap.add_argument("--gather", default="copy", choices=["store", "copy"],
                    help="N > 1: how finished frames reach rank 0's slots (peer stores from the kernel / copy engine)")
    ap.add_argument("--tile-split", action="store_true",
                    help="N > 1: cut every frame into super-tiles across the ranks (strong scaling of heavy frames)")
    ap.add_argument("--spp", type=int, default=0, help="override pre = max sample count (heavy configs at reduced cost)")
    ap.add_argument("--count-one", action="store_true", help="count rays/FLOPs on one frame of the batch only")
    ap.add_argument("--count-fast", action="store_true", help="take the counters from the FAST64 kernel (heavy configs)")
    ap.add_argument("--cpu-fraction", type=float, default=1.0, help="fraction of the frame width the CPU legs render")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
    else:
        run_ours(args, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
