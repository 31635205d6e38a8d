#!/usr/bin/env python
"""bench.py -- headline benchmark of the 3DGS rasterizer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE configs[2] "C3" -- 1M-Gaussian room, 1920x1080, SH degree 3,
seeded synthetic scene (robosimgs_b200/scenes.py:room_scene).  A *step* is one forward render of one
camera of a camera-sharded sweep (global frame f = step*N + rank; cameras are the C3 camera with a
deterministic centimetre-scale jitter so every frame is distinct work of the same size).

Printed JSON line (rank 0): `value` = rendered Mpixels/s over all ranks, inputs resident in HBM,
device-timed with CUDA events, max over ranks; `train` = fwd + MSE loss + bwd iterations/s on the
C3 camera; `e2e` = the same render metric through the public operator with the per-frame camera
coming from pinned host memory and the finished frame copied back to pinned host memory inside the
timed region; `roofline` = dominant kernel against the measured HBM peak; `cpu_baseline` = the CPU
oracle (oracle/, test infrastructure) timed on this box's host cores.

`--impl reference` times the CPU restatement of the reference algorithm (oracle/gs_oracle.c, OpenMP,
all host threads) on the same frames -- the reference repo has no rasterizer of its own to run
(SURVEY.md section 0), so this is kind "port".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rendered Mpixels/s (fwd) + train iters/s (fwd+bwd), 1M Gaussians @1080p"
W_IMG, H_IMG, P_SCENE, SH_DEG = 1920, 1080, 1_000_000, 3
WORKLOAD = ("C3: 1M-Gaussian room background, 1920x1080, SH deg 3 (BASELINE configs[2]); "
            "camera-sharded sweep of jittered C3 cameras")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        self.thr = threading.Thread(target=pump, daemon=True)
        self.thr.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def jittered_cameras(n, first=0):
    """Frames of the sweep: the C3 camera with a deterministic ~1 cm eye jitter per global frame."""
    from robosimgs_b200.cameras import camera_look_at
    cams = []
    for f in range(first, first + n):
        j = [0.01 * math.sin(0.7 * f), 0.01 * math.cos(1.3 * f), 0.005 * math.sin(2.1 * f)]
        cams.append(camera_look_at((-1.8 + j[0], -1.2 + j[1], 0.1 + j[2]), (2.0, 1.0, -0.2), (0, 0, 1), 70.0,
                                   W_IMG, H_IMG))
    return cams


def algorithmic_bytes(P, P_vis, D, M):
    """SURVEY.md 8(d) / BASELINE.md section 2, per frame, fp32."""
    b_in = 12 + 12 + 16 + 4 + 12 * M
    px = W_IMG * H_IMG
    per_stage = {
        "project": P * b_in + P_vis * 48,
        "scan": P * 8,
        "emit_pairs": D * 12,
        "pair_sort": D * 12 * 2,
        "tile_ranges": D * 8,
        "render": D * 40 + px * (12 + 8),
        "render_bwd": D * 40 + px * (12 + 8) + P_vis * 40,
        "project_bwd": P_vis * 40 + P * b_in + P * (b_in + 12),
    }
    fwd = P * b_in + P_vis * 48 + D * 24 + D * 40 + px * 12 + px * 8
    bwd = D * 40 + px * 20 + P_vis * 80 + P * b_in + P * (b_in + 12)
    return per_stage, fwd, bwd


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: oracle/gs_oracle.c (OpenMP, all host threads) on the same frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import __graft_entry__ as ge
    from oracle import gs_oracle
    from robosimgs_b200.scenes import room_scene, settings_from_camera
    gs_oracle.build()
    cores = os.cpu_count() or 1
    gs_oracle.set_num_threads(cores)
    sc, _ = room_scene(P_SCENE, 3, SH_DEG, W_IMG, H_IMG)
    cams = jittered_cameras(args.steps + args.warmup)
    kw = dict(shs=sc.shs.numpy(), scales=sc.scales.numpy(), rotations=sc.rotations.numpy(), dtype=np.float32)
    means, opac = sc.means3D.numpy(), sc.opacities.numpy()
    for c in cams[:args.warmup]:
        gs_oracle.forward(settings_from_camera(c, SH_DEG), means, opac, **kw)
    t0 = time.perf_counter()
    for c in cams[args.warmup:]:
        st = gs_oracle.forward(settings_from_camera(c, SH_DEG), means, opac, **kw)
    dt = time.perf_counter() - t0
    val = args.steps * W_IMG * H_IMG / dt / 1e6
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "P": P_SCENE, "D_reference_rects": int(st.num_rendered),
                   "P_vis": int((st.radii > 0).sum())},
        "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": gs_oracle.num_threads(), "kind": "port",
                         "threads_per_stage": {"preprocess": gs_oracle.num_threads(), "bin_and_sort": 1,
                                               "render": gs_oracle.num_threads()},
                         "sample": f"{args.steps} full 1920x1080 frames (preprocess + bin/sort + render), "
                                   "oracle/gs_oracle.c fp32, OpenMP"},
        "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from robosimgs_b200 import GaussianRasterizer, _cabi, export_rgb8
    from robosimgs_b200.rasterizer import GaussianRasterizationSettings
    from robosimgs_b200.losses import mse_loss as fused_mse_loss
    from robosimgs_b200.scenes import mse_loss, room_scene, room_target, settings_from_camera

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from robosimgs_b200.sweep import bind_rank_to_cores
    cores = bind_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))     # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()

    # ---- scene: generated on rank 0, replicated to every GPU with one NCCL broadcast per tensor ----
    names = ("means3D", "shs", "opacities", "scales", "rotations")
    if rank == 0:
        sc, _ = room_scene(P_SCENE, 3, SH_DEG, W_IMG, H_IMG)
        tens = {k: getattr(sc, k).to(dev) for k in names}
    else:
        M = (SH_DEG + 1) ** 2
        shapes = {"means3D": (P_SCENE, 3), "shs": (P_SCENE, M, 3), "opacities": (P_SCENE, 1),
                  "scales": (P_SCENE, 3), "rotations": (P_SCENE, 4)}
        tens = {k: torch.empty(shapes[k], dtype=torch.float32, device=dev) for k in names}
    if world > 1:
        for k in names:
            dist.broadcast(tens[k], src=0)
    M = tens["shs"].shape[1]
    means2D = torch.zeros_like(tens["means3D"])

    K, Wm, R = args.steps, args.warmup, max(1, args.repeats)
    nframes = K + Wm
    # this rank's frames of the sweep: global frame f = step*world + rank
    cams = [jittered_cameras(1, first=s * world + rank)[0] for s in range(nframes)]
    bg = torch.zeros(3, device=dev)
    settings_dev = [settings_from_camera(c, SH_DEG, device=dev) for c in cams]

    def render(rs):
        return GaussianRasterizer(rs)(tens["means3D"], means2D, tens["opacities"], shs=tens["shs"],
                                      scales=tens["scales"], rotations=tens["rotations"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, fs=None):
        """K steps bracketed by barrier + synchronize; CUDA events on the current stream; max over ranks.
        With fs (FrameStreams) the frame streams fork after the start event and join before the end
        event, so the events on the current stream bracket all the work of the K steps."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        if fs is not None:
            fs.fork()
        for s in range(steps):
            fn(s)
        if fs is not None:
            fs.join()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (also yields P_vis / D of the workload) ----
    with torch.no_grad():
        for s in range(max(Wm, 3) + 5):
            color, radii = render(settings_dev[s % nframes])
    torch.cuda.synchronize()
    m3 = tens["means3D"].detach().requires_grad_(True)
    color, radii = GaussianRasterizer(settings_dev[Wm % nframes])(
        m3, means2D, tens["opacities"], shs=tens["shs"], scales=tens["scales"], rotations=tens["rotations"])
    D = int(color.grad_fn.num_rendered)
    P_vis = int((radii > 0).sum().item())
    del color, m3

    # ---- headline: forward render sweep, inputs resident.  Frames of the sweep are independent, so
    # consecutive frames alternate over `--streams` CUDA streams (robosimgs_b200.sweep.FrameStreams):
    # one frame's latency-bound binning stage overlaps the other's compositing.  The single-stream
    # pass that follows gives the per-frame latency and the per-stage times of the roofline. ----
    from robosimgs_b200.sweep import FrameStreams
    sampler = ClockSampler(local)
    sampler.start()
    fs = FrameStreams(dev, args.streams)

    # cameras resident in HBM (35 floats each); frames stay on the device.  Each frame slot of the
    # renderer replays ONE captured CUDA graph of the forward (pair-count check deferred, validated when
    # the frame is collected), so the host cost per frame is a 140-byte device copy + one graph launch.
    from robosimgs_b200.sweep import SceneRenderer
    cam_blocks = torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1)])
                              for c in cams]).to(dev)
    sweep_r = SceneRenderer(tens, SH_DEG, bg, H_IMG, W_IMG, streams=max(1, args.streams), graphs=not args.no_graphs,
                            host_frames=False)
    fs = sweep_r.fs
    pend = []

    def piped_render(s):
        while len(pend) >= sweep_r.in_flight_limit():
            sweep_r.collect(pend.pop(0))
        pend.append(sweep_r.submit(cams[Wm + s], cam_block=cam_blocks[Wm + s]))
        if s == K - 1:                    # every frame of the timed region is collected (validated) inside it
            while pend:
                sweep_r.collect(pend.pop(0))

    redone = [0]
    with torch.no_grad():
        if args.streams > 1:
            _cabi.launch_count(reset=True)
            sweep_r.collect(sweep_r.submit(cams[Wm], cam_block=cam_blocks[Wm]))     # exact first frame
            launches_per_frame = _cabi.launch_count(reset=True)
            for s in range(2 * sweep_r.in_flight_limit()):      # graph capture per slot
                piped_render(s % max(K - 1, 1))
            while pend:
                sweep_r.collect(pend.pop(0))
            torch.cuda.synchronize()
            sweep_r.redone = 0
            fwd_runs = [timed(piped_render, K, fs) for _ in range(R)]      # R timed runs of K frames each
            fwd_ms = statistics.median(fwd_runs)
            redone[0] = sweep_r.redone
            launches = launches_per_frame * K      # graph replays: the launches of one frame, K times
    _cabi.launch_count(reset=True)
    with torch.no_grad():
        serial_runs = [timed(lambda s: render(settings_dev[Wm + s]), K) for _ in range(R)]
        serial_ms = statistics.median(serial_runs)
        serial_launches = _cabi.launch_count(reset=True) // R
        # per-stage times: one more pass with the library's stage timers on (their ~12 event records per frame cost
        # ~0.04 ms of a single-stream frame, so this pass is not the one that is reported as the frame time)
        _cabi.profile_enable(True)
        _cabi.profile_read(reset=True)
        timed(lambda s: render(settings_dev[Wm + s]), K)
    if args.streams <= 1:
        fwd_ms, fwd_runs, launches = serial_ms, serial_runs, serial_launches
    _cabi.launch_count(reset=True)
    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    stages_fwd = _cabi.profile_read(reset=True)
    _cabi.profile_enable(False)

    # ---- train step: fwd + MSE loss + bwd on the C3 camera ----
    target = room_target(W_IMG, H_IMG).to(dev)
    leaves = {k: tens[k].detach().clone().requires_grad_(True) for k in names}
    m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
    rs_train = settings_from_camera(jittered_cameras(1, first=0)[0], SH_DEG, device=dev)

    from robosimgs_b200.train import backward_or_retry

    def make_train_step(loss_fn, defer):
        # defer: the rasterizer's opt-in training mode -- forward returns without waiting for the pair count, backward
        # launches the adjoint and validates it afterwards (every step, inside the timed region); an overflow would
        # raise PairCapacityExceeded and backward_or_retry would render the step again
        rast = GaussianRasterizer(rs_train)
        rast.defer_pair_check = bool(defer)

        def loss_of_frame():
            col, _ = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                          rotations=leaves["rotations"])
            return loss_fn(col, target)

        def train_step(_s):
            for t in list(leaves.values()) + [m2]:
                t.grad = None
            backward_or_retry(loss_of_frame)
        return train_step

    train_step = make_train_step(fused_mse_loss, True)          # loss fused in libb200gs (2 launches)
    train_step_blocking = make_train_step(fused_mse_loss, False)  # the operator's default: the host learns D inside forward
    train_step_eager = make_train_step(mse_loss, True)          # same loss as eager PyTorch ops (~8 launches)
    for s in range(max(Wm, 3) + 5):
        train_step(s)
        train_step_blocking(s)
        train_step_eager(s)
    _cabi.launch_count(reset=True)
    _cabi.profile_enable(True)
    _cabi.profile_read(reset=True)
    train_ms = timed(train_step, K)
    stages_train = _cabi.profile_read(reset=True)
    _cabi.profile_enable(False)
    launches_train = _cabi.launch_count(reset=True)
    train_runs = [train_ms] + [timed(train_step, K) for _ in range(R - 1)]
    train_ms = statistics.median(train_runs)
    train_eager_ms = timed(train_step_eager, K)
    train_blocking_ms = timed(train_step_blocking, K)
    clocks = sampler.stop()
    del leaves, m2

    # ---- e2e: HOST buffers on both sides through the public per-frame renderer (sweep.SceneRenderer):
    # camera block from pinned host memory in, finished 8-bit frame in pinned host memory out, every
    # frame collected (and its pair-count ticket validated) by the consumer inside the timed region ----
    from robosimgs_b200.sweep import host_frames_in_flight
    NS = args.e2e_streams if args.e2e_streams > 0 else host_frames_in_flight(int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    renderer = SceneRenderer(tens, SH_DEG, bg, H_IMG, W_IMG, streams=NS, graphs=not args.no_graphs)
    h2d_bytes = 35 * 4
    d2h_bytes = H_IMG * W_IMG * 3
    handles = []
    last_frame = [None]

    def e2e_step(s):
        while len(handles) >= renderer.in_flight_limit():
            last_frame[0] = renderer.collect(handles.pop(0))
        handles.append(renderer.submit(cams[(Wm + s) % nframes]))
        if s == K - 1:
            while handles:                      # the last frames are collected inside the timed region
                last_frame[0] = renderer.collect(handles.pop(0))

    for s in range(max(Wm, 3) + 2 * renderer.in_flight_limit()):      # warm-up: exact frame, then graph capture per slot
        e2e_step(s)
    while handles:
        renderer.collect(handles.pop(0))
    torch.cuda.synchronize()
    renderer.redone = 0
    e2e_runs = [timed(e2e_step, K, renderer.fs) for _ in range(R)]
    e2e_ms = statistics.median(e2e_runs)
    checksum = float(last_frame[0].double().mean())
    e2e_redone = [renderer.redone]

    # ---- host ceiling of the e2e path: every rank copies finished frames device -> pinned host memory back to back,
    # all ranks at once, nothing else running (what the box can absorb; the e2e sweep cannot beat it) ----
    d2h_src = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    d2h_dst = [torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    def d2h_step(s_):
        d2h_dst[s_ % 4].copy_(d2h_src, non_blocking=True)
    timed(d2h_step, 50)
    d2h_ms = timed(d2h_step, 400)
    d2h_gbs = world * 400 * d2h_bytes / (d2h_ms * 1e-3) / 1e9
    del d2h_src, d2h_dst

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    px = W_IMG * H_IMG
    value = world * K * px / (fwd_ms * 1e-3) / 1e6
    e2e_val = world * K * px / (e2e_ms * 1e-3) / 1e6
    per_stage_bytes, bytes_fwd, bytes_bwd = algorithmic_bytes(P_SCENE, P_vis, D, M)
    peak, peak_src = measured_peaks()
    st_ms = {k: (v[0] / max(v[1], 1)) for k, v in stages_fwd.items() if v[1] > 0}
    st_ms_train = {k: (v[0] / max(v[1], 1)) for k, v in stages_train.items() if v[1] > 0}
    dom = max(st_ms, key=st_ms.get)
    dom_gbs = per_stage_bytes[dom] / (st_ms[dom] * 1e-3) / 1e9
    stage_gbs = {k: round(per_stage_bytes[k] / (v * 1e-3) / 1e9, 1) for k, v in {**st_ms, **st_ms_train}.items()}
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) comes from the committed `ncu --set full`
    # capture of this same workload (profiles/dominant_kernel_traffic.json) -- a number taken under a profiler
    # cannot be measured inside a timed run; `traffic_source` says so.
    traffic_tab, traffic_src = {}, None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic_tab = json.load(open(tp))
            traffic_src = "static: " + str(traffic_tab.get("_source", "profiles/dominant_kernel_traffic.json"))
        except Exception:
            traffic_tab = {}
    traffic = traffic_tab.get(dom)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    kernels = []
    for k, v in {**st_ms, **st_ms_train}.items():
        row = {"stage": k, "ms": round(v, 5), "algorithmic_bytes": per_stage_bytes[k],
               "algorithmic_GBps": round(per_stage_bytes[k] / (v * 1e-3) / 1e9, 1),
               "algorithmic_frac": round(per_stage_bytes[k] / (v * 1e-3) / 1e9 / peak, 4)}
        if isinstance(traffic_tab.get(k), (int, float)):
            row["dram_bytes"] = traffic_tab[k]
            row["dram_GBps"] = round(traffic_tab[k] / (v * 1e-3) / 1e9, 1)
            row["dram_frac"] = round(traffic_tab[k] / (v * 1e-3) / 1e9 / peak, 4)
        wi = (traffic_tab.get("_warp_instructions") or {}).get(k)
        if isinstance(wi, (int, float)) and sm_count and clocks.get("sm_mhz"):
            # issue-slot view of the same launch (static instruction count from the committed ncu capture, like the
            # DRAM bytes): warp instructions / (SMs x 4 schedulers x SM clock x event-timed duration)
            row["warp_instructions"] = int(wi)
            row["issue_frac"] = round(wi / (sm_count * 4 * clocks["sm_mhz"] * 1e6 * v * 1e-3), 4)
        kernels.append(row)
    out = {
        "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": fwd_ms / K, "repeats": R, "ms_per_step_runs": [round(t / K, 5) for t in fwd_runs],
        "timed_frames": K * R,
        "frame_latency_ms": serial_ms / K,
        "single_stream": {"ms_per_frame": serial_ms / K, "mpixels_per_s": px / (serial_ms / K * 1e-3) / 1e6,
                          "ms_per_frame_runs": [round(t / K, 5) for t in serial_runs],
                          "what": "SURVEY 8(d) definition: W*H / t_fwd with every frame issued on ONE stream through "
                                  "GaussianRasterizer.forward (no graphs, no frames in flight, stage timers off); `value` "
                                  "is the sweep rate"},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "P": P_SCENE, "P_vis": P_vis, "D_pairs": D, "sh_degree": SH_DEG,
                   "image": [W_IMG, H_IMG], "tile": 16, "parallelism": f"camera-sharded x{world}, scene replicated", "binning": "depth-sliced buckets (bucket.cu), no library scan/sort",
                   "streams": f"{args.streams} CUDA streams per GPU, consecutive frames alternate (independent frames "
                              "overlap), one captured CUDA graph per frame slot, pair-count check deferred (every frame "
                              "validated inside the timed region), cameras and frames resident in HBM; "
                              "frame_latency_ms and the roofline stage times are from a single-stream pass",
                   "frames_rendered_twice": redone[0],
                   "l2": "inputs larger than L2: 236 MB of parameters + %.0f MB of pair/slab buffers stream per frame "
                         "(126 MB L2), no explicit flush" % (D * 64 / 1e6),
                   "frame_checksum": checksum},
        "train": {"iters_per_s": world * K / (train_ms * 1e-3), "ms_per_iter": train_ms / K,
                  "what": "fwd + mean((img-target)^2) + bwd, C3 camera, per-GPU replicas (no gradient all-reduce); "
                          "loss fused in libb200gs (robosimgs_b200.losses.mse_loss); median of `repeats` timed runs of K steps; "
                          "GaussianRasterizer.defer_pair_check on: no host wait inside the step, backward validates the "
                          "pair count of every step inside the timed region (train.backward_or_retry)",
                  "ms_per_iter_runs": [round(t / K, 5) for t in train_runs],
                  "ms_per_iter_blocking": round(train_blocking_ms / K, 5),
                  "blocking_what": "the operator's default contract (frame complete when forward returns): the host "
                                   "waits for the pair count inside every forward",
                  "iters_per_s_eager_torch_loss": world * K / (train_eager_ms * 1e-3),
                  "gpu_launches": launches_train},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "Mpixels/s", "ms_per_step": e2e_ms / K,
                "ms_per_step_runs": [round(t / K, 5) for t in e2e_runs],
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "frames_rendered_twice": e2e_redone[0], "streams": NS,
                "GBps_to_host": round(world * K * d2h_bytes / (e2e_ms * 1e-3) / 1e9, 2),
                "host_d2h_ceiling_GBps": round(d2h_gbs, 2),
                "frac_of_host_ceiling": round((world * K * d2h_bytes / (e2e_ms * 1e-3) / 1e9) / d2h_gbs, 3),
                "host_binding": (f"rank bound to {len(cores)} host cores before pinned allocation" if cores else "none"),
                "host_ceiling_what": "all ranks copying 6.2 MB frames device -> pinned host memory back to back, nothing else "
                                     "running (max over ranks, same timing harness): the rate this box absorbs",
                "cuda_graphs": not args.no_graphs,
                "what": "robosimgs_b200.sweep.SceneRenderer.submit/collect per frame: camera (view, proj, campos) from "
                        "pinned host memory, forward with deferred pair check writing the 8-bit RGB frame from the compositing kernel (B200GS_OUT_RGB8), finished frame "
                        "copied to pinned host memory and collected by the consumer; one captured CUDA graph per frame "
                        "slot, consecutive frames alternate streams; scene resident in HBM as in the reference's render loop"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                     "frac": dom_gbs / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernels": kernels,
                     "bound_note": "the contract's two bounds are hbm | tensor; the dominant kernel (compositing) is neither: "
                                   "it is bound by FP32 instruction issue and dependent-instruction latency -- "
                                   "kernels[].issue_frac gives the share of the GPU's issue slots each launch used",
                     "algorithmic_bytes_per_launch": per_stage_bytes[dom], "ms_per_launch": st_ms[dom],
                     "frame_fwd": {"bytes": bytes_fwd, "GBps": bytes_fwd / (fwd_ms / K * 1e-3) / 1e9,
                                   "frac": bytes_fwd / (fwd_ms / K * 1e-3) / 1e9 / peak,
                                   "GBps_single_stream": bytes_fwd / (serial_ms / K * 1e-3) / 1e9},
                     "frame_bwd": {"bytes": bytes_bwd},
                     "stage_ms_fwd": {k: round(v, 4) for k, v in st_ms.items()},
                     "stage_ms_train": {k: round(v, 4) for k, v in st_ms_train.items()},
                     "stage_GBps": stage_gbs,
                     "note": "render/render_bwd are FP32-issue/MUFU bound, not HBM bound (SURVEY 8(d)); the HBM "
                             "fraction is reported as the contract asks"},
    }
    # ---- CPU baseline (oracle port, bounded sample) on rank 0, N == 1 only ----
    if world == 1 and not args.no_cpu_baseline:
        from oracle import gs_oracle
        cores = os.cpu_count() or 1
        gs_oracle.set_num_threads(cores)
        sc_cpu = {k: tens[k].cpu().numpy() for k in names}
        rs_cpu = settings_from_camera(cams[Wm], SH_DEG)
        kw = dict(shs=sc_cpu["shs"], scales=sc_cpu["scales"], rotations=sc_cpu["rotations"], dtype=np.float32)
        gs_oracle.forward(rs_cpu, sc_cpu["means3D"], sc_cpu["opacities"], **kw)      # warm
        n, t0 = 0, time.perf_counter()
        while n < 4 and time.perf_counter() - t0 < 15.0:
            st = gs_oracle.forward(rs_cpu, sc_cpu["means3D"], sc_cpu["opacities"], **kw)
            n += 1
        dt = time.perf_counter() - t0
        pre = gs_oracle.preprocess_only(rs_cpu, sc_cpu["means3D"], sc_cpu["opacities"], sc_cpu["shs"],
                                        sc_cpu["scales"], sc_cpu["rotations"])
        pre(); t1 = time.perf_counter(); pre(); pre_s = time.perf_counter() - t1
        out["cpu_baseline"] = {
            "value": n * px / dt / 1e6, "unit": "Mpixels/s", "cores": gs_oracle.num_threads(), "kind": "port",
            "sample": f"{n} full 1920x1080 frames of the same scene/camera (oracle/gs_oracle.c fp32, OpenMP)",
            "preprocess_only_ms": pre_s * 1e3, "D_reference_rects": int(st.num_rendered),
            "threads_per_stage": {"preprocess": gs_oracle.num_threads(), "bin_and_sort": 1, "render": gs_oracle.num_threads()},
            "note": "a checker, not a tuned CPU renderer: the binning/sort stage of the oracle is serial (17 M pairs, about "
                    "1 s of the frame); the GPU/CPU ratio says nothing about kernel quality",
            "cpu_count": cores,
        }
        out["config"]["D_reference_rects"] = int(st.num_rendered)
        # ---- parity of THIS run against the oracle frame just rendered (same camera, same scene bits):
        # the frame comes out of the timed sweep path (graph replay, deferred pair check, policy bins)
        with torch.no_grad():
            for _ in range(2):
                gpu_frame = sweep_r.collect(sweep_r.submit(cams[Wm], cam_block=cam_blocks[Wm])).clone()
            _, radii_gpu = render(settings_dev[Wm])
        torch.cuda.synchronize()
        img = gpu_frame.cpu().numpy().astype(np.float64)
        mse = float(((img - st.color.astype(np.float64)) ** 2).mean())
        parity = {"psnr_db": 10.0 * math.log10(1.0 / max(mse, 1e-30)),
                  "max_abs_err": float(np.abs(img - st.color).max()),
                  "radii_mismatch": float((radii_gpu.cpu().numpy() != st.radii).mean()),
                  "frame": "sweep path (SceneRenderer graph replay) vs oracle/gs_oracle.c fp32, camera of step 0",
                  "tolerance": "north_star: PSNR >= 60 dB, grad max-rel-err < 1e-3"}
        if not args.no_parity_grad:
            st64 = gs_oracle.forward(settings_from_camera(jittered_cameras(1, first=0)[0], SH_DEG), sc_cpu["means3D"],
                                     sc_cpu["opacities"], shs=sc_cpu["shs"], scales=sc_cpu["scales"],
                                     rotations=sc_cpu["rotations"], dtype=np.float64)
            lv = {k: tens[k].detach().clone().requires_grad_(True) for k in names}
            m2g = torch.zeros_like(lv["means3D"], requires_grad=True)
            col, _ = GaussianRasterizer(rs_train)(lv["means3D"], m2g, lv["opacities"], shs=lv["shs"],
                                                  scales=lv["scales"], rotations=lv["rotations"])
            fused_mse_loss(col, target).backward()
            torch.cuda.synchronize()
            dL = (2.0 * (col.detach().cpu().double() - target.cpu().double()) / target.numel()).numpy()
            ref = gs_oracle.backward(st64, dL)
            errs = {}
            for k in names:
                r_ = getattr(ref, k)
                g_ = lv[k].grad.detach().cpu().double().numpy().reshape(r_.shape)
                errs[k] = float(np.abs(g_ - r_).max() / max(np.abs(r_).max(), 1e-30))
            parity["grad_max_rel"] = max(errs.values())
            parity["grad_max_rel_per_tensor"] = errs
            parity["grad"] = "train step (fwd + fused MSE + bwd) vs gs_oracle.backward fp64, max|got-ref|/max|ref| per tensor"
            del lv, m2g, col
        out["parity"] = parity
        # ---- GPU yardstick for the ">= 1.5x the reference rasterizer" target: the plain-SIMT restatement of the
        # published design (baseline/naive_simt.cu; the real library cannot be installed here), same scene/camera
        try:
            from baseline.naive_runner import NaiveRasterizer
            nv = NaiveRasterizer(tens, H_IMG, W_IMG, SH_DEG)
            rs0 = settings_dev[Wm]
            nv.forward(rs0)
            torch.cuda.synchronize()
            n_mse = float(((nv.out - gpu_frame).double() ** 2).mean())
            n_ms = timed(lambda s_: nv.forward(rs0), 20)
            out["gpu_comparator"] = {
                "kind": "naive SIMT restatement of the published rasterizer (baseline/naive_simt.cu), not the real library",
                "ms": n_ms / 20, "mpixels_per_s": px / (n_ms / 20 * 1e-3) / 1e6,
                "psnr_vs_b200gs_db": 10.0 * math.log10(1.0 / max(n_mse, 1e-30)),
                "ratio_single_stream": (n_ms / 20) / (serial_ms / K), "ratio_sweep": (n_ms / 20) / (fwd_ms / K)}
        except Exception as e:          # yardstick only: never fail the bench over it
            out["gpu_comparator"] = {"unavailable": repr(e)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--repeats", type=int, default=3, help="timed runs of K steps each; the median is reported")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-grad", action="store_true", help="skip the fp64 oracle gradient check of the parity block")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams the frames of the sweep alternate over")
    ap.add_argument("--e2e-streams", type=int, default=0,
                    help="same, for the end-to-end (host buffers) measurement; 0 = sweep.host_frames_in_flight "
                         "(6 for one or two ranks on the box, 4 from four ranks up where the host path is the bound)")
    ap.add_argument("--no-graphs", action="store_true", help="end-to-end path without CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
