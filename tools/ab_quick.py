"""A/B of library builds on one GPU box (dev tool): for each libb200gs build given on the command line, the C3 sweep
rate (SceneRenderer, 4 streams, graphs), the single-stream frame time with per-stage times, and the train step.
usage: python tools/ab_quick.py tag=path/to/lib.so[,ENV=VALUE...] [tag=...]   (runs each in a fresh process)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import statistics, torch
    from robosimgs_b200 import GaussianRasterizer, _cabi
    from robosimgs_b200.losses import mse_loss
    from robosimgs_b200.scenes import room_scene, room_target, settings_from_camera
    from robosimgs_b200.sweep import SceneRenderer
    import bench
    dev = torch.device("cuda:0")
    sc, _ = room_scene()
    tens = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    K = 60
    cams = bench.jittered_cameras(K + 3)
    blocks = torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1)]) for c in cams]).to(dev)
    bg = torch.zeros(3, device=dev)
    r = SceneRenderer(tens, 3, bg, 1080, 1920, streams=int(os.environ.get('AB_STREAMS', '4')), graphs=True, host_frames=False)
    pend = []
    def sweep(n):
        for s in range(n):
            while len(pend) >= r.in_flight_limit():
                r.collect(pend.pop(0))
            pend.append(r.submit(cams[s % len(cams)], cam_block=blocks[s % len(cams)]))
        while pend:
            r.collect(pend.pop(0))
    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); r.fs.fork(); fn(); r.fs.join(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    ONLY_TRAIN = os.environ.get('AB_ONLY', '') == 'train'      # skip the sweep and single-stream passes
    with torch.no_grad():
        sweep(24)
        sw = [timed(lambda: sweep(K)) / K for _ in range(1 if ONLY_TRAIN else 5)]
    # host-frame path (bench.py's e2e): camera from pinned host memory in, 8-bit frame in pinned host memory out
    e2e = []
    if not ONLY_TRAIN:
        rh = SceneRenderer(tens, 3, bg, 1080, 1920, streams=int(os.environ.get('AB_E2E_STREAMS', '6')), graphs=True, host_frames=True)
        ph = []
        def sweep_h(n):
            for s in range(n):
                while len(ph) >= rh.in_flight_limit():
                    rh.collect(ph.pop(0))
                ph.append(rh.submit(cams[s % len(cams)]))
            while ph:
                rh.collect(ph.pop(0))
        def timed_h(fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); rh.fs.fork(); fn(); rh.fs.join(); b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b)
        with torch.no_grad():
            sweep_h(36)
            e2e = [round(timed_h(lambda: sweep_h(K)) / K, 4) for _ in range(5)]
        rh.close(); del rh
    m2d = torch.zeros_like(tens["means3D"])
    rs = [settings_from_camera(c, 3, device=dev) for c in cams]
    def single(n):
        for s in range(n):
            GaussianRasterizer(rs[s % len(rs)])(tens["means3D"], m2d, tens["opacities"], shs=tens["shs"], scales=tens["scales"], rotations=tens["rotations"])
    with torch.no_grad():
        single(10)
        _cabi.profile_enable(True); _cabi.profile_read(True)
        ss = [timed(lambda: single(K)) / K for _ in range(1 if ONLY_TRAIN else 3)]
        st = _cabi.profile_read(True); _cabi.profile_enable(False)
    leaves = {k: v.clone().requires_grad_(True) for k, v in tens.items()}
    m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
    target = room_target().to(dev)
    rast = GaussianRasterizer(rs[0])
    rast.defer_pair_check = os.environ.get('AB_TRAIN_DEFER', '0') == '1'
    def train(n):
        for s in range(n):
            for v in leaves.values(): v.grad = None
            c, _ = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
            mse_loss(c, target).backward()
    train(8)
    tr = [timed(lambda: train(30)) / 30 for _ in range(5)]
    _cabi.profile_enable(True); _cabi.profile_read(True)
    train(10); torch.cuda.synchronize()
    st2 = _cabi.profile_read(True); _cabi.profile_enable(False)
    st.update({k: v for k, v in st2.items() if k in ("render_bwd", "project_bwd")})
    print(json.dumps({"sweep_ms": round(statistics.median(sw), 4), "sweep_runs": [round(x, 4) for x in sw], "single_ms": round(statistics.median(ss), 4),
                      "e2e_ms": (statistics.median(e2e) if e2e else None), "train_ms": round(statistics.median(tr), 4), "train_runs": [round(x, 4) for x in tr], "redone": r.redone, "prio_nodes": list(r.priority_nodes),
                      "stages": {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]}}))
    sys.exit(0)

for spec in sys.argv[1:]:
    tag, rest = spec.split("=", 1)
    path, *envs = rest.split(",")
    env = dict(os.environ, B200GS_LIB_PATH=os.path.abspath(path))
    env.update(dict(e.split("=", 1) for e in envs))
    out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    print(tag, line, flush=True)
