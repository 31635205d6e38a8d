"""Host-side logic of the camera-sharded sweep on CPU: world_size-2 gloo processes, a stand-in
render function (NOT the oracle -- only the sharding / replication plumbing is under test)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robosimgs_b200.sweep import SCENE_FIELDS, render_sweep, replicate_scene, shard_indices


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robosimgs_b200.cameras import orbit_cameras
    g = torch.Generator().manual_seed(0)
    scene = None
    if rank == 0:
        scene = {"means3D": torch.rand(50, 3, generator=g), "shs": torch.rand(50, 4, 3, generator=g),
                 "opacities": torch.rand(50, 1, generator=g), "scales": torch.rand(50, 3, generator=g),
                 "rotations": torch.rand(50, 4, generator=g)}
    rep = replicate_scene(scene, torch.device("cpu"))
    cams = orbit_cameras(7, (0, 0, 0), 2.0, 60.0, 32, 24)
    # stand-in renderer: a frame whose mean identifies (camera, replicated scene)
    render = lambda cam: torch.full((3, 24, 32), float(cam.campos.sum()) + float(rep["means3D"].sum()))
    seen = []
    mine, means = render_sweep(cams, render, on_frame=lambda f, fr: seen.append(f), gather=True)
    q.put((rank, mine, seen, means.tolist(), {k: float(rep[k].sum()) for k in SCENE_FIELDS}))
    dist.destroy_process_group()


def test_shard_indices_partition():
    for world in (1, 2, 3, 8):
        parts = [shard_indices(64, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_two_rank_sweep_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    (r0, mine0, seen0, means0, sums0), (r1, mine1, seen1, means1, sums1) = res
    assert mine0 == [0, 2, 4, 6] and mine1 == [1, 3, 5] and seen0 == mine0 and seen1 == mine1
    assert sums0 == sums1                                   # scene replicated bit-exactly
    assert means0 == means1 and all(m != 0 for m in means0)  # every frame accounted for on both ranks
