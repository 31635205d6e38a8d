"""Host ceiling of the end-to-end (host-buffer) sweep on a multi-GPU box (dev tool, run under torchrun): every rank
copies 1080p 8-bit frames (6.2 MB) device -> pinned host memory back to back, all ranks at once, and rank 0 prints the
aggregate GB/s -- with the default placement and with each rank's thread + pinned buffers bound to its own slice of the
host cores (os.sched_setaffinity before the pinned allocation: first touch places the pages).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_ceiling.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
res = {}
for mode in ("default", "bound"):
    if mode == "bound":
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[rank * per:(rank + 1) * per]) or set(cpus))
    nbytes = 1080 * 1920 * 3
    src = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(4)]
    dst = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    for d in dst:
        d.zero_()                                   # first touch from this rank's cores
    s = torch.cuda.Stream()
    def burst(n):
        with torch.cuda.stream(s):
            for i in range(n):
                dst[i % 4].copy_(src[i % 4], non_blocking=True)
        s.synchronize()
    burst(50)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    burst(2000)
    dt = time.perf_counter() - t0
    t = torch.tensor([2000 * nbytes / dt / 1e9], device=dev, dtype=torch.float64)
    lo = t.clone()
    if world > 1:
        dist.all_reduce(t)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    res[mode] = {"aggregate_GBps": float(t), "slowest_rank_GBps": float(lo)}
if rank == 0:
    topo = os.popen("nvidia-smi topo -m 2>/dev/null | head -14").read()
    print(json.dumps({"n_gpus": world, "frame_bytes": 1080 * 1920 * 3, "d2h": res, "cpus": len(os.sched_getaffinity(0)) if False else os.cpu_count()}))
    print(topo)
if world > 1:
    dist.destroy_process_group()
