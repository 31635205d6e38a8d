"""Camera-batch data-parallel step on CPU: world_size-2 gloo processes, a stand-in differentiable
render (NOT the oracle -- the bucket / all-reduce / replica-consistency plumbing is under test)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robosimgs_b200.train import GradBucket, dp_train_step


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make():
    g = torch.Generator().manual_seed(5)
    params = [torch.randn(40, 3, generator=g).requires_grad_(True), torch.randn(40, 4, 3, generator=g).requires_grad_(True),
              torch.randn(40, 1, generator=g).requires_grad_(True)]
    cams = [float(c) for c in range(1, 6)]                     # 5 "cameras": ragged over 2 ranks
    def loss_fn(cam):                                          # stand-in for render + photometric loss
        return ((params[0] * cam).sum() - 3.0) ** 2 + (params[1] ** 2).mean() * cam      # params[2] gets no gradient
    return params, cams, loss_fn


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, cams, loss_fn = _make()
    opt = torch.optim.SGD(params, lr=1e-3)
    bucket = GradBucket(params)
    for _ in range(3):
        dp_train_step(params, opt, cams, loss_fn, bucket)
    q.put((rank, [p.detach().reshape(-1).tolist() for p in params]))
    dist.destroy_process_group()


def test_grad_bucket_views_and_missing_grads():
    params, cams, loss_fn = _make()
    b = GradBucket(params)
    loss_fn(cams[0]).backward()
    g0 = params[0].grad.clone()
    b.pack(); b.all_reduce(); b.unpack()
    assert torch.equal(params[0].grad, g0) and params[0].grad.data_ptr() == b.views[0].data_ptr()
    assert torch.count_nonzero(params[2].grad) == 0 and b.flat.numel() == sum(p.numel() for p in params)


def test_two_rank_dp_step_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    # single-process reference: the same three steps over the whole camera batch
    params, cams, loss_fn = _make()
    opt = torch.optim.SGD(params, lr=1e-3)
    for _ in range(3):
        opt.zero_grad()
        sum(loss_fn(c) / len(cams) for c in cams).backward()
        opt.step()
    for a, b, ref in zip(res[0][1], res[1][1], params):
        assert a == b                                          # replicas stay bit-identical
        assert torch.allclose(torch.tensor(a), ref.detach().reshape(-1), rtol=1e-5, atol=1e-6)
