"""GPU parity of the training-step neighbours: fused SSIM and fused Adam against
plain PyTorch references of the same operations (the tier's rule for floating-point kernels)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from robosimgs_b200 import _cabi
    _cabi.lib()


def _torch_ssim(img1, img2, window_size=11, sigma=1.5):
    """The SSIM of the public 3DGS trainers, restated in plain PyTorch (float64 for the reference)."""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / (2 * sigma ** 2)) for x in range(window_size)],
                     dtype=img1.dtype, device=img1.device)
    g = g / g.sum()
    C = img1.shape[0]
    win = (g[:, None] @ g[None, :]).expand(C, 1, window_size, window_size).contiguous()
    x, y = img1[None], img2[None]
    conv = lambda t: F.conv2d(t, win, padding=window_size // 2, groups=C)
    mu1, mu2 = conv(x), conv(y)
    s1, s2, s12 = conv(x * x) - mu1 * mu1, conv(y * y) - mu2 * mu2, conv(x * y) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return m.mean()


@pytest.mark.parametrize("shape", [(3, 67, 93), (3, 128, 160), (1, 16, 16), (3, 9, 300)])
def test_fused_ssim_matches_torch(shape):
    from robosimgs_b200.losses import ssim
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.rand(shape, generator=g)
    b = (a + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)        # correlated, like a render vs its target
    x = a.to(dev).requires_grad_(True)
    val = ssim(x, b.to(dev))
    val.backward()
    xr = a.double().to(dev).requires_grad_(True)
    ref = _torch_ssim(xr, b.double().to(dev))
    ref.backward()
    assert abs(float(val) - float(ref)) < 2e-6
    err = float((x.grad.double() - xr.grad).abs().max() / xr.grad.abs().max())
    assert err < 1e-4, err
    # forward only (no derivative maps stored) gives the same value
    with torch.no_grad():
        assert abs(float(ssim(a.to(dev), b.to(dev))) - float(val)) < 1e-6      # atomic summation order


def test_gs_loss_matches_torch_and_scales_upstream():
    from robosimgs_b200.losses import gs_loss
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(3, 121, 201, generator=g), torch.rand(3, 121, 201, generator=g)     # numel % 4 != 0
    x = a.to(dev).requires_grad_(True)
    val = 3.0 * gs_loss(x, b.to(dev), 0.2)
    val.backward()
    xr = a.double().to(dev).requires_grad_(True)
    bd = b.double().to(dev)
    ref = 3.0 * (0.8 * (xr - bd).abs().mean() + 0.2 * (1.0 - _torch_ssim(xr, bd)))
    ref.backward()
    assert abs(float(val) - float(ref)) < 1e-5
    err = float((x.grad.double() - xr.grad).abs().max() / xr.grad.abs().max())
    assert err < 1e-4, err


def test_fused_adam_matches_torch_adam():
    from robosimgs_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    shapes = [(1000, 3), (1000, 16, 3), (1000, 1), (1000, 3), (1000, 4), (7,), (4099,), (3, 5), (2, 2), (33,)]
    lrs = [1.6e-4, 2.5e-3, 5e-2, 5e-3, 1e-3, 1e-2, 1e-3, 1e-3, 1e-3, 1e-3]     # > 8 tensors: two launches
    init = [torch.randn(s, generator=g) for s in shapes]
    ours = [t.clone().to(dev).requires_grad_(True) for t in init]
    theirs = [t.clone().to(dev).requires_grad_(True) for t in init]
    opt_a = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(ours, lrs)], betas=(0.9, 0.999), eps=1e-15)
    opt_b = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(theirs, lrs)], betas=(0.9, 0.999), eps=1e-15)
    for step in range(5):
        grads = [torch.randn(s, generator=g) * (0.1 + step) for s in shapes]
        for p, q, gr in zip(ours, theirs, grads):
            p.grad = gr.to(dev)
            q.grad = gr.to(dev).clone()
        opt_a.step()
        opt_b.step()
    for p, q in zip(ours, theirs):
        assert torch.allclose(p, q, rtol=2e-5, atol=1e-7), float((p - q).abs().max())
    opt_a.zero_grad()
    assert all(p.grad is None for p in ours)
