"""Host-side logic of the operator layer that needs no GPU: bin-size policy, pair tickets, defer options."""
import torch

from robosimgs_b200 import rasterizer as rz


def test_default_bin_shift_matches_the_library_rule():
    # smallest shift with at most 255 bins, capped at 3 (api.cu:bin_shift_for, global-sort path)
    assert rz._default_bin_shift(1080, 1920) == 3
    assert rz._default_bin_shift(800, 800) == 2
    assert rz._default_bin_shift(480, 640) == 2
    assert rz._default_bin_shift(256, 256) == 1
    assert rz._default_bin_shift(64, 64) == 0
    assert rz._default_bin_shift(4320, 7680) == 3


def test_bin_size_policy_tracks_the_splat_extent():
    key = ("cpu-test", 1000, 1080, 1920)
    rz._BIN_POLICY.pop(key, None)
    rz._PAIR_HINTS[key] = 123
    pol, used, flags = rz._bin_flags(key, 1080, 1920)
    assert used == 3 and flags == 0 and pol["shift"] == -1
    radii = torch.zeros(1000, dtype=torch.int32)
    radii[:400] = 5
    # C3-like: 2.08 pairs per touching Gaussian at 128-px bins -> extent ~57 px -> bins of ~170 px -> keep 128
    rz._adapt_bin_size(pol, used, key, int(2.08 * 400), radii, lambda: 0.9)
    assert pol["shift"] == -1 and rz._PAIR_HINTS.get(key) == 123
    # ... and one size coarser when the frame saturates everywhere
    pol["calls"] = 0
    rz._adapt_bin_size(pol, used, key, int(2.08 * 400), radii, lambda: 0.9999)
    assert pol["shift"] == 4 and key not in rz._PAIR_HINTS
    pol["shift"], pol["calls"] = -1, 0
    rz._PAIR_HINTS[key] = 123
    # small splats: 1.16 pairs per touching Gaussian -> extent ~10 px -> 32-px bins; pair hint dropped
    pol["calls"] = 0
    rz._adapt_bin_size(pol, used, key, int(1.16 * 400), radii)
    assert pol["shift"] == 1 and key not in rz._PAIR_HINTS
    pol2, used2, flags2 = rz._bin_flags(key, 1080, 1920)
    assert used2 == 1 and flags2 == (2 << 8)
    # only the first call (and every 256th) looks; in between nothing changes
    rz._adapt_bin_size(pol2, used2, key, 4 * 400, radii)
    assert pol2["shift"] == 1 and pol2["calls"] == 2
    rz._BIN_POLICY.pop(key, None)


def test_pair_ticket_and_defer_options():
    key = ("cpu-test", 10, 8, 8)
    word = torch.zeros(1, dtype=torch.int32)
    rz._PAIR_HINTS.pop(key, None)
    t = rz.PairTicket(1000, key, word)
    word[0] = 900
    assert t.ok() and t.pairs == 900 and rz._PAIR_HINTS[key] == 900
    assert t.ok()                                   # idempotent
    t2 = rz.PairTicket(1000, key, word)
    word[0] = 5000
    assert not t2.ok() and rz._PAIR_HINTS[key] == 5000      # the hint recovers from an overflow
    opts = rz.DeferOptions(capacity=4096, word=word, record_event=False)
    assert opts.capacity == 4096 and opts.word is word and not opts.record_event and len(opts) == 0
    rz._PAIR_HINTS.pop(key, None)


def test_scene_renderer_capacity_never_shrinks_on_overflow():
    """Frames in flight were submitted under different capacities: an overflow report from an older frame
    (small pair count, small old capacity) must not undo the growth a later frame already caused."""
    from robosimgs_b200.sweep import SceneRenderer
    r = SceneRenderer.__new__(SceneRenderer)          # capacity logic only: no device needed
    r.capacity = 0
    r._set_capacity(3_000_000)
    c3 = r.capacity
    assert c3 >= 3_000_000 and c3 % 65536 == 0
    r._set_capacity(5_000_000, grow_only=True)
    c5 = r.capacity
    assert c5 > c3
    r._set_capacity(3_200_000, grow_only=True)        # stale overflow report from an older frame
    assert r.capacity == c5
    r._set_capacity(0)
    assert r.capacity == 0


def test_frames_in_flight_follow_the_ranks_on_the_box():
    # one or two ranks are kernel/launch bound (six frames hide the binning latency); from four ranks up the
    # device->host path of the box is the bound and more frames in flight only add concurrent copies
    from robosimgs_b200.sweep import host_frames_in_flight
    assert [host_frames_in_flight(n) for n in (1, 2, 4, 8)] == [6, 6, 4, 4]


def test_backward_or_retry_repeats_an_overflowed_step_and_gives_up_after_the_retries():
    """train.backward_or_retry: a step whose backward reports PairCapacityExceeded (the deferred training mode of the
    rasterizer) is rendered again; other errors and a persistent overflow propagate."""
    import pytest
    from robosimgs_b200 import PairCapacityExceeded
    from robosimgs_b200.train import backward_or_retry

    class Overflowing(torch.autograd.Function):
        fails_left = 0

        @staticmethod
        def forward(ctx, x):
            return x * 2.0

        @staticmethod
        def backward(ctx, g):
            if Overflowing.fails_left > 0:
                Overflowing.fails_left -= 1
                raise PairCapacityExceeded("9 pairs, capacity 4: run the step again")
            return g * 2.0

    x = torch.ones(3, requires_grad=True)
    calls = []

    def loss_fn(scale):
        calls.append(scale)
        return (Overflowing.apply(x) * scale).sum()

    Overflowing.fails_left = 2
    loss = backward_or_retry(loss_fn, 0.5, retries=2)
    assert len(calls) == 3 and float(loss) == 3.0
    assert torch.equal(x.grad, torch.ones(3))            # only the step that completed reached the leaf
    x.grad = None
    Overflowing.fails_left = 3
    with pytest.raises(PairCapacityExceeded):
        backward_or_retry(loss_fn, 0.5, retries=2)
    assert x.grad is None
    assert issubclass(PairCapacityExceeded, rz._cabi.B200GSError)


def test_defer_options_carry_the_training_and_rgb8_switches():
    o = rz.DeferOptions()
    assert o.train is False and o.rgb8 is None and o.capacity == 0 and o.record_event
    buf = torch.empty((4, 6, 3), dtype=torch.uint8)
    o = rz.DeferOptions(train=True, rgb8=buf)
    assert o.train and o.rgb8 is buf and len(o) == 0
    r = rz.GaussianRasterizer(None)
    assert r.defer_pair_check is False and r.last_ticket is None       # the operator's default contract: blocking
