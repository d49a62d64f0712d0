"""CPU property tests (hypothesis) of the host logic and of the oracles: invariants that hold for every input, next
to the example-based tests.  No GPU, no compute call into the library."""
import os
import sys

import torch
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dqo_map_b200 import binning_policy as bp, sharding  # noqa: E402
from oracle import ssim_oracle as so  # noqa: E402

FAST = settings(max_examples=60, deadline=None, derandomize=True)


@FAST
@given(st.lists(st.integers(0, 500_000), min_size=0, max_size=90), st.integers(1, 8))
def test_assign_objects_is_a_partition_with_the_lpt_bound(counts, world):
    owner, load = sharding.assign_objects(counts, world)
    assert sorted(owner) == list(range(len(counts)))                        # every object exactly once
    assert all(0 <= r < world for r in owner.values())
    for r in range(world):
        assert load[r] == sum(c for i, c in enumerate(counts) if owner[i] == r)
        assert sharding.local_objects(owner, r) == sorted(i for i in owner if owner[i] == r)
    if counts:
        # greedy LPT: no rank exceeds the mean load by more than the largest item
        assert max(load) <= sum(counts) / world + max(counts)
    # deterministic and independent of dict ordering (every rank computes it locally)
    shuffled = dict(reversed(list(enumerate(counts))))
    assert sharding.assign_objects(shuffled, world)[0] == owner


status_words = st.tuples(st.integers(0, 1 << 27), st.floats(0, 1), st.floats(0, 1), st.floats(0, 1), st.booleans())


@FAST
@given(st.lists(status_words, min_size=1, max_size=30), st.integers(1 << 16, 1 << 28))
def test_binning_policy_plans_are_always_valid(calls, capacity):
    """Whatever sequence of status words the device reports, the plan respects the C-ABI contract: front is a multiple
    of 256, front + back fits the capacity, and both are zero together (single phase)."""
    pol = bp.BinningPolicy()
    for R, f_walk, f_front, f_back, overflow in calls:
        front, back = pol.plan(capacity)
        assert (front == 0) == (back == 0)
        assert front % 256 == 0 and 0 <= front and front + back <= capacity
        status = [0] * 8
        status[bp.ST_NUM_RENDERED] = R
        status[bp.ST_WALKED] = int(f_walk * R)
        status[bp.ST_R_FRONT] = min(front, int(f_front * R))
        status[bp.ST_R_BACK] = int(f_back * R)
        status[bp.ST_OVERFLOW] = int(overflow)
        repeat = pol.update(status, front, back)
        assert repeat == bool(overflow)
        assert len(pol.history) <= 8


@FAST
@given(st.integers(1 << 21, 1 << 27), st.floats(0.0, 0.3))
def test_binning_policy_enters_two_phase_under_heavy_occlusion_and_leaves_when_it_does_not_pay(R, walked_frac):
    pol = bp.BinningPolicy()
    status = [0] * 8
    status[bp.ST_NUM_RENDERED], status[bp.ST_WALKED] = R, int(walked_frac * R)
    pol.update(status, 0, 0)
    front, back = pol.plan(1 << 30)
    assert front > 0 and back > 0 and front <= R // 2 + 256
    status[bp.ST_R_FRONT], status[bp.ST_R_BACK] = front, R          # nothing finished early: give up
    pol.update(status, front, back)
    assert pol.plan(1 << 30) == (0, 0) and pol.cooldown > 0


images = st.tuples(st.integers(1, 40), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))


@settings(max_examples=25, deadline=None, derandomize=True)
@given(images)
def test_ssim_oracle_invariants(shape_seed):
    H, W, seed = shape_seed
    g = torch.Generator().manual_seed(seed)
    x, y = torch.rand(3, H, W, generator=g, dtype=torch.float64), torch.rand(3, H, W, generator=g, dtype=torch.float64)
    assert abs(float(so.ssim(x, x)) - 1.0) <= 1e-12                       # identical images
    s_xy = float(so.ssim(x, y))
    assert abs(s_xy - float(so.ssim(y, x))) <= 1e-12 and -1.0 <= s_xy <= 1.0
    assert abs(s_xy - so.ssim_numpy(x.numpy(), y.numpy())) <= 1e-10       # torch conv form == explicit tap loops
    loss, grad = so.ssim_loss_and_grad(x, y.permute(1, 2, 0))
    assert abs(loss - (1 - s_xy)) <= 1e-12 and grad.shape == x.shape and bool(torch.isfinite(grad).all())
    # directional derivative: the autograd gradient predicts a finite difference of the loss
    d = torch.randn(3, H, W, generator=g, dtype=torch.float64)
    eps = 1e-6
    fd = ((1 - float(so.ssim(x + eps * d, y))) - (1 - float(so.ssim(x - eps * d, y)))) / (2 * eps)
    assert abs(fd - float((grad * d).sum())) <= 1e-6 * max(1.0, abs(fd))
