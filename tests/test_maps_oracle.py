"""CPU: pins oracle/maps_oracle.py to the fixtures produced by the reference's own SLAM/utils.py
(tests/golden/make_maps_golden.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import maps_oracle as mo  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))


def test_transmission2tilemask_matches_reference_fixtures():
    for name in "abc":
        pm = torch.from_numpy(G[name + "_pixelmask"])
        for ratio in (0.5, 0.3):
            got = mo.transmission2tilemask(pm, 16, ratio).numpy()
            assert np.array_equal(got, G["%s_tm_%02d" % (name, int(ratio * 10))])


def test_colorerror2tilemask_matches_reference_fixtures():
    for name in "abc":
        err = torch.from_numpy(G[name + "_error"])
        for ratio in (0.4, 0.1):
            got, _ = mo.colorerror2tilemask(err, 16, ratio)
            assert np.array_equal(got.numpy(), G["%s_ce_%02d" % (name, int(ratio * 10))].astype(np.int32))


def test_evaluate_render_range_cases():
    g = torch.Generator().manual_seed(3)
    H, W = 37, 53
    T = torch.where(torch.rand(1, H, W, generator=g) > 0.4, torch.rand(1, H, W, generator=g), torch.ones(1, H, W))
    out = {"T_map": T, "render": torch.rand(3, H, W, generator=g) * (T != 1), "semantic_seg": None}
    gt = torch.rand(3, H, W, generator=g)
    rm, tm, ratio = mo.evaluate_render_range(out)
    assert rm.dtype == torch.bool and tm.shape == (3, 4) and abs(float(ratio) - float(rm.sum()) / (H * W)) < 1e-7
    rm2, tm2, _ = mo.evaluate_render_range(out, global_opt=True)
    assert tm2 is None and torch.equal(rm, rm2)
    rm3, tm3, _ = mo.evaluate_render_range(out, gt_image=gt, global_opt=True, sample_ratio=0.4)
    assert int(tm3.sum()) == int(12 * 0.4) and rm3.shape == (H, W)
    assert torch.equal(rm3, tm3.bool().repeat_interleave(16, 0).repeat_interleave(16, 1)[:H, :W])
