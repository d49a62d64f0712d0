"""GPU: mask builders / error maps (csrc/maps.cu via the C-ABI) against the reference-generated fixtures and the
torch restatement in oracle/maps_oracle.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import maps_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))


def _dev():
    return torch.device("cuda:0")


def test_transmission2tilemask_fixtures():
    from dqo_map_b200 import map_utils
    for name in "abc":
        pm = torch.from_numpy(G[name + "_pixelmask"]).to(_dev())
        for ratio in (0.5, 0.3):
            got = map_utils.transmission2tilemask(pm, 16, ratio)
            assert got.dtype == torch.int32
            assert np.array_equal(got.cpu().numpy(), G["%s_tm_%02d" % (name, int(ratio * 10))])


def test_colorerror2tilemask_fixtures():
    from dqo_map_b200 import map_utils
    for name in "abc":
        err = torch.from_numpy(G[name + "_error"]).to(_dev())
        for ratio in (0.4, 0.1):
            got = map_utils.colorerror2tilemask(err, 16, ratio)
            assert np.array_equal(got.cpu().numpy(), G["%s_ce_%02d" % (name, int(ratio * 10))].astype(np.int32))


@pytest.mark.parametrize("H,W", [(680, 1200), (1080, 1920), (33, 17)])
def test_topk_tilemask_full_size(H, W):
    """Against torch on the same device at the bench sizes; ties (equal tile means, e.g. all-zero tiles) may be ordered
    differently by torch.topk, so mismatching tiles must have a mean equal to the k-th value."""
    from dqo_map_b200 import map_utils
    g = torch.Generator().manual_seed(11)
    err = torch.rand(H, W, generator=g)
    err[: H // 3] = 0  # a band of never-rendered tiles -> ties at zero
    err = err.to(_dev())
    for ratio in (0.4, 0.9):
        got, pix = map_utils.colorerror2tilemask(err, 16, ratio, return_pixel_mask=True)
        want, means = mo.colorerror2tilemask(err.cpu(), 16, ratio)
        k = int(want.numel() * ratio)
        assert int(got.sum()) == k
        kth = torch.sort(means.reshape(-1), descending=True).values[k - 1]
        bad = (got.cpu() != want)
        assert bool((means[bad] == kth).all())
        up = got.bool().repeat_interleave(16, 0).repeat_interleave(16, 1)[:H, :W]
        assert torch.equal(pix, up)


def test_color_error_map_and_or_into():
    from dqo_map_b200 import map_utils
    g = torch.Generator().manual_seed(5)
    H, W = 75, 130
    render = torch.rand(3, H, W, generator=g)
    render[:, :20] = 0
    gt = torch.rand(3, H, W, generator=g)
    got = map_utils.color_error_map(render.to(_dev()), gt.to(_dev()))
    want = mo.color_error_map(render, gt)
    assert torch.equal(got.cpu(), want)
    sem = torch.rand(3, H, W, generator=g)
    gsem = torch.rand(3, H, W, generator=g)
    tm = map_utils.colorerror2tilemask(got, 16, 0.3)
    map_utils.colorerror2tilemask(map_utils.color_error_map(sem.to(_dev()), gsem.to(_dev())), 16, 0.3, out=tm)
    want_tm = mo.colorerror2tilemask(want, 16, 0.3)[0] | mo.colorerror2tilemask(mo.color_error_map(sem, gsem), 16, 0.3)[0]
    assert torch.equal(tm.cpu(), want_tm)


@pytest.mark.parametrize("H,W", [(680, 1200), (37, 53)])
def test_evaluate_render_range_cases(H, W):
    from dqo_map_b200 import map_utils
    g = torch.Generator().manual_seed(7)
    coarse = torch.rand(1, 1, (H + 15) // 16, (W + 15) // 16, generator=g)
    up = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear")[0]
    T = torch.where(up + 0.2 * torch.rand(1, H, W, generator=g) > 0.6, torch.rand(1, H, W, generator=g), torch.ones(1, H, W))
    out = {"T_map": T, "render": torch.rand(3, H, W, generator=g) * (T != 1), "semantic_seg": None}
    gt = torch.rand(3, H, W, generator=g)
    dout = {k: (v.to(_dev()) if v is not None else None) for k, v in out.items()}
    for kw in ({}, {"global_opt": True}, {"global_opt": True, "sample_ratio": 0.4, "gt_image": gt}):
        want = mo.evaluate_render_range(out, **kw)
        dkw = {k: (v.to(_dev()) if torch.is_tensor(v) else v) for k, v in kw.items()}
        got = map_utils.evaluate_render_range(dout, **dkw)
        assert got[0].dtype == torch.bool and torch.equal(got[0].cpu(), want[0])
        assert (got[1] is None) == (want[1] is None)
        if want[1] is not None:
            assert torch.equal(got[1].cpu(), want[1])
        assert abs(float(got[2]) - float(want[2])) < 1e-7


def test_render_error_maps_feed_accumulate():
    from dqo_map_b200 import map_utils
    g = torch.Generator().manual_seed(9)
    H, W, P = 61, 83, 500
    out = {"render": torch.rand(3, H, W, generator=g), "depth": torch.rand(1, H, W, generator=g) * 4,
           "depth_index_map": torch.randint(-1, P, (1, H, W), generator=g, dtype=torch.int32),
           "color_index_map": torch.randint(-1, P, (1, H, W), generator=g, dtype=torch.int32)}
    color_map = torch.rand(H, W, 3, generator=g)
    depth_map = torch.rand(H, W, 1, generator=g) * 4
    depth_map[torch.rand(H, W, 1, generator=g) < 0.2] = 0
    want = mo.render_error_maps(out, color_map, depth_map)
    dout = {k: v.to(_dev()) for k, v in out.items()}
    got = map_utils.render_error_maps(dout, color_map.to(_dev()), depth_map.to(_dev()))
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a.cpu(), b)
    # and straight into the scatter, as error_gaussians_remove does (mapper.py:1026-1047)
    res = map_utils.accumulate_gaussian_error(H, W, P, got[0], got[1], got[2], dout["color_index_map"].permute(1, 2, 0),
                                              dout["depth_index_map"].permute(1, 2, 0), 0.1, 0.05, 0.3, True)
    ce = torch.zeros(P)
    ci = out["color_index_map"].reshape(-1).long()
    valid = ci >= 0
    ce.scatter_reduce_(0, ci[valid], want[0].reshape(-1)[valid], reduce="amax")
    assert torch.equal(res[0].cpu().reshape(-1), ce)
