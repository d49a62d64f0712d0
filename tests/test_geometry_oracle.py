"""CPU: the geometry oracle (oracle/geometry_oracle.py) against the fixture produced by the reference's own
GaussianPointCloud.update_geometry / get_radius / bbox_filter (tests/golden/make_scale_init_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_oracle as go
from oracle import oracle

CASES = ["a", "b", "c", "d"]
CFG = dict(min_radius=0.001, max_radius=0.05, scale_factor=1.0, xyz_factor=(1.0, 1.0, 0.1))


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "scale_init.npz"))


@pytest.mark.parametrize("case", CASES)
def test_update_geometry_matches_reference(golden, case):
    g = golden
    t = lambda k: torch.from_numpy(g[case + "_" + k])
    xyz, ls, ex, er = t("xyz"), t("log_scaling"), t("extra_xyz"), t("extra_radius")
    assert torch.equal(go.get_radius(ls), t("radius"))
    if ex.shape[0]:
        assert np.array_equal(go.bbox_filter(xyz, ex).numpy(), g[case + "_inbbox"])
    log_scales, invalid = go.update_geometry(xyz, ls, ex, er, knn_fn=oracle.knn, **CFG)
    assert np.array_equal(invalid.numpy(), g[case + "_invalid"])
    assert (log_scales is not None) == bool(g[case + "_scaling_updated"])
    if log_scales is not None:
        assert torch.equal(log_scales, t("new_scaling"))  # same torch ops in the same order: bit-identical


def test_fixture_covers_the_branches(golden):
    g = golden
    assert 0 < g["b_invalid"].sum() < g["b_invalid"].size            # mixed survivors / deletions
    assert g["c_invalid"].all() and not bool(g["c_scaling_updated"])   # everything deleted: scaling untouched
    assert not g["d_invalid"].any()
    assert np.exp(g["d_new_scaling"]).max() <= 0.05 + 1e-7           # clipped to max_radius


def test_knn_points3_contract():
    g = torch.Generator().manual_seed(3)
    q, r = torch.rand(200, 3, generator=g), torch.rand(500, 3, generator=g)
    d2, idx = go.knn_points3(q, r)
    full = ((q[:, None] - r[None]) ** 2).sum(-1)
    assert torch.allclose(d2, torch.sort(full, dim=1).values[:, :3], atol=1e-7)
    assert bool((d2[:, 0] <= d2[:, 1]).all() and (d2[:, 1] <= d2[:, 2]).all())
    assert torch.allclose(full.gather(1, idx), d2, atol=1e-7)
    d2, idx = go.knn_points3(q, r[:2])       # fewer than K reference points: zero padding
    assert bool((d2[:, 2] == 0).all() and (idx[:, 2] == 0).all())
    assert go.temp_points_filter(q, r[:0], torch.zeros(0)) is None
