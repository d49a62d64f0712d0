"""CPU: the quadric oracle against the fixture produced by the reference's own quadrics.py
(tests/golden/make_quadric_golden.py)."""
import os

import numpy as np
import pytest

from oracle import quadric_oracle as qo


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "quadric.npz"))


def test_object_init(g):
    for i in range(g["init_bbox"].shape[0]):
        ax, R, c = qo.object_init(g["init_bbox"][i], g["init_depth_stats"][i], g["K"], g["init_Rt"][i])
        np.testing.assert_allclose(ax, g["init_axes"][i], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(R, g["init_R"][i], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(c, g["init_center"][i], rtol=1e-12, atol=1e-12)


def test_projection_bbox(g):
    for i in range(g["init_bbox"].shape[0]):
        bb, ell = qo.project_bbox(g["init_axes"][i], g["init_R"][i], g["init_center"][i], g["proj_P"][i])
        np.testing.assert_allclose(bb, g["proj_bbox"][i], rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(ell, g["proj_ellipse"][i], rtol=1e-8, atol=1e-8)


def test_refinement_matches_reference_loop(g):
    for i in range(g["init_bbox"].shape[0]):
        ax, R, c, _ = qo.refine(g["init_axes"][i], g["init_R"][i], g["init_center"][i], g["obs_bboxes"][i], g["Ps"][i],
                                g["view_choice"][i])
        np.testing.assert_allclose(ax, g["refined_axes"][i], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(R, g["refined_R"][i], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(c, g["refined_center"][i], rtol=2e-5, atol=2e-6)
