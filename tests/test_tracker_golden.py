"""CPU: sanity of tests/golden/icp.npz (outputs of the reference's own SLAM/icp.py, tests/golden/make_icp_golden.py):
the reference's coarse-to-fine ICP recovers the synthetic ground-truth motion, so the fixture is a meaningful pin."""
import os

import numpy as np

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "icp.npz"))


def test_reference_icp_fixture_is_consistent():
    assert G["vertex0_l2"].shape == (120, 160, 3) and G["vertex0_l0"].shape == (30, 40, 3)
    np.testing.assert_allclose(G["pose_after_l2"], G["true_pose10"], atol=2e-3)
    assert np.abs(G["pose_after_l0"] - G["true_pose10"]).max() > np.abs(G["pose_after_l2"] - G["true_pose10"]).max()
    assert G["first_jtj"].shape == (6, 6) and np.allclose(G["first_jtj"], G["first_jtj"].T, atol=1e-3)
    n = G["normal0_l2"]
    norms = np.linalg.norm(n, axis=-1)
    assert np.all((norms < 1e-6) | (np.abs(norms - 1) < 1e-4))
