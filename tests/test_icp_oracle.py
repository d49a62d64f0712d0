"""CPU: pins oracle/icp_oracle.py to tests/golden/icp.npz (outputs of the reference's own SLAM/icp.py / SLAM/utils.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import icp_oracle as io  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "icp.npz"))


def test_vertex_normal_pyramids_bit_exact():
    K = torch.from_numpy(G["K"])
    for f in ("0", "1"):
        d = torch.from_numpy(G["depth" + f])
        vp = io.build_vertex_pyramid(d.view(*d.shape, 1), [2, 1, 0], K.clone())
        for lvl in range(3):
            assert np.array_equal(vp[lvl].numpy(), G["vertex%s_l%d" % (f, lvl)])
            assert np.array_equal(io.compute_normal_map(vp[lvl]).numpy(), G["normal%s_l%d" % (f, lvl)])


def test_icp_matches_reference_run():
    K = torch.from_numpy(G["K"])
    t = lambda n: torch.from_numpy(G[n])
    Kd = K * 0.25
    Kd[2, 2] = 1.0
    res, J, valid = io.residuals_jacobian(t("vertex1_l0"), t("vertex0_l0"), t("normal1_l0"), t("normal0_l0"), torch.eye(4), Kd,
                                          0.1, float(np.cos(np.deg2rad(20))))
    assert int(valid.sum()) == int(G["first_valid"])
    np.testing.assert_allclose((J.T @ J).numpy(), G["first_jtj"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose((J.T @ res).numpy(), G["first_jtr"].reshape(6), rtol=1e-4, atol=1e-5)
    pose, ratio = io.predict_pose(t("depth0").view(120, 160, 1), t("depth1").view(120, 160, 1), K)
    np.testing.assert_allclose(pose.numpy(), G["pose_after_l2"], rtol=0, atol=1e-5)
    assert abs(float(ratio) - float(G["ratio_l2"])) < 1e-6
