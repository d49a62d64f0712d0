"""GPU: tracker front-end (csrc/tracker.cu via the C-ABI) against tests/golden/icp.npz, produced by importing and running
the reference's own SLAM/icp.py + SLAM/utils.py in float32 on the CPU (tests/golden/make_icp_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "icp.npz"))
DEV = "cuda:0"


def _t(name):
    return torch.from_numpy(G[name]).to(DEV)


def test_vertex_and_normal_pyramids_match_reference():
    from dqo_map_b200 import icp
    K = _t("K")
    builder = icp.ImagePyramids([2, 1, 0], "max")
    for f in ("0", "1"):
        d = _t("depth" + f)
        vp = icp.build_vertex_pyramid(d.view(*d.shape, 1), builder, K)
        npyr = icp.build_normal_pyramid(vp)
        for lvl in range(3):
            v_ref, n_ref = G["vertex%s_l%d" % (f, lvl)], G["normal%s_l%d" % (f, lvl)]
            assert vp[lvl].shape == v_ref.shape
            np.testing.assert_allclose(vp[lvl].cpu().numpy(), v_ref, rtol=1e-6, atol=1e-6)
            # the Sobel sums run in a different order than the reference's conv2d: compare directions, except where the
            # cross product is numerically zero (flat depth steps) and the reference normalises noise
            got = npyr[lvl].cpu().numpy()
            both = (np.linalg.norm(n_ref, axis=-1) > 0.5) & (np.linalg.norm(got, axis=-1) > 0.5)
            assert (np.linalg.norm(n_ref, axis=-1) > 0.5).mean() > 0.2   # the far wall sits at depth.max() and is masked
            assert np.abs((got * n_ref).sum(-1)[both] - 1.0).max() < 1e-4
            assert ((np.linalg.norm(got, axis=-1) > 0.5) != (np.linalg.norm(n_ref, axis=-1) > 0.5)).mean() < 2e-3
        # the fused depth -> (vertex, normal) entry point equals the two-step path
        v, n = icp.vertex_normal_map(d, K)
        assert torch.equal(v, icp.compute_vertex_map(d, K)) and torch.equal(n, icp.compute_normal_map(v))


def test_icp_first_iteration_normal_equations():
    """J^T J, J^T r and the valid count of the first Gauss-Newton iteration at the coarsest level, recovered from the
    pose update of a 1-iteration run with zero damping: exp(xi) = pose since the initial pose is the identity."""
    from dqo_map_b200 import icp
    K = _t("K") * 0.25
    K[2, 2] = 1.0
    tr = icp.ICP(1, damping=1e-4, distance_threshold=0.1, normal_threshold=20)
    pose, ratio = tr.icp(torch.eye(4, device=DEV), _t("vertex1_l0"), _t("vertex0_l0"), _t("normal1_l0"), _t("normal0_l0"), K)
    H, W = G["vertex1_l0"].shape[:2]
    assert abs(float(ratio) * H * W - float(G["first_valid"])) <= 2
    JtJ, JtR = G["first_jtj"].astype(np.float64), G["first_jtr"].astype(np.float64).reshape(6)
    Hm = JtJ + np.trace(JtJ) * 1e-4 * np.eye(6)
    xi = -np.linalg.solve(Hm, JtR)
    w, v = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    Wh = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    E = np.eye(3) + Wh * np.sin(th) / th + Wh @ Wh * (1 - np.cos(th)) / th ** 2
    Jm = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Wh + (th - np.sin(th)) / th ** 3 * Wh @ Wh
    want = np.eye(4)
    want[:3, :3], want[:3, 3] = E, Jm @ v
    np.testing.assert_allclose(pose.cpu().numpy(), want, rtol=0, atol=2e-5)


def test_coarse_to_fine_pose_matches_reference_and_truth():
    from dqo_map_b200 import icp
    K = _t("K")
    builder = icp.ImagePyramids([2, 1, 0], "max")
    d0, d1 = _t("depth0"), _t("depth1")
    vp0, vp1 = icp.build_vertex_pyramid(d0, builder, K), icp.build_vertex_pyramid(d1, builder, K)
    np0, np1 = icp.build_normal_pyramid(vp0), icp.build_normal_pyramid(vp1)
    pose = torch.eye(4, device=DEV)
    for lvl, scale in enumerate((0.25, 0.5, 1.0)):
        Kd = K.clone() * scale
        Kd[2, 2] = 1.0
        pose, ratio = icp.ICP(5, damping=1e-4, distance_threshold=0.1, normal_threshold=20).icp(
            pose, vp1[lvl], vp0[lvl], np1[lvl], np0[lvl], Kd)
        # nearest-neighbour association flips on a handful of pixels (float rounding of the projection), so the
        # trajectories agree to ~1e-4 rather than to the last bit
        np.testing.assert_allclose(pose.cpu().numpy(), G["pose_after_l%d" % lvl], rtol=0, atol=5e-4)
        assert abs(float(ratio) - float(G["ratio_l%d" % lvl])) < 5e-3
    p2, r2 = icp.predict_pose(d0, d1, K)
    assert torch.allclose(p2, pose, atol=1e-6)
    np.testing.assert_allclose(pose.cpu().numpy(), G["true_pose10"], rtol=0, atol=2e-3)   # and it is the right answer


def test_against_pinned_oracle_on_fresh_views():
    """Other motion, other resolution (not a multiple of the pyramid stride): CUDA vs oracle/icp_oracle.py (pinned to the
    reference fixtures by tests/test_icp_oracle.py)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_icp_golden import synth_depth
    from dqo_map_b200 import icp
    from oracle import icp_oracle as io
    H, W = 187, 333
    Kn = np.array([[260.0, 0, 166.0], [0, 260.0, 93.0], [0, 0, 1.0]])
    p1 = np.eye(4)
    a = -0.015
    p1[:3, :3] = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    p1[:3, 3] = [-0.02, 0.015, -0.01]
    d0, d1 = torch.from_numpy(synth_depth(H, W, Kn, np.eye(4))), torch.from_numpy(synth_depth(H, W, Kn, p1))
    K = torch.from_numpy(Kn).float()
    want, want_ratio = io.predict_pose(d0.view(H, W, 1), d1.view(H, W, 1), K.clone())
    got, ratio = icp.predict_pose(d0.to(DEV).view(H, W, 1), d1.to(DEV).view(H, W, 1), K.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=5e-4)
    assert abs(float(ratio) - float(want_ratio)) < 5e-3
    np.testing.assert_allclose(got.cpu().numpy(), p1.astype(np.float32), rtol=0, atol=3e-3)
    # vertex maps are bit-identical to the oracle; degenerate input is left alone
    v_got = icp.compute_vertex_map(d0.to(DEV).view(H, W, 1), K.to(DEV))
    assert np.array_equal(v_got.cpu().numpy(), io.compute_vertex_map(d0.view(H, W, 1), K).numpy())
    z = torch.zeros(H, W, 3, device=DEV)
    pose, r0 = icp.ICP(3).icp(torch.eye(4, device=DEV), z, z, z, z, K.to(DEV))
    assert torch.equal(pose, torch.eye(4, device=DEV)) and float(r0) == 0.0
