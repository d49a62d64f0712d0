#!/usr/bin/env python
"""Generates tests/golden/icp.npz by importing the reference's own SLAM/utils.py and SLAM/icp.py (unmodified, from
/root/reference) in THIS container and running, on the CPU in float32:

  * ImagePyramids("max") + build_vertex_pyramid / build_normal_pyramid   (SLAM/icp.py:342-360, SLAM/utils.py:65-125,542-559)
  * ICP.compute_residuals_jacobian / compute_jtj / compute_jtr            (SLAM/icp.py:52-121) for the first iteration
  * the coarse-to-fine pose prediction loop of IcpTracker.predict_pose     (SLAM/icp.py:424-441) via ICP.icp per level

on two synthetic depth maps (a wall with a sphere and a box, seen from two nearby poses).  Same stand-in technique as
make_maps_golden.py for the packages that are not installed here; no reference source is modified or copied.
Run:  python tests/golden/make_icp_golden.py"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_maps_golden import load_reference_utils  # noqa: E402

REF = os.environ.get("DQO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icp.npz")


def load_reference_icp():
    utils_mod = load_reference_utils()
    pkg = types.ModuleType("SLAM")
    pkg.__path__ = []
    pkg.utils = utils_mod
    sys.modules["SLAM"] = pkg
    sys.modules["SLAM.utils"] = utils_mod
    spec = importlib.util.spec_from_file_location("ref_slam_icp", os.path.join(REF, "SLAM/icp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return utils_mod, mod


def synth_depth(H, W, K, pose_c2w):
    """Ray-cast depth of a wall (z = 3), a sphere and a box from camera pose `pose_c2w` (numpy, float64)."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    j, i = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    d = np.stack([(i - cx) / fx, (j - cy) / fy, np.ones_like(i, dtype=np.float64)], -1)
    R, t = pose_c2w[:3, :3], pose_c2w[:3, 3]
    dw = d @ R.T
    o = t
    best = np.full((H, W), np.inf)
    # wall z = 3 (world)
    tw = (3.0 - o[2]) / dw[..., 2]
    best = np.where(tw > 0, np.minimum(best, tw), best)
    # sphere centre (0.3, 0.1, 2.0) radius 0.45
    c, r = np.array([0.3, 0.1, 2.0]), 0.45
    oc = o - c
    b = (dw * oc).sum(-1)
    a = (dw * dw).sum(-1)
    disc = b * b - a * ((oc * oc).sum() - r * r)
    ts = (-b - np.sqrt(np.maximum(disc, 0))) / a
    best = np.where((disc > 0) & (ts > 0), np.minimum(best, ts), best)
    # box: slab x in [-1.0, -0.4], y in [-0.3, 0.5], z in [1.6, 2.2]
    lo, hi = np.array([-1.0, -0.3, 1.6]), np.array([-0.4, 0.5, 2.2])
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (lo - o) / dw, (hi - o) / dw
    tn, tf = np.minimum(t1, t2).max(-1), np.maximum(t1, t2).min(-1)
    best = np.where((tn < tf) & (tn > 0), np.minimum(best, tn), best)
    depth = best  # ray parameter along d with d.z = 1 in the camera frame == camera-space depth
    depth[~np.isfinite(depth)] = 0
    return depth.astype(np.float32)


def main():
    ref_utils, ref_icp = load_reference_icp()
    H, W = 120, 160
    K = np.array([[140.0, 0, 79.5], [0, 140.0, 59.5], [0, 0, 1.0]])
    pose0 = np.eye(4)
    ang = 0.02
    pose1 = np.eye(4)
    pose1[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    pose1[:3, 3] = [0.03, -0.01, 0.02]
    d0 = torch.from_numpy(synth_depth(H, W, K, pose0))
    d1 = torch.from_numpy(synth_depth(H, W, K, pose1))
    Kt = torch.from_numpy(K).float()
    out = {"depth0": d0.numpy(), "depth1": d1.numpy(), "K": K.astype(np.float32)}
    builder = ref_icp.ImagePyramids([2, 1, 0], "max")
    vp0 = ref_utils.build_vertex_pyramid(d0.view(H, W, 1), builder, Kt.clone())
    vp1 = ref_utils.build_vertex_pyramid(d1.view(H, W, 1), builder, Kt.clone())
    np0, np1 = ref_utils.build_normal_pyramid(vp0), ref_utils.build_normal_pyramid(vp1)
    for lvl in range(3):
        out["vertex0_l%d" % lvl], out["normal0_l%d" % lvl] = vp0[lvl].numpy(), np0[lvl].numpy()
        out["vertex1_l%d" % lvl], out["normal1_l%d" % lvl] = vp1[lvl].numpy(), np1[lvl].numpy()
    downscales, iters = [0.25, 0.5, 1.0], [5, 5, 5]
    pose = torch.eye(4)
    for lvl in range(3):
        icp = ref_icp.ICP(iters[lvl], damping=1e-4, distance_threshold=0.1, normal_threshold=20)
        Kd = Kt * downscales[lvl]
        Kd[2, 2] = 1.0
        if lvl == 0:
            mask0 = vp1[lvl][..., -1] > 0.0
            res, J, valid = icp.compute_residuals_jacobian(vp1[lvl], vp0[lvl], np1[lvl], np0[lvl], mask0, pose, Kd,
                                                           icp.distance_threshold, icp.normal_threshold)
            out["first_jtj"], out["first_jtr"] = icp.compute_jtj(J).numpy(), icp.compute_jtr(J, res).numpy()
            out["first_valid"] = np.array(int(valid.sum()))
        # predict_pose passes (vertex_t1, vertex_t0, normal_t1, normal_t0): frame 1 is the template (icp.py:436-439)
        pose, ratio = icp.icp(pose, vp1[lvl], vp0[lvl], np1[lvl], np0[lvl], Kd)
        out["pose_after_l%d" % lvl], out["ratio_l%d" % lvl] = pose.numpy().copy(), np.array(float(ratio))
    out["true_pose10"] = (np.linalg.inv(pose0) @ pose1).astype(np.float32)   # x0 = T x1: what predict_pose estimates
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)
    print("estimated\n", out["pose_after_l2"], "\ntrue\n", out["true_pose10"])


if __name__ == "__main__":
    main()
