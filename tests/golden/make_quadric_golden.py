#!/usr/bin/env python
"""Generates tests/golden/quadric.npz by importing the reference's own SLAM/multiprocess/quadrics.py
(unmodified, from /root/reference) in THIS container and running its classes on seeded inputs:

  * Object.__init__                (quadrics.py:429-487)   -> init_axes / init_R / init_center
  * Ellipsoid.project + ComputeBbox (quadrics.py:388-425, 148-248) -> proj_bbox / proj_ellipse
  * Object_Optimize_only            (quadrics.py:2234-2298) -> refined axes / R / center, with Python's
    `random` seeded so the per-iteration view schedule is reproducible (stored as `view_choice`).

The module hard-codes device="cuda" and imports plotting packages that are not installed here; the script
substitutes empty stand-in modules for the plotting imports and routes device="cuda" tensor factories to the CPU.
No reference source is modified or copied.  Run:  python tests/golden/make_quadric_golden.py
"""
import os
import random
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = os.environ.get("DQO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quadric.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda k: mock.MagicMock()
    sys.modules[name] = m
    return m


def load_reference_quadrics():
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d", "PIL",
              "plyfile", "cv2"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = mock.MagicMock()
    sys.modules["PIL"].Image = mock.MagicMock()
    sys.modules["PIL"].ImageDraw = mock.MagicMock()
    sys.modules["plyfile"].PlyData = mock.MagicMock()
    sys.modules["plyfile"].PlyElement = mock.MagicMock()

    def cpuify(fn):
        def wrapped(*a, **k):
            if k.get("device", None) == "cuda":
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    torch.tensor = cpuify(torch.tensor)
    torch.eye = cpuify(torch.eye)
    torch.zeros = cpuify(torch.zeros)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_quadrics", os.path.join(REF, "SLAM/multiprocess/quadrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(n=6, views=4, seed=3):
    rng = np.random.RandomState(seed)
    K = np.array([[600.0, 0, 599.5], [0, 600.0, 339.5], [0, 0, 1.0]])
    objs = []
    for i in range(n):
        center = np.array([rng.uniform(-1, 1), rng.uniform(-0.5, 0.5), rng.uniform(2.5, 4.5)])
        axes = rng.uniform(0.1, 0.45, 3)
        Rts, bbs = [], []
        for v in range(views):
            ang = rng.uniform(-0.25, 0.25)
            ca, sa = np.cos(ang), np.sin(ang)
            Rcw = np.array([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]])
            tcw = rng.uniform(-0.2, 0.2, 3)
            Rt = np.concatenate([Rcw, tcw[:, None]], axis=1)
            Rts.append(Rt)
        objs.append((center, axes, Rts))
    return K, objs


def main():
    q = load_reference_quadrics()
    K, objs = make_inputs()
    n, V = len(objs), len(objs[0][2])
    iters = 20
    out = {"K": K}
    init_axes, init_R, init_center, bbs0, dstats, Rt0 = [], [], [], [], [], []
    proj_bbox, proj_ell = [], []
    obs = np.zeros((n, V, 4)); Ps = np.zeros((n, V, 3, 4))
    ref_axes, ref_R, ref_center, choices = [], [], [], []
    rng = np.random.RandomState(11)
    for i, (center, axes, Rts) in enumerate(objs):
        gt = q.Ellipsoid(axes, np.eye(3), center)
        bbs = []
        for v, Rt in enumerate(Rts):
            P = K @ Rt
            ell = gt.project(P)
            bb = ell.ComputeBbox() + rng.uniform(-2, 2, 4)  # noisy detections
            bbs.append(bb)
            obs[i, v] = bb
            Ps[i, v] = P
        # single-view construction from the first detection
        depth_c = (Rts[0][:3, :3] @ center + Rts[0][:3, 3])[2]
        dstat = [depth_c, 0.15]
        obj = q.Object(cat=i, bb=bbs[0], ell=None, score=1.0, depth_data=dstat, K=K, Rt=Rts[0], frame_idx=0, kf=False)
        init_axes.append(obj.ellipsoid_.axes_.copy()); init_R.append(obj.ellipsoid_.R_.copy())
        init_center.append(obj.ellipsoid_.center_.copy())
        bbs0.append(bbs[0]); dstats.append(dstat); Rt0.append(Rts[0])
        # projection of the constructed ellipsoid into the last view
        e2 = obj.ellipsoid_.project(K @ Rts[-1])
        proj_bbox.append(e2.ComputeBbox().copy())
        proj_ell.append([e2.GetAxes()[0], e2.GetAxes()[1], e2.GetAngle(), e2.GetCenter()[0], e2.GetCenter()[1]])
        # refinement through the reference's own function; record the view schedule it will draw
        obj.bboxes_ = [np.asarray(b, dtype=np.float64) for b in bbs]
        obj.Rts_ = list(Rts)
        random.seed(100 + i)
        sched = []
        for it in range(iters):
            k = random.randint(0, len(obj.bboxes_) - 1)
            if it > iters / 4:
                k = -1
            sched.append(k)
        choices.append(sched)
        random.seed(100 + i)
        det = {"obj": obj, "is_validate": True, "node_id": i}
        Map_global = {i: obj}
        q.Object_Optimize_only([det], Map_global, K, Rts[-1])
        ref_axes.append(np.asarray(obj.ellipsoid_.axes_, dtype=np.float64))
        ref_R.append(np.asarray(obj.ellipsoid_.R_, dtype=np.float64))
        ref_center.append(np.asarray(obj.ellipsoid_.center_, dtype=np.float64))
    out.update(init_bbox=np.array(bbs0), init_depth_stats=np.array(dstats), init_Rt=np.array(Rt0),
               init_axes=np.array(init_axes), init_R=np.array(init_R), init_center=np.array(init_center),
               proj_P=np.array([K @ o[2][-1] for o in objs]), proj_bbox=np.array(proj_bbox), proj_ellipse=np.array(proj_ell),
               obs_bboxes=obs, Ps=Ps, view_choice=np.array(choices, dtype=np.int32),
               refined_axes=np.array(ref_axes), refined_R=np.array(ref_R), refined_center=np.array(ref_center))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)
    print("init axes[0]", out["init_axes"][0], "refined axes[0]", out["refined_axes"][0])


if __name__ == "__main__":
    main()
