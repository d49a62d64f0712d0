"""Batched dual-quadric operators (reference SLAM/multiprocess/quadrics.py) over the C-ABI.

  quadric_init     <- Object.__init__          quadrics.py:451-487 (single-view construction, fp64)
  quadric_project  <- Ellipsoid.project + Ellipse.ComputeBbox   quadrics.py:388-425, 148-248 (fp64)
  quadric_refine   <- Object_Optimize_only     quadrics.py:2234-2298 (20-iteration IoU-Adam, fp32)
NOTE: the reference contains no SVD / least-squares quadric fit; nothing of the sort is invented here.
"""
import random

import torch

from ._lib import check, lib, ptr


def _s():
    return torch.cuda.current_stream().cuda_stream


def _f64(t, dev):
    return torch.as_tensor(t, dtype=torch.float64, device=dev).contiguous()


def quadric_init(bboxes, depth_stats, K, Rts, device="cuda"):
    """bboxes [n,4] (x0,y0,x1,y1), depth_stats [n,2] (avg_depth, diff_depth), K [3,3], Rts [n,3,4] ->
    (axes [n,3], R [n,3,3], center [n,3]) float64."""
    dev = torch.device(device)
    bb, ds, Kt, Rt = _f64(bboxes, dev), _f64(depth_stats, dev), _f64(K, dev), _f64(Rts, dev)
    n = bb.shape[0]
    axes = torch.empty((n, 3), dtype=torch.float64, device=dev)
    R = torch.empty((n, 3, 3), dtype=torch.float64, device=dev)
    center = torch.empty((n, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().dqo_quadric_init(n, ptr(bb), ptr(ds), ptr(Kt), ptr(Rt), ptr(axes), ptr(R), ptr(center), _s()),
              "dqo_quadric_init")
    return axes, R, center


def quadric_project(axes, R, center, P):
    """Projects n ellipsoids with n 3x4 matrices P = K [R|t]; returns (bbox [n,4], ellipse [n,5] = ax0, ax1, angle, cx, cy)."""
    dev = axes.device
    ax, Rm, c, Pm = _f64(axes, dev), _f64(R, dev), _f64(center, dev), _f64(P, dev)
    n = ax.shape[0]
    bbox = torch.empty((n, 4), dtype=torch.float64, device=dev)
    ell = torch.empty((n, 5), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().dqo_quadric_project(n, ptr(ax), ptr(Rm), ptr(c), ptr(Pm), ptr(bbox), ptr(ell), _s()),
              "dqo_quadric_project")
    return bbox, ell


def reference_view_schedule(n_views, iters=20, rng=random):
    """The per-iteration observation choice of Object_Optimize_only (quadrics.py:2264-2266): uniform for
    iter <= iters/4, afterwards always the latest (-1)."""
    sched = []
    for it in range(iters):
        k = rng.randint(0, n_views - 1)
        if it > iters / 4:
            k = -1
        sched.append(k)
    return sched


def quadric_refine(axes, R, center, obs_bboxes, Ps, n_views, view_choice, iters=20, lr_axes=0.01, lr_center=0.001,
                   lr_R=0.01):
    """Refines n ellipsoids in one launch.  axes [n,3], R [n,3,3], center [n,3] float32 (returned updated, inputs
    untouched); obs_bboxes [n,V,4], Ps [n,V,3,4], n_views [n] int32, view_choice [n,iters] int32 (negative = from the end).
    Returns (axes, R, center, last_loss)."""
    dev = obs_bboxes.device
    f = dict(dtype=torch.float32, device=dev)
    ax = torch.as_tensor(axes, **f).clone().contiguous()
    Rm = torch.as_tensor(R, **f).clone().contiguous()
    c = torch.as_tensor(center, **f).clone().contiguous()
    ob = torch.as_tensor(obs_bboxes, **f).contiguous()
    Pm = torch.as_tensor(Ps, **f).contiguous()
    nv = torch.as_tensor(n_views, dtype=torch.int32, device=dev).contiguous()
    vc = torch.as_tensor(view_choice, dtype=torch.int32, device=dev).contiguous()
    n, V = ob.shape[0], ob.shape[1]
    if vc.shape != (n, iters):
        raise ValueError("view_choice must be [n, iters]")
    last = torch.empty((n,), **f)
    with torch.cuda.device(dev):
        check(lib().dqo_quadric_refine(n, iters, V, ptr(nv), ptr(ob), ptr(Pm), ptr(vc), lr_axes, lr_center, lr_R,
                                       ptr(ax), ptr(Rm), ptr(c), ptr(last), _s()), "dqo_quadric_refine")
    return ax, Rm, c, last
