"""CPU restatement of the reference's dual-quadric math (SLAM/multiprocess/quadrics.py) — TEST INFRASTRUCTURE ONLY.

Pinned against tests/golden/quadric.npz, produced by tests/golden/make_quadric_golden.py which imports and runs the
reference's own module (Object.__init__, Ellipsoid.project/Ellipse.ComputeBbox, Object_Optimize_only).
"""
import numpy as np
import torch


def object_init(bb, depth_data, K, Rt):
    """Object.__init__, quadrics.py:451-487.  Returns (axes, R_world, center_world) as float64 numpy."""
    avg_depth, diff_depth = depth_data
    bb_center = np.array([(bb[0] + bb[2]) / 2, (bb[1] + bb[3]) / 2])
    u = (bb_center[0] - K[0, 2]) / K[0, 0]
    v = (bb_center[1] - K[1, 2]) / K[1, 1]
    c_cam = np.array([u * avg_depth, v * avg_depth, avg_depth])
    Rcw, tcw = Rt[:3, :3], Rt[:3, 3]
    center_world = Rcw.T @ c_cam + (-Rcw.T @ tcw)
    zc = c_cam / np.linalg.norm(c_cam)
    xc = np.cross(-np.array([0, -1, 0]), zc)
    xc = xc / np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    rot_cam = np.stack([xc, yc, zc], axis=1)
    rot_world = Rcw.T @ rot_cam
    w, h = bb[2] - bb[0], bb[3] - bb[1]
    axes = np.array([w * avg_depth / K[0, 0] * 0.5, h * avg_depth / K[1, 1] * 0.5, diff_depth * 0.5])
    return axes, rot_world, center_world


def dual_quadric(axes, R, center):
    """Ellipsoid.__init__, quadrics.py:388-403."""
    Q = np.diag([axes[0] ** 2, axes[1] ** 2, axes[2] ** 2, -1.0])
    T = np.eye(4)
    T[:3, 3] = center
    Rw = np.eye(4)
    Rw[:3, :3] = R
    tr = T @ Rw
    Q = tr @ Q @ tr.T
    Q = 0.5 * (Q + Q.T)
    return Q / -Q[3, 3]


def project_bbox(axes, R, center, P):
    """Ellipsoid.project + Ellipse.decompose/ComputeBbox, quadrics.py:404-406, 178-230.
    Returns (bbox[4], (ax0, ax1, angle, cx, cy))."""
    C = P @ dual_quadric(axes, R, center) @ P.T
    C = 0.5 * (C + C.T)
    C = C / -C[2, 2]
    center2 = -C[:2, 2]
    Tc = np.eye(3)
    Tc[:2, 2] = -center2
    tmp = Tc @ C @ Tc.T
    Cc = 0.5 * (tmp + tmp.T)
    vals, vecs = np.linalg.eigh(Cc[:2, :2])
    if np.linalg.det(vecs) < 0:
        vecs[:, 1] *= -1
    if vecs[0, 0] < 0:
        vecs *= -1
    ax = np.sqrt(np.abs(vals))
    ang = np.arctan2(vecs[1, 0], vecs[0, 0])
    c, s = np.cos(ang), np.sin(ang)
    xmax = np.sqrt(ax[0] ** 2 * c ** 2 + ax[1] ** 2 * s ** 2)
    ymax = np.sqrt(ax[0] ** 2 * s ** 2 + ax[1] ** 2 * c ** 2)
    bbox = np.array([center2[0] - xmax, center2[1] - ymax, center2[0] + xmax, center2[1] + ymax])
    return bbox, (ax[0], ax[1], ang, center2[0], center2[1])


def _bbox_tensor(axes, R, center, P):
    """Ellipsoid_tensor.forward + Ellipse_tensor, quadrics.py:2178-2225, 2018-2091 (device parameterised: CPU)."""
    Qd = torch.diag(torch.cat([axes[:3] ** 2, torch.tensor([-1.0])]))
    T = torch.eye(4)
    T = T.clone()
    T[:3, 3] = center
    Rw = torch.eye(4).clone()
    Rw[:3, :3] = R
    tr = T @ Rw
    Q = tr @ Qd @ tr.T
    Q = 0.5 * (Q + Q.T)
    Q = Q / -Q[3, 3]
    C = P @ Q @ P.T
    C = 0.5 * (C + C.T)
    C = C / -C[2, 2]
    c2 = -C[:2, 2]
    Tc = torch.eye(3).clone()
    Tc[:2, 2] = -c2
    tmp = Tc @ C @ Tc.T
    Cc = 0.5 * (tmp + tmp.T)
    vals, vecs = torch.linalg.eig(Cc[:2, :2])
    vals, vecs = vals.real, vecs.real
    if torch.det(vecs) < 0:
        vecs = vecs.clone()
        vecs[:, 1] *= -1
    if vecs[0, 0] < 0:
        vecs = vecs.clone()
        vecs *= -1
    ax = torch.sqrt(torch.abs(vals))
    ang = torch.atan2(vecs[1, 0], vecs[0, 0])
    c, s = torch.cos(ang), torch.sin(ang)
    xmax = torch.sqrt(ax[0] ** 2 * c ** 2 + ax[1] ** 2 * s ** 2)
    ymax = torch.sqrt(ax[0] ** 2 * s ** 2 + ax[1] ** 2 * c ** 2)
    return torch.stack([c2[0] - xmax, c2[1] - ymax, c2[0] + xmax, c2[1] + ymax])


def _iou(bb1, bb2):
    """bboxes_iou with python min/max on (float, tensor), quadrics.py:283-290."""
    inter_w = max(min(bb1[2], bb2[2]) - max(bb1[0], bb2[0]), 0)
    inter_h = max(min(bb1[3], bb2[3]) - max(bb1[1], bb2[1]), 0)
    area_inter = inter_w * inter_h
    a1 = (bb1[2] - bb1[0]) * (bb1[3] - bb1[1])
    a2 = (bb2[2] - bb2[0]) * (bb2[3] - bb2[1])
    return area_inter / (a1 + a2 - area_inter)


def refine(axes, R, center, obs_bboxes, Ps, view_choice, iters=20, lrs=(0.01, 0.001, 0.01)):
    """Object_Optimize_only's per-object loop, quadrics.py:2245-2295, with the view schedule given explicitly."""
    ax = torch.nn.Parameter(torch.tensor(np.asarray(axes), dtype=torch.float32))
    Rm = torch.nn.Parameter(torch.tensor(np.asarray(R), dtype=torch.float32))
    c = torch.nn.Parameter(torch.tensor(np.asarray(center), dtype=torch.float32))
    opt = torch.optim.Adam([{"params": [ax], "lr": lrs[0]}, {"params": [c], "lr": lrs[1]}, {"params": [Rm], "lr": lrs[2]}],
                           eps=1e-15)
    last = 0.0
    for it in range(iters):
        opt.zero_grad()
        k = int(view_choice[it])
        obs = [float(x) for x in obs_bboxes[k]]
        P = torch.tensor(np.asarray(Ps[k]), dtype=torch.float32)
        bb = _bbox_tensor(ax, Rm, c, P)
        loss = 1.0 - _iou(obs, bb)
        last = float(loss)
        if loss == 1:
            continue
        loss.backward()
        opt.step()
    return ax.detach().numpy(), Rm.detach().numpy(), c.detach().numpy(), last
