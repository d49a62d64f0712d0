"""torch restatement (float32) of the tracker front-end -- TEST INFRASTRUCTURE ONLY (tests/, bench baselines).

Follows SLAM/utils.py:65-125,542-559 (vertex / normal maps, pyramids) and SLAM/icp.py:33-121,128-145,230-335 (projective
point-to-plane ICP) line by line, with a device argument instead of the hard-coded CUDA sentinels.  Pinned:
tests/test_icp_oracle.py checks it against tests/golden/icp.npz, which was produced by importing and running the
reference's own modules (tests/golden/make_icp_golden.py)."""
import math

import torch
import torch.nn.functional as F


def compute_vertex_map(depth, K):  # utils.py:65-75
    H, W = depth.shape[:2]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t().to(depth.device), j.t().to(depth.device)
    return torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], -1) * depth


def feature_gradient(img):  # utils.py:77-100 with normalize_gradient=False
    H, W, C = img.shape
    wx = torch.tensor([[-1.0, 0, 1], [-2, 0, 2], [-1, 0, 1]]).view(1, 1, 3, 3).to(img)
    wy = torch.tensor([[-1.0, -2, -1], [0, 0, 0], [1, 2, 1]]).view(1, 1, 3, 3).to(img)
    pad = F.pad(img.permute(2, 0, 1).reshape(-1, 1, H, W), (1, 1, 1, 1), mode="replicate")
    dx = F.conv2d(pad, wx).squeeze().permute(1, 2, 0)
    dy = F.conv2d(pad, wy).squeeze().permute(1, 2, 0)
    return dx, dy


def compute_normal_map(vertex_map):  # utils.py:102-125
    H, W, C = vertex_map.shape
    dx, dy = feature_gradient(vertex_map)
    normal = torch.linalg.cross(dy.reshape(-1, 3), dx.reshape(-1, 3)).view(H, W, 3)
    normal = normal / (torch.norm(normal, p=2, dim=-1, keepdim=True) + 1e-8)
    depth = vertex_map[:, :, -1]
    invalid = (depth <= depth.min()) | (depth >= depth.max())
    return torch.where(invalid[..., None], torch.zeros_like(normal), normal)


def depth_pyramid(depth, scales):  # icp.py:342-360, pool='max'
    H, W = depth.shape[:2]
    x = depth.reshape(1, 1, H, W)
    return [F.max_pool2d(x, 1 << i, 1 << i) for i in scales]


def build_vertex_pyramid(depth, scales, K):  # utils.py:542-553
    pyr = depth_pyramid(depth, scales)
    out = []
    for i, d in enumerate(pyr):
        Hs, Ws = d.shape[2:4]
        Kd = K * (1 / 2 ** (len(pyr) - i - 1))
        Kd[2, 2] = 1.0
        out.append(compute_vertex_map(d.reshape(Hs, Ws, 1), Kd))
    return out


def warp_features(Feat, u, v):  # icp.py:128-145
    H, W, C = Feat.shape
    grid = torch.cat(((u / ((W - 1) / 2) - 1).view(1, H, W, 1), (v / ((H - 1) / 2) - 1).view(1, H, W, 1)), dim=-1)
    out = F.grid_sample(Feat.unsqueeze(0).permute(0, 3, 1, 2), grid, mode="nearest", padding_mode="border", align_corners=True)
    return out.squeeze(0).permute(1, 2, 0)


def residuals_jacobian(v0, v1, n0, n1, pose10, K, dist_thr, normal_thr):  # icp.py:52-104
    R, t = pose10[:3, :3], pose10[:3, 3]
    H, W, _ = v0.shape
    q = (R @ v0.view(-1, 3).T).T.view(H, W, 3) + t[None, None, :]
    m = (R @ n0.view(-1, 3).T).T.view(H, W, 3)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u = (q[..., 0] / q[..., 2]) * fx + cx
    v = (q[..., 1] / q[..., 2]) * fy + cy
    inview = (u > 0) & (u < W - 1) & (v > 0) & (v < H - 1)
    r_v1, r_n1 = warp_features(v1, u, v), warp_features(n1, u, v)
    diff = q - r_v1
    res = (r_n1 * diff).sum(-1)
    J = torch.cat((torch.linalg.cross(q.view(-1, 3), r_n1.view(-1, 3)), r_n1.view(-1, 3)), dim=-1).view(H, W, 6)
    invalid = ~inview | (diff.norm(p=2, dim=-1) > dist_thr) | ~(v0[..., -1] > 0) | ~(r_v1[..., -1] > 0) | \
        ~((m * r_n1).sum(-1) > normal_thr)
    J = torch.where(invalid[..., None], torch.zeros_like(J), J)
    res = torch.where(invalid, torch.zeros_like(res), res)
    return res.view(-1), J.view(-1, 6), ~invalid


def exp_se3(xi):  # icp.py:272-310
    w, v = xi[:3], xi[3:6]
    What = torch.tensor([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]]).to(xi)
    W2 = What @ What
    th = torch.norm(w)
    eye = torch.eye(3).to(xi)
    if th <= 1e-8:
        E, Jm = eye, eye
    else:
        E = eye + What * torch.sin(th) / th + W2 * (1.0 - torch.cos(th)) / th ** 2
        Jm = eye + (1 - torch.cos(th)) / th ** 2 * What + (th - torch.sin(th)) / th ** 3 * W2
    T = torch.eye(4).to(xi)
    T[:3, :3] = E
    T[:3, 3] = Jm @ v
    return T


def icp(pose10, v0, v1, n0, n1, K, iters, damping=1e-6, distance_threshold=0.2, normal_threshold=20):  # icp.py:33-48
    thr = math.cos(math.radians(normal_threshold))
    valid = None
    for _ in range(iters):
        res, J, valid = residuals_jacobian(v0, v1, n0, n1, pose10, K, distance_threshold, thr)
        JtJ = J.T @ J                                      # icp.py:106-111
        JtR = J.T @ res                                    # icp.py:113-121
        Hm = JtJ + torch.trace(JtJ) * damping * torch.eye(6).to(JtJ)   # icp.py:248-256
        xi = -(torch.inverse(Hm.cpu()).to(Hm) @ JtR)       # icp.py:313-335 (the reference inverts on the host as well)
        pose10 = exp_se3(xi) @ pose10
    H, W = v0.shape[:2]
    return pose10, valid.sum() / H / W


def predict_pose(depth0, depth1, K, downscales=(0.25, 0.5, 1.0), iters=(5, 5, 5), distance_threshold=0.1,
                 normal_threshold=20, damping=1e-4):  # icp.py:424-441
    scales = list(range(len(downscales) - 1, -1, -1))
    vp0, vp1 = build_vertex_pyramid(depth0, scales, K.clone()), build_vertex_pyramid(depth1, scales, K.clone())
    np0, np1 = [compute_normal_map(v) for v in vp0], [compute_normal_map(v) for v in vp1]
    pose = torch.eye(4, device=depth0.device)
    ratio = None
    for lvl, s in enumerate(downscales):
        Kd = K * s
        Kd[2, 2] = 1.0
        pose, ratio = icp(pose, vp1[lvl], vp0[lvl], np1[lvl], np0[lvl], Kd, iters[lvl], damping, distance_threshold,
                          normal_threshold)
    return pose, ratio
