"""Host-side mirror of the tracker front-end (reference SLAM/icp.py and SLAM/utils.py:65-125,542-559) over the C-ABI.

Same names and argument meaning as the reference: ImagePyramids, compute_vertex_map, compute_normal_map,
build_vertex_pyramid, build_normal_pyramid, ICP (with .icp()) and the coarse-to-fine loop of IcpTracker.predict_pose as
`predict_pose`.  One ICP iteration is two kernel launches and the pose stays on the device; the reference runs ~40 torch
kernels per iteration and inverts the 6x6 system on the host (icp.py:313-326)."""
import math

import torch

from ._lib import check, lib, ptr


def _s():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s expects CUDA tensors (there is no CPU path)" % name)


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)


class ImagePyramids:
    """ImagePyramids(scales, pool='max') of SLAM/icp.py:342-360 for depth images: level i is MaxPool2d(1 << i)."""

    def __init__(self, scales, pool="max"):
        if pool != "max":
            raise NotImplementedError("only the max pyramid of the tracker (icp.py:383) is provided")
        self.scales = list(scales)

    def __call__(self, x):
        _need_cuda(x, "ImagePyramids")
        H, W = x.shape[-2:]
        d = x.reshape(H, W).contiguous().float()
        outs = []
        for lvl in self.scales:
            o = torch.empty((1, 1, H >> lvl, W >> lvl), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                check(lib().dqo_depth_maxpool(W, H, int(lvl), ptr(d), ptr(o), _s()), "dqo_depth_maxpool")
            outs.append(o)
        return outs


def _intrinsics(K):
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


def vertex_normal_map(depth, K, want_normal=True):
    """Fused compute_vertex_map + compute_normal_map (SLAM/utils.py:65-125): (vertex [H,W,3], normal [H,W,3] or None)."""
    _need_cuda(depth, "vertex_normal_map")
    H, W = depth.shape[:2]
    d = depth.reshape(H, W).contiguous().float()
    fx, fy, cx, cy = _intrinsics(K)
    vertex = torch.empty((H, W, 3), dtype=torch.float32, device=depth.device)
    normal = torch.empty((H, W, 3), dtype=torch.float32, device=depth.device) if want_normal else None
    ws = _ws(lib().dqo_vertex_normal_workspace_bytes(), depth.device)
    with torch.cuda.device(depth.device):
        check(lib().dqo_vertex_normal_map(W, H, ptr(d), fx, fy, cx, cy, ptr(vertex), ptr(normal), ptr(ws), _s()),
              "dqo_vertex_normal_map")
    return vertex, normal


def compute_vertex_map(depth, K):
    return vertex_normal_map(depth, K, want_normal=False)[0]


def compute_normal_map(vertex_map):
    _need_cuda(vertex_map, "compute_normal_map")
    H, W = vertex_map.shape[:2]
    v = vertex_map.contiguous().float()
    normal = torch.empty((H, W, 3), dtype=torch.float32, device=v.device)
    ws = _ws(lib().dqo_vertex_normal_workspace_bytes(), v.device)
    with torch.cuda.device(v.device):
        check(lib().dqo_normal_map(W, H, ptr(v), ptr(normal), ptr(ws), _s()), "dqo_normal_map")
    return normal


def build_vertex_pyramid(depth, pyramid_builder, K):
    """SLAM/utils.py:542-553: vertex maps of the depth pyramid, coarsest first, with K scaled per level."""
    H, W = depth.shape[:2]
    depth_pyramid = pyramid_builder(depth.reshape(1, 1, H, W))
    out = []
    for i, d in enumerate(depth_pyramid):
        Hs, Ws = d.shape[2:4]
        scale = 1.0 / 2 ** (len(depth_pyramid) - i - 1)
        Kd = K.clone().float() * scale
        Kd[2, 2] = 1.0
        out.append(compute_vertex_map(d.reshape(Hs, Ws, 1), Kd))
    return out


def build_normal_pyramid(vertex_pyramid):
    return [compute_normal_map(v) for v in vertex_pyramid]


class ICP:
    """Projective point-to-plane ICP (SLAM/icp.py:16-130): same constructor arguments, `icp()` returns (pose10, valid_ratio)
    as device tensors."""

    def __init__(self, max_iter=3, damping=1e-6, distance_threshold=0.2, normal_threshold=20, verbose=False):
        self.max_iterations = int(max_iter)
        self.distance_threshold = float(distance_threshold)
        self.normal_threshold = math.cos(math.radians(normal_threshold))
        self.damping = float(damping)
        self.verbose = verbose

    def icp(self, pose10, vertex_t0, vertex_t1, normal_t0, normal_t1, K):
        _need_cuda(vertex_t0, "ICP.icp")
        H, W = vertex_t0.shape[:2]
        dev = vertex_t0.device
        pose = torch.as_tensor(pose10, dtype=torch.float32, device=dev).clone().contiguous()
        ratio = torch.zeros(1, dtype=torch.float32, device=dev)
        fx, fy, cx, cy = _intrinsics(K)
        ws = _ws(lib().dqo_icp_workspace_bytes(), dev)
        args = [t.contiguous().float() for t in (vertex_t0, vertex_t1, normal_t0, normal_t1)]
        with torch.cuda.device(dev):
            check(lib().dqo_icp_level(W, H, ptr(args[0]), ptr(args[1]), ptr(args[2]), ptr(args[3]), fx, fy, cx, cy,
                                      self.distance_threshold, self.normal_threshold, self.damping, self.max_iterations,
                                      ptr(pose), ptr(ratio), ptr(ws), _s()), "dqo_icp_level")
        return pose, ratio[0]


def predict_pose(depth_t0, depth_t1, K, downscales=(0.25, 0.5, 1.0), iters=(5, 5, 5), distance_threshold=0.1,
                 normal_threshold=20, damping=1e-4):
    """The coarse-to-fine loop of IcpTracker.predict_pose (SLAM/icp.py:424-441): pose_t1_t0 from two depth images.
    Frame 1 is the template of every level, exactly as the reference calls ICP.icp(pose, v_t1, v_t0, n_t1, n_t0, K)."""
    builder = ImagePyramids(list(range(len(downscales) - 1, -1, -1)), "max")
    vp0, vp1 = build_vertex_pyramid(depth_t0, builder, K), build_vertex_pyramid(depth_t1, builder, K)
    np0, np1 = build_normal_pyramid(vp0), build_normal_pyramid(vp1)
    pose = torch.eye(4, dtype=torch.float32, device=depth_t0.device)
    ratio = None
    for lvl, scale in enumerate(downscales):
        Kd = K.clone().float() * scale
        Kd[2, 2] = 1.0
        pose, ratio = ICP(iters[lvl], damping, distance_threshold, normal_threshold).icp(pose, vp1[lvl], vp0[lvl], np1[lvl],
                                                                                         np0[lvl], Kd)
    return pose, ratio
