"""Deterministic synthetic scenes of Replica / Cube-Diorama shape.

Camera conventions follow the reference exactly: `world_view_transform` and `full_proj_transform` are the
TRANSPOSED (column-major) W2C and W2C·Proj float32 matrices (scene/cameras.py:138-154, utils/graphics_utils.py:52-86).
Everything is generated on the CPU with torch.Generator(seed) and is independent of the device.
"""
import math

import numpy as np
import torch

C0 = 0.28209479177387814


def RGB2SH(rgb):  # utils/sh_utils.py
    return (rgb - 0.5) / C0


def get_world2view2(R, t, translate=np.array([0.0, 0.0, 0.0]), scale=1.0):
    """utils/graphics_utils.py:52-64 (R is stored transposed, like the reference's Camera.R)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    cam_center = C2W[:3, 3]
    cam_center = (cam_center + translate) * scale
    C2W[:3, 3] = cam_center
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def get_projection_matrix(znear, zfar, fovX, fovY):
    """utils/graphics_utils.py:67-86."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class SynthCamera:
    """The subset of scene.cameras.Camera that SLAM/render.py reads (render.py:140-162)."""

    def __init__(self, W, H, fx, fy, cx, cy, R=None, T=None, znear=0.01, zfar=100.0):
        self.image_width, self.image_height = W, H
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        self.FoVx = 2 * math.atan(W / (2 * fx))
        self.FoVy = 2 * math.atan(H / (2 * fy))
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
        self.T = np.zeros(3) if T is None else np.asarray(T, dtype=np.float64)
        self.world_view_transform = torch.tensor(get_world2view2(self.R, self.T)).transpose(0, 1).contiguous()
        self.projection_matrix = get_projection_matrix(znear, zfar, self.FoVx, self.FoVy).transpose(0, 1)
        self.full_proj_transform = (
            self.world_view_transform.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0))).squeeze(0).contiguous()
        self.camera_center = self.world_view_transform.inverse()[3, :3].contiguous()

    def to(self, device):
        self.world_view_transform = self.world_view_transform.to(device)
        self.full_proj_transform = self.full_proj_transform.to(device)
        self.camera_center = self.camera_center.to(device)
        return self

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)


CONFIGS = {
    # name: (P, W, H, fx, fy, cx, cy, sh_degree)
    "tiny": (2000, 160, 96, 120.0, 120.0, 79.5, 47.5, 0),
    "small": (20000, 320, 240, 260.0, 260.0, 159.5, 119.5, 3),
    "deg1": (12000, 320, 240, 260.0, 260.0, 159.5, 119.5, 1),
    "ragged": (15000, 333, 187, 270.0, 270.0, 166.0, 93.0, 2),   # neither dimension a multiple of the 16-pixel tile
    "huge": (30000, 4112, 4100, 2000.0, 2000.0, 2055.5, 2049.5, 0),   # 257 x 257 = 66049 tiles: 32-bit tile keys
    "c1": (100000, 640, 480, 525.0, 525.0, 319.5, 239.5, 0),
    "c2": (1000000, 1200, 680, 600.0, 600.0, 599.5, 339.5, 3),
    "c5": (3000000, 1920, 1080, 960.0, 960.0, 959.5, 539.5, 3),
}


def make_camera(name, pose_index=0):
    P, W, H, fx, fy, cx, cy, deg = CONFIGS[name]
    # small deterministic rotation / translation so that no matrix entry is trivially 0 or 1
    ang = 0.05 + 0.02 * pose_index
    ca, sa = math.cos(ang), math.sin(ang)
    Ry = np.array([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]])
    ang2 = -0.03 + 0.01 * pose_index
    cb, sb = math.cos(ang2), math.sin(ang2)
    Rx = np.array([[1, 0, 0], [0, cb, -sb], [0, sb, cb]])
    R = Ry @ Rx
    T = np.array([0.03 + 0.01 * pose_index, -0.02, 0.05])
    return SynthCamera(W, H, fx, fy, cx, cy, R=R, T=T)


def make_gaussians(name, seed=2024, P=None, sh_degree=None, opaque_fraction=0.7):
    """Surfel-like Gaussians in the camera frustum (10 % outside), already ACTIVATED like the tensors the
    reference hands to the rasterizer (post-exp scales, post-sigmoid opacity, unit quaternions)."""
    Pn, W, H, fx, fy, cx, cy, deg = CONFIGS[name]
    P = Pn if P is None else P
    deg = deg if sh_degree is None else sh_degree
    g = torch.Generator(device="cpu").manual_seed(seed)
    cam = make_camera(name)
    z = 0.5 + 4.5 * torch.rand(P, generator=g)
    # 10 % of the points leave the frustum (|ndc| up to 1.6) to exercise culling
    spread = torch.where(torch.rand(P, generator=g) < 0.1, torch.tensor(1.6), torch.tensor(1.0))
    u = (torch.rand(P, generator=g) * 2 - 1) * spread
    v = (torch.rand(P, generator=g) * 2 - 1) * spread
    x_c = u * z * (W / (2 * fx))
    y_c = v * z * (H / (2 * fy))
    pts_c = torch.stack([x_c, y_c, z], dim=1).double()
    # camera -> world with the reference's convention: p_c = W2C p_w, W2C = world_view_transform^T
    W2C = cam.world_view_transform.transpose(0, 1).double()
    C2W = torch.linalg.inv(W2C)
    xyz = (pts_c @ C2W[:3, :3].T + C2W[:3, 3]).float()
    s = torch.exp(math.log(0.002) + (math.log(0.05) - math.log(0.002)) * torch.rand(P, 2, generator=g))
    scales = torch.cat([s, 0.1 * s.min(dim=1, keepdim=True).values], dim=1)
    perm = torch.argsort(torch.rand(P, 3, generator=g), dim=1)  # the flat axis is not always z
    scales = torch.gather(scales, 1, perm).contiguous()
    q = torch.randn(P, 4, generator=g)
    rotations = (q / q.norm(dim=1, keepdim=True)).contiguous()
    op = torch.where(torch.rand(P, generator=g) < opaque_fraction, torch.tensor(0.99),
                     0.05 + 0.85 * torch.rand(P, generator=g))
    opacity = op.unsqueeze(1).contiguous()
    M = (deg + 1) ** 2
    rgb = torch.rand(P, 3, generator=g)
    shs = torch.zeros(P, M, 3)
    shs[:, 0, :] = RGB2SH(rgb)
    if M > 1:
        shs[:, 1:, :] = 0.05 * torch.randn(P, M - 1, 3, generator=g)
    return {"xyz": xyz.contiguous(), "scales": scales, "rotations": rotations, "opacity": opacity,
            "shs": shs.contiguous(), "rgb": rgb.contiguous(), "sh_degree": deg}


def make_tile_mask(name, kind="ones", seed=7):
    P, W, H = CONFIGS[name][:3]
    th, tw = (H + 15) // 16, (W + 15) // 16
    if kind == "ones":
        return torch.ones(th, tw, dtype=torch.int32)
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(th, tw, generator=g) < 0.5).to(torch.int32)


RENDER_DEFAULTS = dict(opaque_threshold=0.6, normal_threshold=math.cos(math.radians(60.0)), depth_threshold=1.0,
                       color_sigma=3.0, T_threshold=0.0001, scale_modifier=1.0)


# ---------------------------------------------------------------------------------------------------------------------
# Multi-object scene (BASELINE config 3, "Cube-Diorama room-shaped": ~20 object IDs + background), SURVEY.md 8d/8e.
# Every object owns its Gaussians; object 0 is the background shell (the largest unit of the bin packing).
# ---------------------------------------------------------------------------------------------------------------------
def object_counts(n_objects=20, seed=2024, background=200_000):
    """Gaussians per object, object 0 = background.  Deterministic: every rank computes the same table."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    counts = (20_000 * (1.0 + 3.0 * torch.rand(n_objects, generator=g))).long().tolist()
    return [int(background)] + [int(c) for c in counts]


def make_object(obj_id, count, cam_name="c1", seed=2024, sh_degree=3):
    """Activated Gaussians of one object: surfels on an ellipsoid placed in the camera frustum (object 0: a far wall)."""
    Pn, W, H, fx, fy, cx, cy, _ = CONFIGS[cam_name]
    g = torch.Generator(device="cpu").manual_seed(seed * 1000 + obj_id)
    cam = make_camera(cam_name)
    P = int(count)
    if obj_id == 0:
        z = 4.5 + 0.05 * torch.randn(P, generator=g)
        u, v = torch.rand(P, generator=g) * 2 - 1, torch.rand(P, generator=g) * 2 - 1
        pts_c = torch.stack([u * z * (W / (2 * fx)), v * z * (H / (2 * fy)), z], dim=1)
    else:
        zc = 1.5 + 2.5 * float(torch.rand(1, generator=g))
        uc, vc = (float(torch.rand(1, generator=g)) * 1.4 - 0.7), (float(torch.rand(1, generator=g)) * 1.4 - 0.7)
        centre = torch.tensor([uc * zc * (W / (2 * fx)), vc * zc * (H / (2 * fy)), zc])
        axes = 0.15 + 0.3 * torch.rand(3, generator=g)
        d = torch.randn(P, 3, generator=g)
        d = d / d.norm(dim=1, keepdim=True)
        pts_c = centre + d * axes * (1.0 + 0.02 * torch.randn(P, 1, generator=g))
    W2C = cam.world_view_transform.transpose(0, 1).double()
    C2W = torch.linalg.inv(W2C)
    xyz = (pts_c.double() @ C2W[:3, :3].T + C2W[:3, 3]).float()
    s = torch.exp(math.log(0.004) + (math.log(0.03) - math.log(0.004)) * torch.rand(P, 2, generator=g))
    scales = torch.cat([s, 0.1 * s.min(dim=1, keepdim=True).values], dim=1)
    scales = torch.gather(scales, 1, torch.argsort(torch.rand(P, 3, generator=g), dim=1)).contiguous()
    q = torch.randn(P, 4, generator=g)
    op = torch.where(torch.rand(P, generator=g) < 0.7, torch.tensor(0.99), 0.05 + 0.85 * torch.rand(P, generator=g))
    M = (sh_degree + 1) ** 2
    base = torch.rand(3, generator=g)
    shs = torch.zeros(P, M, 3)
    shs[:, 0, :] = RGB2SH((base + 0.15 * torch.randn(P, 3, generator=g)).clamp(0, 1))
    if M > 1:
        shs[:, 1:, :] = 0.05 * torch.randn(P, M - 1, 3, generator=g)
    return {"xyz": xyz.contiguous(), "scales": scales, "rotations": (q / q.norm(dim=1, keepdim=True)).contiguous(),
            "opacity": op.unsqueeze(1).contiguous(), "shs": shs.contiguous(), "sh_degree": sh_degree, "obj_id": obj_id}


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 5 shape: 3 M Gaussians at 1920x1080 with 64 objects + background.  The background
# shell is split SPATIALLY into pieces of about one object's size ("split spatially only if it dominates"):
# every unit of the bin packing is then small against a rank's share and LPT balances to a few percent at 8 ranks.
# ---------------------------------------------------------------------------------------------------------------------
def scene_units_c5(n_objects=64, total=3_000_000, background_fraction=0.2, seed=2024):
    """[(unit_id, n_gaussians, kind, piece, n_pieces)]: units 1..n_objects are objects, the others background pieces.
    Deterministic: every rank computes the same table."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = 0.5 + torch.rand(n_objects, generator=g)
    n_obj_total = int(total * (1.0 - background_fraction))
    counts = (w / w.sum() * n_obj_total).long().tolist()
    mean = n_obj_total // n_objects
    n_pieces = max(1, round((total - n_obj_total) / mean))
    piece = (total - sum(counts)) // n_pieces
    units = [(1 + i, int(c), "object", 0, 1) for i, c in enumerate(counts)]
    units += [(1 + n_objects + k, int(piece), "background", k, n_pieces) for k in range(n_pieces)]
    return units


def make_background_piece(unit_id, count, piece, n_pieces, cam_name="c5", seed=2024, sh_degree=3):
    """Vertical strip `piece` of `n_pieces` of the far wall (object 0 of make_object), same surfel statistics."""
    Pn, W, H, fx, fy, cx, cy, _ = CONFIGS[cam_name]
    g = torch.Generator(device="cpu").manual_seed(seed * 1000 + unit_id)
    cam = make_camera(cam_name)
    P = int(count)
    z = 4.5 + 0.05 * torch.randn(P, generator=g)
    lo, hi = -1.0 + 2.0 * piece / n_pieces, -1.0 + 2.0 * (piece + 1) / n_pieces
    u, v = lo + (hi - lo) * torch.rand(P, generator=g), torch.rand(P, generator=g) * 2 - 1
    pts_c = torch.stack([u * z * (W / (2 * fx)), v * z * (H / (2 * fy)), z], dim=1)
    W2C = cam.world_view_transform.transpose(0, 1).double()
    C2W = torch.linalg.inv(W2C)
    xyz = (pts_c.double() @ C2W[:3, :3].T + C2W[:3, 3]).float()
    s = torch.exp(math.log(0.004) + (math.log(0.03) - math.log(0.004)) * torch.rand(P, 2, generator=g))
    scales = torch.cat([s, 0.1 * s.min(dim=1, keepdim=True).values], dim=1)
    scales = torch.gather(scales, 1, torch.argsort(torch.rand(P, 3, generator=g), dim=1)).contiguous()
    q = torch.randn(P, 4, generator=g)
    op = torch.where(torch.rand(P, generator=g) < 0.7, torch.tensor(0.99), 0.05 + 0.85 * torch.rand(P, generator=g))
    M = (sh_degree + 1) ** 2
    base = torch.rand(3, generator=g)
    shs = torch.zeros(P, M, 3)
    shs[:, 0, :] = RGB2SH((base + 0.15 * torch.randn(P, 3, generator=g)).clamp(0, 1))
    if M > 1:
        shs[:, 1:, :] = 0.05 * torch.randn(P, M - 1, 3, generator=g)
    return {"xyz": xyz.contiguous(), "scales": scales, "rotations": (q / q.norm(dim=1, keepdim=True)).contiguous(),
            "opacity": op.unsqueeze(1).contiguous(), "shs": shs.contiguous(), "sh_degree": sh_degree, "obj_id": unit_id}
