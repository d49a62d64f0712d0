"""Map geometry around distCUDA2: scale initialisation of new Gaussians and the temp-point filter.

Host-side mirrors of `GaussianPointCloud.update_geometry` (reference SLAM/gaussian_pointcloud.py:519-570), `bbox_filter`
(SLAM/utils.py:801-808), `GaussianPointCloud.get_radius` (gaussian_pointcloud.py:739-743) and `Mapping.temp_points_filter`
(SLAM/multiprocess/mapper.py:1351-1380) over the C-ABI.  Same argument meaning and results; the point-cloud objects
stay with the caller (these functions return the new log-scales / the delete masks instead of mutating a class).
"""
import torch

from ._lib import check, lib, ptr
from .knn import distCUDA2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s expects CUDA tensors (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s expects float32 tensors" % name)
    return t.contiguous()


def bbox_filter(local_xyz, total_xyz, padding=0.05):
    """bool [n_total]: rows of total_xyz strictly inside the bounding box of local_xyz grown by `padding`."""
    local_c, total_c = _f32(local_xyz, "bbox_filter"), _f32(total_xyz, "bbox_filter")
    n_total = total_c.shape[0]
    mask = torch.empty((n_total,), dtype=torch.uint8, device=total_c.device)
    if n_total == 0:
        return mask.bool()
    ws = torch.empty((32,), dtype=torch.uint8, device=total_c.device)
    with torch.cuda.device(total_c.device):
        check(lib().dqo_bbox_mask(local_c.shape[0], ptr(local_c), n_total, ptr(total_c), float(padding), ptr(mask), None,
                                  ptr(ws), _stream()), "dqo_bbox_mask")
    return mask.view(torch.bool)


def get_radius(log_scaling):
    """(sum(exp(s)) - min(exp(s))) / 2 per Gaussian, from the raw (log) scaling parameter [P,3]."""
    s = _f32(log_scaling, "get_radius")
    out = torch.empty((s.shape[0],), dtype=torch.float32, device=s.device)
    if s.shape[0]:
        with torch.cuda.device(s.device):
            check(lib().dqo_gaussian_radius(s.shape[0], ptr(s), ptr(out), _stream()), "dqo_gaussian_radius")
    return out


def update_geometry(xyz, log_scaling, extra_xyz, extra_radius, min_radius=0.001, max_radius=0.05, scale_factor=1.0,
                    xyz_factor=(1.0, 1.0, 0.1)):
    """Scale initialisation of the points `xyz` [P,3] (raw scaling `log_scaling` [P,3]) against themselves and the
    existing Gaussians (`extra_xyz` [E,3], `extra_radius` [E]) -- GaussianPointCloud.update_geometry.

    Returns (log_scales [P,3] or None, invalid_mask bool [P]).  The reference assigns `_scaling = log_scales` and then
    deletes the rows of invalid_mask; when every row is invalid it deletes without touching `_scaling`: log_scales is
    None in that case (one 4-byte read-back decides it, as the reference's `(~mask).sum() == 0` does)."""
    xyz_c = _f32(xyz, "update_geometry")
    P = xyz_c.shape[0]
    dev = xyz_c.device
    if P == 0:
        return None, torch.zeros((0,), dtype=torch.bool, device=dev)
    radius = get_radius(log_scaling)
    if extra_xyz is not None and extra_xyz.numel() > 0:
        keep = bbox_filter(xyz_c, extra_xyz)
        extra_xyz, extra_radius = extra_xyz[keep], extra_radius[keep]
        total_xyz = torch.cat([xyz_c, _f32(extra_xyz, "update_geometry")])
        total_radius = torch.cat([radius, _f32(extra_radius.reshape(-1), "update_geometry")])
    else:
        total_xyz, total_radius = xyz_c, radius
    _, knn_idx = distCUDA2(total_xyz)
    log_scales = torch.empty((P, 3), dtype=torch.float32, device=dev)
    invalid = torch.empty((P,), dtype=torch.uint8, device=dev)
    valid_count = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().dqo_scale_init(P, total_xyz.shape[0], ptr(total_xyz), ptr(total_radius), ptr(knn_idx),
                                   float(min_radius), float(max_radius), float(scale_factor), float(xyz_factor[0]),
                                   float(xyz_factor[1]), float(xyz_factor[2]), ptr(log_scales), ptr(invalid),
                                   ptr(valid_count), _stream()), "dqo_scale_init")
    if int(valid_count.item()) == 0:
        return None, invalid.view(torch.bool)
    return log_scales, invalid.view(torch.bool)


def knn_points3(query, ref):
    """(squared distances [Q,3] ascending, indices [Q,3] int32) of the 3 nearest rows of `ref` for every row of `query`:
    pytorch3d.ops.knn_points(query[None], ref[None], K=3) without the batch dimension."""
    q, r = _f32(query, "knn_points3"), _f32(ref, "knn_points3")
    Q, R = q.shape[0], r.shape[0]
    d2 = torch.zeros((Q, 3), dtype=torch.float32, device=q.device)
    idx = torch.zeros((Q, 3), dtype=torch.int32, device=q.device)
    if Q == 0:
        return d2, idx
    L = lib()
    nbytes = L.dqo_knn_cross3_workspace_bytes(Q, R)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        check(L.dqo_knn_cross3(Q, ptr(q), R, ptr(r), ptr(d2), ptr(idx), ptr(ws), nbytes, _stream()), "dqo_knn_cross3")
    return d2, idx


def temp_points_filter(temp_xyz, exist_xyz, exist_radius, ratio=0.6):
    """bool [T] delete mask of Mapping.temp_points_filter: temp points that fall within `ratio` x radius of one of their
    3 nearest existing (unstable) Gaussians inside the temp points' bounding box.  Returns None where the reference
    returns without deleting (no existing Gaussian in the box)."""
    t = _f32(temp_xyz, "temp_points_filter")
    if exist_xyz is None or exist_xyz.numel() == 0 or t.numel() == 0:
        return None
    keep = bbox_filter(t, exist_xyz)
    ex, er = _f32(exist_xyz[keep], "temp_points_filter"), _f32(exist_radius.reshape(-1)[keep], "temp_points_filter")
    if ex.shape[0] == 0:
        return None
    d2, idx = knn_points3(t, ex)
    mask = torch.empty((t.shape[0],), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        check(lib().dqo_inside_mask(t.shape[0], ptr(d2), ptr(idx), ptr(er), float(ratio), ptr(mask), _stream()),
              "dqo_inside_mask")
    return mask.view(torch.bool)
