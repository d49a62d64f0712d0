"""Python face of the CPU oracle (oracle/dqo_oracle.c + numpy/torch-CPU restatements).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product.  Parity status: pinned against tests/golden/*.npz, which tests/golden/make_golden.py
produced by running the unmodified reference CUDA extensions (oracle/_ref) on a B200.
"""
import ctypes as C
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
THREADS = 1


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "dqo_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "_build/liboracle.so"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_higher_msb.restype = C.c_uint32
        _lib.orc_higher_msb.argtypes = [C.c_uint32]
    return _lib


def set_threads(n):
    global THREADS
    THREADS = max(1, int(n))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _split(n, parts):
    parts = max(1, min(parts, n)) if n > 0 else 1
    step = (n + parts - 1) // parts if n > 0 else 0
    return [(i, min(n, i + step)) for i in range(0, n, step)] if n > 0 else []


def _run_ranges(fn, n, chunks_per_thread=4):
    rngs = _split(n, THREADS * chunks_per_thread if THREADS > 1 else 1)
    if THREADS <= 1 or len(rngs) <= 1:
        for b, e in rngs:
            fn(b, e)
        return
    with ThreadPoolExecutor(max_workers=THREADS) as ex:
        list(ex.map(lambda r: fn(*r), rngs))


def higher_msb(n):
    return int(lib().orc_higher_msb(int(n)))


class Scene:
    """Inputs of one rasterizer call in numpy form (same meaning as the 27 pybind arguments)."""

    def __init__(self, xyz, scales, rotations, opacity, view, proj, campos, W, H, tanfovx, tanfovy, cx, cy, bg,
                 tile_mask, shs=None, sh_degree=0, colors_precomp=None, scale_modifier=1.0, color_sigma=3.0,
                 opaque_threshold=0.6, depth_threshold=1.0, normal_threshold=0.5, T_threshold=1e-4):
        self.xyz, self.scales, self.rotations = _f32(xyz), _f32(scales), _f32(rotations)
        self.opacity = _f32(opacity).reshape(-1)
        self.view, self.proj, self.campos = _f32(view).reshape(-1), _f32(proj).reshape(-1), _f32(campos).reshape(-1)
        self.W, self.H = int(W), int(H)
        self.tanfovx, self.tanfovy, self.cx, self.cy = float(tanfovx), float(tanfovy), float(cx), float(cy)
        self.bg = _f32(bg)
        self.tile_mask = np.ascontiguousarray(tile_mask, dtype=np.int32)
        self.shs = _f32(shs) if shs is not None and np.size(shs) else None
        self.colors_precomp = _f32(colors_precomp) if colors_precomp is not None and np.size(colors_precomp) else None
        self.D = int(sh_degree)
        self.M = self.shs.shape[1] if self.shs is not None else 0
        self.scale_modifier, self.color_sigma = float(scale_modifier), float(color_sigma)
        self.opaque_threshold, self.depth_threshold = float(opaque_threshold), float(depth_threshold)
        self.normal_threshold, self.T_threshold = float(normal_threshold), float(T_threshold)
        self.P = self.xyz.shape[0]

    @classmethod
    def from_golden(cls, g):
        pre = bool(int(g["precomp"]))
        return cls(g["xyz"], g["scales"], g["rotations"], g["opacity"], g["viewmatrix"], g["projmatrix"], g["campos"],
                   int(g["W"]), int(g["H"]), float(g["tanfovx"]), float(g["tanfovy"]), float(g["cx"]), float(g["cy"]),
                   g["bg"], g["tile_mask"], shs=None if pre else g["shs"], sh_degree=0 if pre else int(g["sh_degree"]),
                   colors_precomp=g["rgb_in"] if pre else None, normal_threshold=math.cos(math.radians(60.0)))


def mark_visible(xyz, view, proj):
    xyz = _f32(xyz)
    out = np.zeros(xyz.shape[0], np.uint8)
    lib().orc_mark_visible(C.c_int(xyz.shape[0]), _p(xyz), _p(_f32(view).reshape(-1)), _p(_f32(proj).reshape(-1)), _p(out))
    return out.astype(bool)


def preprocess(sc):
    P = sc.P
    o = {"radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32), "depths": np.zeros(P, np.float32),
         "cov3D": np.zeros((P, 6), np.float32), "rgb": np.zeros((P, 3), np.float32),
         "conic_opacity": np.zeros((P, 4), np.float32), "clamped": np.zeros((P, 3), np.uint8),
         "tiles_touched": np.zeros(P, np.uint32)}
    L = lib()

    def run(b, e):
        L.orc_preprocess(C.c_int(P), C.c_int(sc.D), C.c_int(sc.M), C.c_float(sc.color_sigma), _p(sc.xyz), _p(sc.scales),
                         C.c_float(sc.scale_modifier), _p(sc.rotations), _p(sc.opacity), _p(sc.shs), None,
                         _p(sc.colors_precomp), _p(sc.view), _p(sc.proj), _p(sc.campos), _p(sc.tile_mask), C.c_int(sc.W),
                         C.c_int(sc.H), C.c_float(sc.tanfovx), C.c_float(sc.tanfovy), C.c_float(sc.cx), C.c_float(sc.cy),
                         _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
                         _p(o["conic_opacity"]), _p(o["clamped"]), _p(o["tiles_touched"]), C.c_int(b), C.c_int(e))

    _run_ranges(run, P, 1)
    return o


def binning(sc, pre):
    tiles = ((sc.H + 15) // 16) * ((sc.W + 15) // 16)
    R = int(pre["tiles_touched"].sum())
    keys = np.zeros(max(R, 1), np.uint64)
    vals = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((tiles, 2), np.uint32)
    tile_indices = np.zeros(tiles, np.int32)
    n = lib().orc_binning(C.c_int(sc.P), C.c_int(sc.W), C.c_int(sc.H), _p(pre["radii"]), _p(pre["means2D"]),
                          _p(pre["depths"]), _p(pre["tiles_touched"]), _p(sc.tile_mask), _p(keys), _p(vals), _p(ranges),
                          _p(tile_indices))
    if n < 0:
        raise RuntimeError("oracle binning: instance count mismatch")
    return {"num_rendered": R, "keys_sorted": keys[:R], "point_list": vals[:R], "ranges": ranges,
            "tile_indices": tile_indices, "tile_num": int(n)}


def render_forward(sc, pre, bn, need_n_touched=True):
    H, W, P = sc.H, sc.W, sc.P
    N = H * W
    o = {"color": np.zeros((3, H, W), np.float32), "depth": np.zeros((1, H, W), np.float32),
         "hit_depth": np.zeros((1, H, W), np.int32), "hit_color": np.zeros((1, H, W), np.int32),
         "hit_color_weight": np.zeros((1, H, W), np.float32), "hit_depth_weight": np.zeros((1, H, W), np.float32),
         "T_map": np.ones((1, H, W), np.float32), "n_touched": np.zeros(P, np.int32),
         "accum_alpha": np.ones((H, W), np.float32), "n_contrib": np.zeros((H, W), np.uint32),
         "hit_normal_c": np.zeros((N, 3), np.float32), "hit_point_c": np.zeros((N, 3), np.float32),
         "weight_sum": np.zeros((H, W), np.float32)}
    feats = sc.colors_precomp if sc.colors_precomp is not None else pre["rgb"]
    L = lib()

    def run(b, e):
        L.orc_render_forward(
            C.c_int(W), C.c_int(H), C.c_int(P), C.c_float(sc.tanfovx), C.c_float(sc.tanfovy), C.c_float(sc.cx),
            C.c_float(sc.cy), C.c_float(sc.scale_modifier), _p(sc.view), _p(sc.xyz), _p(sc.scales), _p(sc.rotations),
            _p(sc.bg), C.c_float(sc.opaque_threshold), C.c_float(sc.depth_threshold), C.c_float(sc.normal_threshold),
            C.c_float(sc.T_threshold), _p(bn["ranges"]), _p(bn["point_list"]), _p(bn["tile_indices"]),
            C.c_int(bn["tile_num"]), _p(pre["means2D"]), _p(feats), _p(pre["depths"]), _p(pre["conic_opacity"]),
            _p(o["accum_alpha"]), _p(o["n_contrib"]), _p(o["hit_normal_c"]), _p(o["hit_point_c"]), _p(o["color"]),
            _p(o["depth"]), _p(o["hit_depth"]), _p(o["hit_color"]), _p(o["hit_color_weight"]), _p(o["hit_depth_weight"]),
            _p(o["T_map"]), _p(o["weight_sum"]), _p(o["n_touched"]) if need_n_touched else None, C.c_int(b), C.c_int(e))

    _run_ranges(run, bn["tile_num"])
    return o


def forward(sc):
    pre = preprocess(sc)
    bn = binning(sc, pre)
    img = render_forward(sc, pre, bn)
    return pre, bn, img


def backward(sc, pre, bn, img, grad_color, grad_depth):
    P, M = sc.P, sc.M
    gc, gd = _f32(grad_color), _f32(grad_depth).reshape(-1)
    g = {"dL_dmeans2D": np.zeros((P, 3), np.float32), "dL_dconic": np.zeros((P, 4), np.float32),
         "dL_dopacity": np.zeros((P, 1), np.float32), "dL_dcolors": np.zeros((P, 3), np.float32),
         "dL_dmeans3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
         "dL_dsh": np.zeros((P, M, 3), np.float32), "dL_dscales": np.zeros((P, 3), np.float32),
         "dL_drotations": np.zeros((P, 4), np.float32)}
    feats = sc.colors_precomp if sc.colors_precomp is not None else pre["rgb"]
    hit = np.ascontiguousarray(img["hit_depth"].reshape(-1))
    L = lib()

    def run(b, e):
        L.orc_render_backward(
            C.c_int(sc.W), C.c_int(sc.H), C.c_float(sc.tanfovx), C.c_float(sc.tanfovy), C.c_float(sc.cx), C.c_float(sc.cy),
            C.c_float(sc.normal_threshold), C.c_float(sc.depth_threshold), _p(sc.view), _p(sc.scales), _p(sc.rotations),
            _p(sc.xyz), _p(sc.bg), _p(bn["ranges"]), _p(bn["point_list"]), _p(bn["tile_indices"]), C.c_int(bn["tile_num"]),
            _p(pre["means2D"]), _p(pre["conic_opacity"]), _p(feats), _p(img["accum_alpha"]), _p(img["n_contrib"]), _p(gc),
            _p(gd), _p(hit), _p(img["hit_normal_c"]), _p(img["hit_point_c"]), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
            _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_drotations"]), C.c_int(b), C.c_int(e))

    _run_ranges(run, bn["tile_num"])

    def run2(b, e):
        L.orc_preprocess_backward(
            C.c_int(P), C.c_int(sc.D), C.c_int(M), _p(sc.xyz), _p(pre["radii"]), _p(sc.shs), _p(pre["clamped"]),
            _p(sc.scales), _p(sc.rotations), C.c_float(sc.scale_modifier), _p(pre["cov3D"]), _p(sc.view), _p(sc.proj),
            C.c_float(sc.tanfovx), C.c_float(sc.tanfovy), C.c_int(sc.W), C.c_int(sc.H), _p(sc.campos), _p(g["dL_dmeans2D"]),
            _p(g["dL_dconic"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcolors"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]),
            _p(g["dL_dscales"]), _p(g["dL_drotations"]), C.c_int(b), C.c_int(e))

    _run_ranges(run2, P, 1)
    return g


def knn(points):
    pts = _f32(points)
    P = pts.shape[0]
    md = np.zeros(P, np.float32)
    ki = np.zeros((P, 3), np.int32)
    lib().orc_knn(C.c_int(P), _p(pts), _p(md), _p(ki))
    return md, ki


def accumulate_error(H, W, P, ce, de, ne, ci, di, cthr, dthr, nthr, check_max):
    outs = [np.zeros((P, 1), np.float32) for _ in range(4)]
    lib().orc_accumulate_error(C.c_int(W), C.c_int(H), C.c_int(P), _p(_f32(ce)), _p(_f32(de)), _p(_f32(ne)),
                               _p(np.ascontiguousarray(ci, np.int32)), _p(np.ascontiguousarray(di, np.int32)),
                               C.c_float(cthr), C.c_float(dthr), C.c_float(nthr), C.c_int(int(check_max)), _p(outs[0]),
                               _p(outs[1]), _p(outs[2]), _p(outs[3]))
    return outs


# -------------------------------------------------------------------------------------------------
# mapping step restatements (reference: inline PyTorch, SLAM/multiprocess/mapper.py:830-875, torch.optim.Adam)
# -------------------------------------------------------------------------------------------------
def masked_l1_loss(image, depth, hit_depth, gt_color, gt_depth, render_mask, color_weight, depth_weight, depth_err_thres):
    """torch-CPU autograd restatement of the loss lines of Mapping.loss_update.  Returns
    (total, colour, depth, dL_dimage[3,H,W], dL_ddepth[1,H,W])."""
    import torch
    img = torch.tensor(np.asarray(image), dtype=torch.float32, requires_grad=True)
    dep = torch.tensor(np.asarray(depth), dtype=torch.float32, requires_grad=True)
    gtc = torch.tensor(np.asarray(gt_color), dtype=torch.float32)
    gtd = torch.tensor(np.asarray(gt_depth), dtype=torch.float32).reshape(dep.shape[1], dep.shape[2], 1)
    hit = torch.tensor(np.asarray(hit_depth)).permute(1, 2, 0)
    image_p, depth_p = img.permute(1, 2, 0), dep.permute(1, 2, 0)
    if render_mask is None:
        mask = torch.ones(image_p.shape[:2], dtype=torch.bool)
    else:
        mask = torch.tensor(np.asarray(render_mask)).bool()
    color_loss = torch.abs(image_p[mask] - gtc[mask]).mean()
    depth_loss = torch.tensor(0.0)
    if depth_weight > 0:
        depth_error = depth_p - gtd
        valid = (hit != -1).squeeze(-1) & (gtd > 0).squeeze(-1) & (depth_error < depth_err_thres).squeeze(-1) & mask
        depth_loss = torch.abs(depth_error[valid]).mean()
    total = depth_weight * depth_loss + color_weight * color_loss
    if torch.isfinite(total):
        total.backward()
    gi = img.grad if img.grad is not None else torch.zeros_like(img)
    gd = dep.grad if dep.grad is not None else torch.zeros_like(dep)
    return float(total), float(color_loss), float(depth_loss), gi.numpy(), gd.numpy()


def adam_reference(params, grads_per_step, lrs, eps=1e-15, betas=(0.9, 0.999)):
    """torch.optim.Adam on CPU (single-tensor path) over several steps; returns final params, exp_avg, exp_avg_sq."""
    import torch
    ps = [torch.nn.Parameter(torch.tensor(p, dtype=torch.float32)) for p in params]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)], lr=0.0, eps=eps, betas=betas,
                           foreach=False)
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = None if g is None else torch.tensor(g, dtype=torch.float32)
        opt.step()
    outs = []
    for p in ps:
        st = opt.state.get(p, {})
        outs.append((p.detach().numpy(), st.get("exp_avg", torch.zeros_like(p)).numpy(),
                     st.get("exp_avg_sq", torch.zeros_like(p)).numpy()))
    return outs
