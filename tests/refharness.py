"""Test-side helpers: load the compiled reference (oracle/_ref), decode its private buffers, run both
implementations on the same inputs.  Test infrastructure only — never imported by the product."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

from dqo_map_b200 import synthetic  # noqa: E402


def _load_pkg(name, path):
    init = os.path.join(path, "__init__.py")
    spec = importlib.util.spec_from_file_location(name, init, submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "diff_gaussian_rasterization_depth", "_C_depth.so"))


_ref_cache = {}


def load_reference():
    """Returns (rast_pkg, rast_C, knn_C, cuda_utils_C) of the UNMODIFIED reference built by oracle/build_ref.py."""
    if "r" in _ref_cache:
        return _ref_cache["r"]
    rast = _load_pkg("ref_dgr_depth", os.path.join(REF_DIR, "diff_gaussian_rasterization_depth"))
    rast_C = sys.modules["ref_dgr_depth._C_depth"]
    knn_pkg = _load_pkg("ref_simple_knn", os.path.join(REF_DIR, "simple_knn"))
    import importlib
    knn_C = importlib.import_module("ref_simple_knn._C")
    cu_pkg = _load_pkg("ref_cuda_utils", os.path.join(REF_DIR, "cuda_utils"))
    cu_C = importlib.import_module("ref_cuda_utils._C")
    _ref_cache["r"] = (rast, rast_C, knn_C, cu_C)
    return _ref_cache["r"]


def _al(x, a=128):
    return (x + a - 1) // a * a


def decode_ref_buffers(geom, binning, img, P, R, W, H):
    """Decodes the reference's three byte buffers (rasterizer_impl.cu:159-201 `fromChunk` layouts; every
    sub-array is aligned to 128 bytes from an (at least) 128-byte aligned base)."""
    g = geom.cpu().numpy()
    out = {}
    off = 0

    def take(buf, off, nbytes, dtype, shape):
        off = _al(off)
        arr = np.frombuffer(buf[off:off + nbytes].tobytes(), dtype=dtype).reshape(shape)
        return arr, off + nbytes

    out["depths"], off = take(g, off, 4 * P, np.float32, (P,))
    out["clamped"], off = take(g, off, 3 * P, np.uint8, (P, 3))
    out["internal_radii"], off = take(g, off, 4 * P, np.int32, (P,))
    out["means2D"], off = take(g, off, 8 * P, np.float32, (P, 2))
    out["cov3D"], off = take(g, off, 24 * P, np.float32, (P, 6))
    out["conic_opacity"], off = take(g, off, 16 * P, np.float32, (P, 4))
    out["rgb"], off = take(g, off, 12 * P, np.float32, (P, 3))
    out["tiles_touched"], off = take(g, off, 4 * P, np.uint32, (P,))
    if R > 0:
        b = binning.cpu().numpy()
        off = 0
        out["point_list"], off = take(b, off, 4 * R, np.uint32, (R,))
        _, off = take(b, off, 4 * R, np.uint32, (R,))
        out["keys_sorted"], off = take(b, off, 8 * R, np.uint64, (R,))
        out["keys_unsorted"], off = take(b, off, 8 * R, np.uint64, (R,))
    else:
        out["point_list"] = np.zeros((0,), np.uint32)
        out["keys_sorted"] = np.zeros((0,), np.uint64)
    im = img.cpu().numpy()
    N = W * H
    off = 0
    out["accum_alpha"], off = take(im, off, 4 * N, np.float32, (H, W))
    out["n_contrib"], off = take(im, off, 4 * N, np.uint32, (H, W))
    ranges, off = take(im, off, 8 * N, np.uint32, (N, 2))
    tiles = ((H + 15) // 16) * ((W + 15) // 16)
    out["ranges"] = ranges[:tiles].copy()
    return out


def make_inputs(cfg, device, seed=2024, P=None, sh_degree=None, mask="ones", precomp=False):
    gs = synthetic.make_gaussians(cfg, seed=seed, P=P, sh_degree=sh_degree)
    cam = synthetic.make_camera(cfg).to(device)
    d = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in gs.items()}
    d["cam"] = cam
    d["tile_mask"] = synthetic.make_tile_mask(cfg, mask).to(device)
    d["bg"] = torch.tensor([0.0, 0.0, 0.0], device=device) if not precomp else torch.tensor([0.1, 0.2, 0.3], device=device)
    d["precomp"] = precomp
    return d


def raster_args(inp, debug=False):
    """27-tuple of `_C_depth.rasterize_gaussians` (RAST/.../__init__.py:68-96)."""
    cam = inp["cam"]
    rd = synthetic.RENDER_DEFAULTS
    empty = torch.Tensor([])
    if inp["precomp"]:
        colors, sh, deg = inp["rgb"], empty, 0
    else:
        colors, sh, deg = empty, inp["shs"], inp["sh_degree"]
    return (inp["bg"], inp["xyz"], colors, inp["opacity"], inp["scales"], inp["rotations"], rd["scale_modifier"],
            empty, cam.world_view_transform, cam.full_proj_transform, inp["tile_mask"], cam.tanfovx, cam.tanfovy,
            cam.image_height, cam.image_width, cam.cx, cam.cy, sh, deg, rd["color_sigma"], cam.camera_center,
            rd["opaque_threshold"], rd["depth_threshold"], rd["normal_threshold"], rd["T_threshold"], False, debug)


def backward_args(inp, fwd, grad_color, grad_depth, debug=False):
    """29-tuple of `_C_depth.rasterize_gaussians_backward` (RAST/.../__init__.py:208-238)."""
    cam = inp["cam"]
    rd = synthetic.RENDER_DEFAULTS
    empty = torch.Tensor([])
    (rendered, tile_num, color, depth, hit_color, hit_depth, hcw, hdw, T_map, radii, geom, binning, img, tile_indices,
     n_touched) = fwd
    if inp["precomp"]:
        colors, sh, deg = inp["rgb"], empty, 0
    else:
        colors, sh, deg = empty, inp["shs"], inp["sh_degree"]
    return (tile_indices, tile_num, inp["bg"], inp["xyz"], radii, colors, inp["scales"], inp["rotations"],
            rd["scale_modifier"], empty, cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy,
            cam.cx, cam.cy, rd["depth_threshold"], rd["normal_threshold"], grad_color, grad_depth, sh, deg,
            cam.camera_center, geom, rendered, binning, img, hit_depth, debug)


def make_pixel_grads(H, W, device, seed=11):
    g = torch.Generator(device="cpu").manual_seed(seed)
    gc = torch.randn(3, H, W, generator=g) * 1e-3
    gd = torch.randn(1, H, W, generator=g) * 1e-3
    return gc.to(device), gd.to(device)


def export_ours(st, P, W, H):
    """Reference-format view of this library's private state (dqo_rast_export_state)."""
    from dqo_map_b200 import _lib
    L = _lib.lib()
    dev = st.status.device
    cap = max(int(st.capacity), 1)
    tiles = ((H + 15) // 16) * ((W + 15) // 16)
    keys = torch.zeros(cap, dtype=torch.int64, device=dev)
    plist = torch.zeros(cap, dtype=torch.int32, device=dev)
    ranges = torch.zeros((tiles, 2), dtype=torch.int32, device=dev)
    ncontrib = torch.zeros((H, W), dtype=torch.int32, device=dev)
    final_T = torch.zeros((H, W), dtype=torch.float32, device=dev)
    means2D = torch.zeros((P, 2), dtype=torch.float32, device=dev)
    depths = torch.zeros((P,), dtype=torch.float32, device=dev)
    conic = torch.zeros((P, 4), dtype=torch.float32, device=dev)
    rgb = torch.zeros((P, 3), dtype=torch.float32, device=dev)
    tiles_touched = torch.zeros((P,), dtype=torch.int32, device=dev)
    p = _lib.ptr
    lists = st.settings.front_instances == 0  # the full sorted list only exists after single-phase binning
    if not lists:
        keys = plist = ranges = None
    _lib.check(L.dqo_rast_export_state(st.settings, p(st.geom), p(st.binning), st.capacity, p(st.image), p(st.status),
                                       p(keys), p(plist), p(ranges), p(ncontrib), p(final_T), p(means2D), p(depths),
                                       p(conic), p(rgb), p(tiles_touched), torch.cuda.current_stream().cuda_stream),
               "dqo_rast_export_state")
    torch.cuda.synchronize()
    R = int(st.status[0].item())
    return {
        "keys_sorted": keys.cpu().numpy().view(np.uint64)[:R] if lists else None,
        "point_list": plist.cpu().numpy().view(np.uint32)[:R] if lists else None,
        "ranges": ranges.cpu().numpy().view(np.uint32) if lists else None,
        "n_contrib": ncontrib.cpu().numpy().view(np.uint32),
        "accum_alpha": final_T.cpu().numpy(), "means2D": means2D.cpu().numpy(), "depths": depths.cpu().numpy(),
        "conic_opacity": conic.cpu().numpy(), "rgb": rgb.cpu().numpy(),
        "tiles_touched": tiles_touched.cpu().numpy().view(np.uint32),
    }
