"""Host-side mirror of the reference rasterizer binding.

Same names, argument meaning and error behaviour as
`submodules/diff-gaussian-rasterizer-depth/diff_gaussian_rasterization_depth/__init__.py`
(GaussianRasterizationSettings :288-307, GaussianRasterizer :310-376, _RasterizeGaussians :53-285) and as the
pybind functions of `rasterize_points.cu` (rasterize_gaussians :37-155, rasterize_gaussians_backward :157-249,
mark_visible :251-270), implemented over the C-ABI of include/dqo_b200.h.  PyTorch only owns memory and streams.
"""
from typing import NamedTuple

import copy

import torch
import torch.nn as nn

from . import _lib
from ._lib import RastSettings, check, lib, ptr
from .binning_policy import BinningPolicy


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    opaque_threshold: float
    normal_threshold: float
    depth_threshold: float
    prefiltered: bool
    debug: bool
    cx: float
    cy: float
    color_sigma: float = 3.0
    T_threshold: float = 0.0001


# ---------------------------------------------------------------------------------------------------
# workspace policy: the instance capacity is remembered per device and grows geometrically.  The forward
# never synchronises in the middle of the pipeline; one status read-back at the end tells whether the
# capacity was sufficient (otherwise the pass is re-run with a larger buffer).
# ---------------------------------------------------------------------------------------------------
_capacity_hint = {}
NEED_N_TOUCHED = True  # reference behaviour (forward.cu:833-835); set False to skip the per-pair counter
# "auto": binning_policy.BinningPolicy picks single-phase or occlusion-aware two-phase binning per device from the status
# words of earlier calls (identical results either way); "single": always bin every instance, which is what
# dqo_rast_export_state needs to rebuild the reference's full sorted list.
BINNING = "auto"
_FIXED = (0, 0)
_policy = {}


def set_binning_mode(mode, front_instances=0, back_instances=0):
    """'auto' (default), 'single', or 'fixed' with explicit (front_instances, back_instances) -- the latter raises on
    overflow instead of adapting (tests, experiments)."""
    global BINNING, _FIXED
    if mode not in ("auto", "single", "fixed"):
        raise ValueError("binning mode must be 'auto', 'single' or 'fixed'")
    if mode == "fixed" and (front_instances <= 0 or front_instances % 256 or back_instances <= 0):
        raise ValueError("fixed binning needs front_instances (multiple of 256) and back_instances > 0")
    BINNING, _FIXED = mode, (int(front_instances), int(back_instances))
    _policy.clear()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t):
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        raise TypeError("expected float32 tensor, got %s" % t.dtype)
    return t.contiguous()


def _make_settings(P, D, M, W, H, tanfovx, tanfovy, cx, cy, scale_modifier, color_sigma, opaque_threshold,
                   depth_threshold, normal_threshold, T_threshold, prefiltered, debug, need_n_touched=True,
                   front_instances=0, back_instances=0, geom_clean=False):
    return RastSettings(P, D, M, W, H, tanfovx, tanfovy, cx, cy, scale_modifier, color_sigma, opaque_threshold,
                        depth_threshold, normal_threshold, T_threshold, int(bool(prefiltered)), int(bool(debug)),
                        int(bool(need_n_touched)), int(front_instances), int(back_instances), int(geom_clean))


class ForwardState:
    """Everything the backward needs; plays the role of (geomBuffer, binningBuffer, imgBuffer, tile_indices,
    num_rendered, num_tile) in the reference."""
    __slots__ = ("settings", "geom", "binning", "image", "tile_indices", "status", "capacity", "status_host", "radii")


def _forward_impl(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                  projmatrix, render_mask, tan_fovx, tan_fovy, image_height, image_width, cx, cy, sh, degree,
                  color_sigma, campos, opaque_threshold, hit_depth_threshold, hit_normal_threshold, T_threshold,
                  prefiltered, debug, sync=True, tile_indices_len=None):
    L = lib()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:67-70
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (there is no CPU path)")
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    sh_c, colors_c = _f32c(sh), _f32c(colors)
    M = sh_c.size(1) if sh_c is not None else 0
    means_c, opac_c = _f32c(means3D), _f32c(opacity)
    scales_c, rot_c, cov_c = _f32c(scales), _f32c(rotations), _f32c(cov3D_precomp)
    bg_c, view_c, proj_c, campos_c = _f32c(background), _f32c(viewmatrix), _f32c(projmatrix), _f32c(campos)
    if render_mask is None:
        raise TypeError("tile_mask must be an int32 CUDA tensor of shape [ceil(H/16), ceil(W/16)]")
    mask_c = render_mask.contiguous()
    tiles = ((H + 15) // 16) * ((W + 15) // 16)
    if mask_c.dtype != torch.int32 or mask_c.numel() != tiles:
        raise TypeError("tile_mask must be an int32 tensor with %d elements" % tiles)

    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    out_color = torch.empty((3, H, W), **f32)
    out_depth = torch.empty((1, H, W), **f32)
    out_hit_depth = torch.empty((1, H, W), **i32)
    out_hit_color = torch.empty((1, H, W), **i32)
    out_hit_cw = torch.empty((1, H, W), **f32)
    out_hit_dw = torch.empty((1, H, W), **f32)
    out_T = torch.empty((1, H, W), **f32)
    radii = torch.empty((P,), **i32)
    n_touched = torch.empty((P,), **i32)
    tile_indices = torch.empty((tile_indices_len or tiles,), **i32)
    if tile_indices.numel() > tiles:
        tile_indices[tiles:] = -1
    status = torch.empty((_lib.ST_WORDS,), **i32)

    st = ForwardState()
    st.image = torch.empty((L.dqo_rast_image_bytes(W, H),), dtype=torch.uint8, device=dev)
    st.tile_indices, st.status = tile_indices, status
    st.geom = torch.empty((L.dqo_rast_geom_bytes(P) if P > 0 else 0,), dtype=torch.uint8, device=dev)
    # Host-side state is keyed by what determines the workload, not by the device alone: the instance capacity by
    # (device, image size, cloud size), the binning policy additionally by the view (the camera tensors of a keyframe
    # persist, so their address identifies it) -- alternating keyframes with different R keep their own front / back split
    # instead of thrashing one policy.  Both tables are bounded.
    key = (dev.index, W, H, P)
    pkey = key + (viewmatrix.data_ptr(),)
    if len(_policy) > 512:
        _policy.clear()
    if len(_capacity_hint) > 512:
        _capacity_hint.clear()
    policy = None
    if BINNING == "auto" and sync:
        policy = _policy.get(pkey)
        if policy is None:  # a view seen for the first time starts from what the last view of the same workload learned
            last = _policy.get(key)
            policy = _policy[pkey] = copy.copy(last) if last is not None else BinningPolicy()
            policy.history = list(policy.history)
        _policy[key] = policy
    capacity = max(_capacity_hint.get(key, 0), 4 * P, 1 << 16) if P > 0 else 0
    while True:
        front, back = policy.plan(1 << 30) if (policy is not None and P > 0) else (0, 0)
        if BINNING == "fixed" and P > 0:
            front, back = _FIXED
        cap = front + back if front > 0 else capacity
        st.settings = _make_settings(P, int(degree), M, W, H, tan_fovx, tan_fovy, cx, cy, scale_modifier, color_sigma,
                                     opaque_threshold, hit_depth_threshold, hit_normal_threshold, T_threshold,
                                     prefiltered, debug, NEED_N_TOUCHED, front, back)
        st.capacity = cap
        st.binning = torch.empty((L.dqo_rast_binning_bytes(cap) if P > 0 else 0,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            code = L.dqo_rast_forward(
                st.settings, ptr(bg_c), ptr(means_c), ptr(sh_c), ptr(colors_c), ptr(opac_c), ptr(scales_c), ptr(rot_c),
                ptr(cov_c), ptr(view_c), ptr(proj_c), ptr(campos_c), ptr(mask_c), ptr(st.geom), ptr(st.binning),
                cap, ptr(st.image), ptr(tile_indices), ptr(out_color), ptr(out_depth), ptr(out_hit_depth),
                ptr(out_hit_color), ptr(out_hit_cw), ptr(out_hit_dw), ptr(out_T), ptr(radii), ptr(n_touched),
                ptr(status), _stream())
        check(code, "dqo_rast_forward")
        if not sync:
            st.status_host = None
            break
        host = status.tolist()  # the single host synchronisation of the forward, after all launches
        st.status_host = host
        retry = policy.update(host, front, back) if policy is not None else bool(host[_lib.ST_OVERFLOW])
        if retry and BINNING == "fixed":
            raise _lib.DqoError("fixed two-phase binning overflowed: back phase needs %d instances, %d given"
                                % (host[_lib.ST_R_BACK], back))
        if front == 0:
            capacity = int(host[_lib.ST_NUM_RENDERED] * 1.25) + 1024
            _capacity_hint[key] = capacity if retry else max(_capacity_hint.get(key, 0), capacity)
        if not retry:
            break
    st.radii = radii
    outs = (out_color, out_depth, out_hit_color, out_hit_depth, out_hit_cw, out_hit_dw, out_T, radii, n_touched)
    return st, outs


def _backward_impl(st, background, means3D, radii, colors, scales, rotations, cov3D_precomp, viewmatrix, projmatrix,
                   dL_dout_color, dL_dout_depth, sh, campos, hit_image):
    L = lib()
    dev = means3D.device
    s = st.settings
    P, M = s.P, s.M
    f32 = dict(dtype=torch.float32, device=dev)
    dL_dmeans3D = torch.empty((P, 3), **f32)
    dL_dmeans2D = torch.empty((P, 3), **f32)
    dL_dcolors = torch.empty((P, 3), **f32)
    dL_dconic = torch.empty((P, 2, 2), **f32)
    dL_dopacity = torch.empty((P, 1), **f32)
    dL_dcov3D = torch.empty((P, 6), **f32)
    dL_dsh = torch.empty((P, M, 3), **f32)
    dL_dscales = torch.empty((P, 3), **f32)
    dL_drotations = torch.empty((P, 4), **f32)
    if P != 0:
        g_color, g_depth = _f32c(dL_dout_color), _f32c(dL_dout_depth)
        with torch.cuda.device(dev):
            code = L.dqo_rast_backward(
                s, ptr(_f32c(background)), ptr(_f32c(means3D)), ptr(_f32c(sh)), ptr(_f32c(colors)),
                ptr(_f32c(scales)), ptr(_f32c(rotations)), ptr(_f32c(cov3D_precomp)), ptr(_f32c(viewmatrix)),
                ptr(_f32c(projmatrix)), ptr(_f32c(campos)), ptr(radii), ptr(st.geom), ptr(st.binning), st.capacity,
                ptr(st.image), ptr(st.status), ptr(g_color), ptr(g_depth), ptr(hit_image.contiguous()),
                ptr(dL_dmeans2D), ptr(dL_dconic), ptr(dL_dopacity), ptr(dL_dcolors), ptr(dL_dmeans3D), ptr(dL_dcov3D),
                ptr(dL_dsh), ptr(dL_dscales), ptr(dL_drotations), _stream())
        check(code, "dqo_rast_backward")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


class RasterPipeline:
    """Pre-allocated, synchronisation-free forward+backward for hot loops (bench.py, fused mapping step).

    All buffers (outputs, workspaces, gradients) are allocated once for a fixed (P, M, W, H, capacity); `forward()` and
    `backward()` only enqueue kernels.  `check()` reads the device status words (one host sync) and raises if the
    instance capacity was exceeded — call it whenever convenient, e.g. together with the loss read-back."""

    def __init__(self, P, M, W, H, capacity, device, front_instances=0, back_instances=0):
        L = lib()
        dev = torch.device(device)
        if front_instances and front_instances + back_instances > capacity:
            raise ValueError("front + back instances exceed the capacity")
        self.front, self.back = int(front_instances), int(back_instances)
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.P, self.M, self.W, self.H, self.capacity, self.dev = P, M, W, H, int(capacity), dev
        tiles = ((H + 15) // 16) * ((W + 15) // 16)
        self.color, self.depth = torch.empty((3, H, W), **f32), torch.empty((1, H, W), **f32)
        self.hit_depth, self.hit_color = torch.empty((1, H, W), **i32), torch.empty((1, H, W), **i32)
        self.hit_cw, self.hit_dw, self.T = (torch.empty((1, H, W), **f32) for _ in range(3))
        self.radii, self.n_touched = torch.empty((P,), **i32), torch.empty((P,), **i32)
        self.tile_indices, self.status = torch.empty((tiles,), **i32), torch.zeros((_lib.ST_WORDS,), **i32)
        self.geom = torch.empty((L.dqo_rast_geom_bytes(P),), **u8)
        # persistent workspace: the gradient accumulators are cleared once here and kept clean by every backward pass
        with torch.cuda.device(device):
            check(L.dqo_rast_geom_init(P, ptr(self.geom), _stream()), "dqo_rast_geom_init")
        self.binning = torch.empty((L.dqo_rast_binning_bytes(self.capacity),), **u8)
        self.image = torch.empty((L.dqo_rast_image_bytes(W, H),), **u8)
        # gradient tensors: owned by the pipeline, zero-initialised once; with geom_clean = 2 the backward only rewrites
        # the rows that are or were non-zero (do not write into them from outside)
        self.g_means2D, self.g_conic = torch.zeros((P, 3), **f32), torch.zeros((P, 4), **f32)
        self.g_opacity, self.g_colors = torch.zeros((P, 1), **f32), torch.zeros((P, 3), **f32)
        self.g_means3D, self.g_cov3D = torch.zeros((P, 3), **f32), torch.zeros((P, 6), **f32)
        self.g_sh = torch.zeros((P, max(M, 0), 3), **f32)
        self.g_scales, self.g_rot = torch.zeros((P, 3), **f32), torch.zeros((P, 4), **f32)
        self.settings = None

    def forward(self, rs, means3D, opacities, scales, rotations, tile_mask, shs=None, colors_precomp=None,
                need_n_touched=True):
        self.settings = _make_settings(self.P, int(rs.sh_degree), self.M, self.W, self.H, rs.tanfovx, rs.tanfovy, rs.cx,
                                       rs.cy, rs.scale_modifier, rs.color_sigma, rs.opaque_threshold, rs.depth_threshold,
                                       rs.normal_threshold, rs.T_threshold, rs.prefiltered, rs.debug, need_n_touched,
                                       self.front, self.back, geom_clean=2)  # persistent, zero-initialised gradient buffers
        self._in = (rs, means3D, shs, colors_precomp, scales, rotations)
        check(lib().dqo_rast_forward(
            self.settings, ptr(rs.bg), ptr(means3D), ptr(shs), ptr(colors_precomp), ptr(opacities), ptr(scales),
            ptr(rotations), None, ptr(rs.viewmatrix), ptr(rs.projmatrix), ptr(rs.campos), ptr(tile_mask), ptr(self.geom),
            ptr(self.binning), self.capacity, ptr(self.image), ptr(self.tile_indices), ptr(self.color), ptr(self.depth),
            ptr(self.hit_depth), ptr(self.hit_color), ptr(self.hit_cw), ptr(self.hit_dw), ptr(self.T), ptr(self.radii),
            ptr(self.n_touched), ptr(self.status), _stream()), "dqo_rast_forward")

    def backward(self, dL_dcolor, dL_ddepth):
        rs, means3D, shs, colors_precomp, scales, rotations = self._in
        check(lib().dqo_rast_backward(
            self.settings, ptr(rs.bg), ptr(means3D), ptr(shs), ptr(colors_precomp), ptr(scales), ptr(rotations), None,
            ptr(rs.viewmatrix), ptr(rs.projmatrix), ptr(rs.campos), ptr(self.radii), ptr(self.geom), ptr(self.binning),
            self.capacity, ptr(self.image), ptr(self.status), ptr(dL_dcolor), ptr(dL_ddepth), ptr(self.hit_depth),
            ptr(self.g_means2D), ptr(self.g_conic), ptr(self.g_opacity), ptr(self.g_colors), ptr(self.g_means3D),
            ptr(self.g_cov3D), ptr(self.g_sh), ptr(self.g_scales), ptr(self.g_rot), _stream()), "dqo_rast_backward")

    def check(self):
        host = self.status.tolist()
        if host[_lib.ST_OVERFLOW]:
            raise _lib.DqoError("instance capacity %d (front %d, back %d) exceeded (R = %d, back needs %d): results invalid, "
                                "re-create the pipeline larger" % (self.capacity, self.front, self.back,
                                                                   host[_lib.ST_NUM_RENDERED], host[_lib.ST_R_BACK]))
        return host


def plan_binning(rs, means3D, opacities, scales, rotations, tile_mask, shs=None, colors_precomp=None):
    """Sizing helper for the pre-allocated paths (RasterPipeline, mapping.FusedMappingStep): one synchronous
    single-phase forward on the given view and, when binning_policy opts for two-phase binning, one two-phase trial.
    Returns (R, front_instances, back_instances); front == 0 means single phase (capacity ~ 1.05 R), otherwise the
    buffers need front + back instances."""
    P = means3D.shape[0]
    M = shs.shape[1] if shs is not None else 0
    H, W = int(rs.image_height), int(rs.image_width)
    probe = RasterPipeline(P, M, W, H, max(4 * P, 1 << 16), means3D.device)
    while True:
        probe.forward(rs, means3D, opacities, scales, rotations, tile_mask, shs=shs, colors_precomp=colors_precomp)
        host = probe.status.tolist()
        if not host[_lib.ST_OVERFLOW]:
            break
        probe = RasterPipeline(P, M, W, H, int(host[_lib.ST_NUM_RENDERED] * 1.05) + 4096, means3D.device)
    R = host[_lib.ST_NUM_RENDERED]
    policy = BinningPolicy()
    policy.update(host, 0, 0)
    front, back = policy.plan(1 << 30)
    del probe
    if front > 0:
        trial = RasterPipeline(P, M, W, H, front + back, means3D.device, front, back)
        trial.forward(rs, means3D, opacities, scales, rotations, tile_mask, shs=shs, colors_precomp=colors_precomp)
        policy.update(trial.check(), front, back)
        front, back = policy.plan(1 << 30)
    return R, front, back


# ---------------------------------------------------------------------------------------------------
# pybind-level compatibility functions (`_C_depth.*`, RAST/ext.cpp:15-19)
# ---------------------------------------------------------------------------------------------------
def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, render_mask, tan_fovx, tan_fovy, image_height, image_width, cx, cy,
                        sh, degree, color_sigma, campos, opaque_threshold, hit_depth_threshold, hit_normal_threshold,
                        T_threshold, prefiltered, debug):
    """Same 27 arguments and 15-tuple as the reference's `rasterize_gaussians`.  `geomBuffer` carries the
    private forward state object (callers only pass it back to `rasterize_gaussians_backward`)."""
    st, outs = _forward_impl(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                             viewmatrix, projmatrix, render_mask, tan_fovx, tan_fovy, image_height, image_width, cx,
                             cy, sh, degree, color_sigma, campos, opaque_threshold, hit_depth_threshold,
                             hit_normal_threshold, T_threshold, prefiltered, debug, sync=True,
                             tile_indices_len=int(image_height) * int(image_width))
    (color, depth, hit_color, hit_depth, hit_cw, hit_dw, T_map, radii, n_touched) = outs
    rendered = st.status_host[_lib.ST_NUM_RENDERED]
    tile_num = st.status_host[_lib.ST_TILE_NUM]
    st.geom._dqo_state = st  # keep the state reachable from the returned buffer
    return (rendered, tile_num, color, depth, hit_color, hit_depth, hit_cw, hit_dw, T_map, radii, st.geom, st.binning,
            st.image, st.tile_indices, n_touched)


def rasterize_gaussians_backward(tile_indices, tile_num, background, means3D, radii, colors, scales, rotations,
                                 scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, cx, cy,
                                 depth_threshold, normal_threshold, dL_dout_color, dL_dout_depth, sh, degree, campos,
                                 geomBuffer, R, binningBuffer, imageBuffer, hit_image, debug):
    """Same 29 arguments and 8-tuple as the reference's `rasterize_gaussians_backward`."""
    st = getattr(geomBuffer, "_dqo_state", None)
    if st is None:
        raise RuntimeError("geomBuffer was not produced by this library's rasterize_gaussians")
    return _backward_impl(st, background, means3D, radii, colors, scales, rotations, cov3D_precomp, viewmatrix,
                          projmatrix, dL_dout_color, dL_dout_depth, sh, campos, hit_image)


def blend_extra_colors(state, colors_precomp, background):
    """[3,H,W] colour image of the view held by `state` (the ForwardState of a previous forward, reachable as
    `geomBuffer._dqo_state` / `ctx.state` / `GaussianRasterizer.last_state`) blended with `colors_precomp` [P,3] instead of
    the colours it was rendered with -- what a second full rasterizer call with colors_precomp would return as its first
    output (SLAM/render.py:227-262), for the cost of one blend pass.  No gradient."""
    s = state.settings
    if colors_precomp.shape[0] != s.P or colors_precomp.dim() != 2 or colors_precomp.shape[1] != 3:
        raise ValueError("colors_precomp must be [P,3] for the P Gaussians of the rendered view")
    dev = state.status.device
    out = torch.empty((3, s.H, s.W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().dqo_rast_blend_extra(s, ptr(_f32c(background)), ptr(_f32c(colors_precomp.detach())), ptr(state.geom),
                                         ptr(state.binning), state.capacity, ptr(state.image), ptr(state.status),
                                         ptr(out), _stream()), "dqo_rast_blend_extra")
    return out


class _BlendExtraColors(torch.autograd.Function):
    """Differentiable form of `blend_extra_colors`: forward = one blend pass over the lists of the main render, backward =
    one reverse blend + the per-Gaussian chain (`dqo_rast_blend_extra_backward`).  The geometric inputs are arguments only
    so that autograd routes the extra image's share of their gradients to them."""

    @staticmethod
    def forward(ctx, colors, means3D, opacities, scales, rotations, state, rs):
        ctx.state, ctx.rs, ctx.opacity_shape = state, rs, tuple(opacities.shape)
        ctx.save_for_backward(colors, means3D, scales, rotations)
        return blend_extra_colors(state, colors, rs.bg)

    @staticmethod
    def backward(ctx, grad_image):
        st, rs = ctx.state, ctx.rs
        colors, means3D, scales, rotations = ctx.saved_tensors
        s = st.settings
        P, dev = s.P, means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        g_colors, g_means2D = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32)
        g_conic, g_opacity = torch.empty((P, 2, 2), **f32), torch.empty((P, 1), **f32)
        g_means3D, g_cov3D = torch.empty((P, 3), **f32), torch.empty((P, 6), **f32)
        g_scales, g_rot = torch.empty((P, 3), **f32), torch.empty((P, 4), **f32)
        acc = torch.zeros((4 * P,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib().dqo_rast_blend_extra_backward(
                s, ptr(_f32c(rs.bg)), ptr(_f32c(colors)), ptr(_f32c(means3D)), ptr(_f32c(scales)), ptr(_f32c(rotations)),
                None, ptr(_f32c(rs.viewmatrix)), ptr(_f32c(rs.projmatrix)), ptr(_f32c(rs.campos)), ptr(st.radii),
                ptr(st.geom), ptr(st.binning), st.capacity, ptr(st.image), ptr(st.status), ptr(_f32c(grad_image)),
                ptr(acc), ptr(g_colors), ptr(g_means2D), ptr(g_conic), ptr(g_opacity), ptr(g_means3D), ptr(g_cov3D),
                ptr(g_scales), ptr(g_rot), _stream()), "dqo_rast_blend_extra_backward")
        return g_colors, g_means3D, g_opacity.reshape(ctx.opacity_shape), g_scales, g_rot, None, None


def blend_extra_colors_grad(state, raster_settings, colors_precomp, means3D, opacities, scales, rotations):
    """`blend_extra_colors` with gradients: what the reference obtains by running the whole rasterizer a second time with
    colors_precomp and back-propagating through it (SLAM/render.py:227-262) -- the image is identical, the gradients reach
    colors_precomp and, through alpha and the projected geometry, means3D / opacities / scales / rotations -- for one extra
    blend forward and one reverse blend instead of a second preprocess + binning + sort + render.  `state` must be the
    ForwardState of a forward pass over exactly these geometric inputs (scale/rotation form, not cov3D_precomp)."""
    if state.settings.P == 0:
        return blend_extra_colors(state, colors_precomp, raster_settings.bg)
    return _BlendExtraColors.apply(colors_precomp, means3D, opacities, scales, rotations, state, raster_settings)


def mark_visible(means3D, viewmatrix, projmatrix):
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        with torch.cuda.device(means3D.device):
            check(lib().dqo_mark_visible(P, ptr(_f32c(means3D)), ptr(_f32c(viewmatrix)), ptr(_f32c(projmatrix)),
                                         ptr(present), _stream()), "dqo_mark_visible")
    return present


# ---------------------------------------------------------------------------------------------------
# autograd wrapper (same structure as the reference's _RasterizeGaussians)
# ---------------------------------------------------------------------------------------------------
def rasterize_gaussians_fn(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, tile_mask,
                           raster_settings):
    return _RasterizeGaussians.apply(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     tile_mask, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, tile_mask,
                raster_settings):
        rs = raster_settings
        st, outs = _forward_impl(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                                 cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, tile_mask, rs.tanfovx, rs.tanfovy,
                                 rs.image_height, rs.image_width, rs.cx, rs.cy, sh, rs.sh_degree, rs.color_sigma,
                                 rs.campos, rs.opaque_threshold, rs.depth_threshold, rs.normal_threshold,
                                 rs.T_threshold, rs.prefiltered, rs.debug, sync=True)
        (color, depth, hit_color, hit_depth, hit_cw, hit_dw, T_map, radii, n_touched) = outs
        ctx.raster_settings = rs
        ctx.state = st
        _RasterizeGaussians.last_state = st
        ctx.num_rendered = st.status_host[_lib.ST_NUM_RENDERED]
        ctx.num_tile = st.status_host[_lib.ST_TILE_NUM]
        ctx.save_for_backward(colors_precomp, hit_depth, means3D, scales, rotations, cov3Ds_precomp, radii, sh)
        ctx.mark_non_differentiable(hit_color, hit_depth, n_touched, radii)
        return color, depth, hit_color, hit_depth, hit_cw, hit_dw, T_map, n_touched, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth, grad_hit_color, grad_hit_depth, grad_hit_color_weight,
                 grad_hit_depth_weight, grad_T_map, grad_n_touched, _):
        # only colour and depth gradients are consumed, like the reference (__init__.py:176-187, N5)
        rs = ctx.raster_settings
        colors_precomp, hit_depth, means3D, scales, rotations, cov3Ds_precomp, radii, sh = ctx.saved_tensors
        H, W = int(rs.image_height), int(rs.image_width)
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, H, W), dtype=torch.float32, device=means3D.device)
        if grad_out_depth is None:
            grad_out_depth = torch.zeros((1, H, W), dtype=torch.float32, device=means3D.device)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = _backward_impl(ctx.state, rs.bg, means3D, radii, colors_precomp, scales, rotations,
                                          cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, grad_out_color,
                                          grad_out_depth, sh, rs.campos, hit_depth)

        def _opt(g, x):  # gradients for absent (empty) inputs are dropped
            return g if (x is not None and x.numel() != 0) else None

        return (grad_means3D, _opt(grad_sh, sh), _opt(grad_colors_precomp, colors_precomp), grad_opacities,
                _opt(grad_scales, scales), _opt(grad_rotations, rotations), _opt(grad_cov3Ds_precomp, cov3Ds_precomp),
                None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, tile_mask=None, normal_w=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians_fn(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                      tile_mask, raster_settings)
