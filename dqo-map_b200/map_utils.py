"""`cuda_utils._C.accumulate_gaussian_error` over the C-ABI (reference submodules/cuda_utils/cuda_utils.cu:17-60)."""
import torch

from ._lib import check, lib, ptr


def accumulate_gaussian_error(H, W, P, screen_color_error, screen_depth_error, screen_normal_error, screen_color_index,
                              screen_depth_index, color_threshold, depth_threshold, normal_threshold, check_max):
    """Same argument order (note H, W) and 4-tuple of [P,1] tensors as the reference:
    (gs_color_error, gs_depth_error, gs_normal_error, gs_rescale_counter)."""
    dev = screen_color_error.device
    if not screen_color_error.is_cuda:
        raise RuntimeError("accumulate_gaussian_error expects CUDA tensors (there is no CPU path)")
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    outs = [torch.empty((P, 1), **f32) for _ in range(3)]
    counters = [torch.empty((P, 1), **i32) for _ in range(3)]
    rescale = torch.empty((P, 1), **f32)
    ce = screen_color_error.contiguous().float()
    de = screen_depth_error.contiguous().float()
    ne = screen_normal_error.contiguous().float()
    ci = screen_color_index.contiguous().to(torch.int32)
    di = screen_depth_index.contiguous().to(torch.int32)
    for t in (ce, de, ne, ci, di):
        if t.numel() != H * W:
            raise ValueError("screen maps must have H*W elements")
    with torch.cuda.device(dev):
        check(lib().dqo_accumulate_error(W, H, P, ptr(ce), ptr(de), ptr(ne), ptr(ci), ptr(di), color_threshold,
                                         depth_threshold, normal_threshold, int(bool(check_max)), ptr(outs[0]),
                                         ptr(outs[1]), ptr(outs[2]), ptr(counters[0]), ptr(counters[1]),
                                         ptr(counters[2]), ptr(rescale), torch.cuda.current_stream().cuda_stream),
              "dqo_accumulate_error")
    return outs[0], outs[1], outs[2], rescale


# ---------------------------------------------------------------------------------------------------------------------
# Mask builders (SLAM/utils.py:752-796) and the image-space halves of Mapping.evaluate_render_range /
# Mapping.error_gaussians_remove (SLAM/multiprocess/mapper.py:930-1047).  Same names, arguments and return values as
# the reference functions; one kernel launch each instead of a chain of torch ops.
# ---------------------------------------------------------------------------------------------------------------------
def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s expects CUDA tensors (there is no CPU path)" % name)


def _stride16(stride, name):
    if int(stride) != 16:
        raise ValueError("%s: only the rasterizer's tile size (stride 16) is supported" % name)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def transmission2tilemask(pixelmask, stride=16, tile_mask_ratio=0.5):
    """int32 [ceil(H/16), ceil(W/16)] tile mask: 1 where more than `tile_mask_ratio` of the (zero padded) tile is set
    (SLAM/utils.py:752-763)."""
    _need_cuda(pixelmask, "transmission2tilemask")
    _stride16(stride, "transmission2tilemask")
    H, W = pixelmask.shape[:2]
    pm = (pixelmask != 0).contiguous().view(torch.uint8) if pixelmask.dtype != torch.bool else pixelmask.contiguous().view(torch.uint8)
    tile_mask = torch.empty(((H + 15) // 16, (W + 15) // 16), dtype=torch.int32, device=pixelmask.device)
    with torch.cuda.device(pixelmask.device):
        check(lib().dqo_pixelmask_to_tilemask(W, H, ptr(pm), float(tile_mask_ratio), ptr(tile_mask), None, _stream()),
              "dqo_pixelmask_to_tilemask")
    return tile_mask


def colorerror2tilemask(color_error, stride=16, top_ratio=0.4, out=None, return_pixel_mask=False):
    """int32 tile mask with the int(n_tiles * top_ratio) tiles of largest mean error set (SLAM/utils.py:765-796).
    `out`: an existing tile mask to OR into (mapper.py:969).  With return_pixel_mask also the nearest-upsampled,
    image-cropped bool mask of mapper.py:971-980."""
    _need_cuda(color_error, "colorerror2tilemask")
    _stride16(stride, "colorerror2tilemask")
    H, W = color_error.shape[:2]
    err = color_error.contiguous().float()
    th, tw = (H + 15) // 16, (W + 15) // 16
    k = int(th * tw * top_ratio)
    dev = color_error.device
    tile_mask = out if out is not None else torch.empty((th, tw), dtype=torch.int32, device=dev)
    if tile_mask.dtype != torch.int32 or tuple(tile_mask.shape) != (th, tw) or not tile_mask.is_contiguous():
        raise ValueError("out must be a contiguous int32 [ceil(H/16), ceil(W/16)] tensor")
    pixel = torch.empty((H, W), dtype=torch.bool, device=dev) if return_pixel_mask else None
    ws = torch.empty(lib().dqo_topk_tilemask_workspace_bytes(W, H), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().dqo_topk_tilemask(W, H, ptr(err), k, int(out is not None), ptr(tile_mask),
                                      ptr(pixel) if pixel is not None else None, ptr(ws), _stream()), "dqo_topk_tilemask")
    return (tile_mask, pixel) if return_pixel_mask else tile_mask


def color_error_map(render_image, gt_image):
    """[H,W] sum over channels of |render - gt| with never-rendered pixels (render.sum(c) == 0) zeroed
    (mapper.py:948-955).  Both inputs are [3,H,W] as the renderer and the frame hold them."""
    _need_cuda(render_image, "color_error_map")
    _, H, W = render_image.shape
    r, g = render_image.contiguous().float(), gt_image.contiguous().float()
    out = torch.empty((H, W), dtype=torch.float32, device=r.device)
    with torch.cuda.device(r.device):
        check(lib().dqo_color_error_map(W, H, ptr(r), ptr(g), ptr(out), _stream()), "dqo_color_error_map")
    return out


def evaluate_render_range(render_output, gt_image=None, gt_semantic=None, global_opt=False, sample_ratio=-1,
                          pixel_num=None):
    """Image-space half of Mapping.evaluate_render_range (mapper.py:944-987) given the renderer's output dict
    ("T_map" [1,H,W], "render" [3,H,W], optional "semantic_seg").  Returns (render_mask bool [H,W], tile_mask int32 or
    None, render_ratio 0-dim tensor) exactly as the reference does for the three cases
    local (transmission mask), global with sampling (top error tiles), global without (transmission, no tile mask)."""
    T_map = render_output["T_map"]
    _need_cuda(T_map, "evaluate_render_range")
    H, W = T_map.shape[-2:]
    dev = T_map.device
    pixel_num = H * W if pixel_num is None else pixel_num
    if global_opt and sample_ratio > 0:
        err = color_error_map(render_output["render"], gt_image)
        tile_mask = colorerror2tilemask(err, 16, sample_ratio)
        sem = render_output.get("semantic_seg", None)
        if sem is not None:
            colorerror2tilemask(color_error_map(sem, gt_semantic), 16, sample_ratio, out=tile_mask)
        render_mask = torch.empty((H, W), dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            check(lib().dqo_tilemask_to_pixelmask(W, H, ptr(tile_mask), ptr(render_mask), _stream()),
                  "dqo_tilemask_to_pixelmask")
        return render_mask, tile_mask, render_mask.sum() / pixel_num
    T = T_map.contiguous().float()
    render_mask = torch.empty((H, W), dtype=torch.bool, device=dev)
    tile_mask = torch.empty(((H + 15) // 16, (W + 15) // 16), dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().dqo_render_range(W, H, ptr(T), 0.5, ptr(render_mask), ptr(tile_mask), ptr(count), _stream()),
              "dqo_render_range")
    return render_mask, (None if global_opt else tile_mask), count[0] / pixel_num


def render_error_maps(render_output, color_map, depth_map):
    """Error images of Mapping.error_gaussians_remove (mapper.py:1008-1025): returns (color_error, depth_error,
    normal_error) as [H,W,1] float tensors ready for accumulate_gaussian_error.  `color_map` [H,W,3] and `depth_map`
    [H,W,1] are the observed frame maps; render_output holds "render" [3,H,W], "depth" [1,H,W], "depth_index_map"
    [1,H,W]."""
    color, depth, dindex = render_output["render"], render_output["depth"], render_output["depth_index_map"]
    _need_cuda(color, "render_error_maps")
    _, H, W = color.shape
    dev = color.device
    outs = [torch.empty((H, W, 1), dtype=torch.float32, device=dev) for _ in range(3)]
    c, d = color.contiguous().float(), depth.contiguous().float()
    gc, gd = color_map.contiguous().float(), depth_map.contiguous().float()
    di = dindex.contiguous().to(torch.int32)
    with torch.cuda.device(dev):
        check(lib().dqo_render_error_maps(W, H, ptr(c), ptr(d), ptr(gc), ptr(gd), ptr(di), ptr(outs[0]), ptr(outs[1]),
                                          ptr(outs[2]), _stream()), "dqo_render_error_maps")
    return outs[0], outs[1], outs[2]
