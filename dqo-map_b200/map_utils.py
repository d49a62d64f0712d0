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
