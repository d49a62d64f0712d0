"""CPU restatement (torch, float32) of the mapper's mask builders and error maps -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module; the product path never does.
Pinned: transmission2tilemask / colorerror2tilemask against tests/golden/maps.npz, produced by importing the reference's
own SLAM/utils.py in this container (tests/golden/make_maps_golden.py).  evaluate_render_range / error maps follow the
cited mapper.py lines literally (methods of a class that needs the whole SLAM stack: not importable, parity pinned only
through the two functions above and torch semantics)."""
import torch
import torch.nn.functional as F


def transmission2tilemask(pixelmask, stride, tile_mask_ratio=0.5):
    """SLAM/utils.py:752-763."""
    h, w = pixelmask.shape[:2]
    pad_h = (h + stride - 1) // stride * stride - h
    pad_w = (w + stride - 1) // stride * stride - w
    padded = F.pad(pixelmask, (0, pad_w, 0, pad_h))
    pooled = F.avg_pool2d(padded[None, None].float(), kernel_size=stride, stride=stride)
    return (pooled > tile_mask_ratio).int()[0, 0]


def tile_means(color_error, stride):
    """SLAM/utils.py:773-786 (pad with zeros, avg_pool2d)."""
    h, w = color_error.shape[:2]
    pad_h = (h + stride - 1) // stride * stride - h
    pad_w = (w + stride - 1) // stride * stride - w
    padded = F.pad(color_error, (0, pad_w, 0, pad_h), value=0)
    return F.avg_pool2d(padded[None, None].float(), kernel_size=stride, stride=stride)[0, 0]


def colorerror2tilemask(color_error, stride, top_ratio=0.4):
    """SLAM/utils.py:765-796; returns (tile_mask int32, tile means) -- the means let the tests tell genuine
    mismatches from torch.topk's unspecified order among equal values."""
    down = tile_means(color_error, stride)
    k = int(down.numel() * top_ratio)
    _, idx = torch.topk(down.reshape(-1), k=k)
    mask = torch.zeros(down.numel(), dtype=torch.int32)
    mask[idx] = 1
    return mask.view_as(down), down


def color_error_map(render_image, gt_image):
    """mapper.py:948-955; inputs [3,H,W]."""
    r = render_image.permute(1, 2, 0)
    diff = (r - gt_image.permute(1, 2, 0)).abs()
    err = torch.sum(diff, dim=-1, keepdim=False)
    err[r.sum(dim=-1) == 0] = 0
    return err


def evaluate_render_range(render_output, gt_image=None, gt_semantic=None, global_opt=False, sample_ratio=-1,
                          pixel_num=None):
    """mapper.py:944-987."""
    T_map = render_output["T_map"]
    H, W = T_map.shape[-2:]
    pixel_num = H * W if pixel_num is None else pixel_num
    if global_opt:
        if sample_ratio > 0:
            err = color_error_map(render_output["render"], gt_image)
            tile_mask, _ = colorerror2tilemask(err, 16, sample_ratio)
            if render_output.get("semantic_seg", None) is not None:
                serr = color_error_map(render_output["semantic_seg"], gt_semantic)
                tile_mask = tile_mask | colorerror2tilemask(serr, 16, sample_ratio)[0]
            render_mask = F.interpolate(tile_mask.float()[None, None], scale_factor=16, mode="nearest")[0, 0].bool()[:H, :W]
        else:
            render_mask = (T_map != 1).squeeze(0)
            tile_mask = None
    else:
        render_mask = (T_map != 1).squeeze(0)
        tile_mask = transmission2tilemask(render_mask, 16, 0.5)
    return render_mask, tile_mask, render_mask.sum() / pixel_num


def render_error_maps(render_output, color_map, depth_map):
    """mapper.py:1008-1025."""
    color = render_output["render"].permute(1, 2, 0)
    depth = render_output["depth"].permute(1, 2, 0)
    depth_index = render_output["depth_index_map"].permute(1, 2, 0)
    depth_error = torch.abs(depth_map - depth)
    depth_error[(depth_map - depth) < 0] = 0
    color_error = torch.sum(torch.abs(color_map - color), dim=-1, keepdim=True)
    normal_error = torch.zeros_like(depth_error)
    invalid = ((depth_map == 0) | (depth_index == -1)).squeeze()
    depth_error[invalid] = 0
    color_error[depth_map == 0] = 0
    normal_error[invalid] = 0
    return color_error, depth_error, normal_error
