"""Host-side mirror of `Renderer.render` (reference SLAM/render.py:134-266): one call that returns the colour / depth /
normal / index / transmission maps and, when the Gaussians carry them, the semantic and instance colour images.

The reference obtains the two extra images by running the complete rasterizer two more times with colors_precomp
(render.py:227-262); those calls are differentiable: with the default configuration (use_semantics, semantic_color_weight
0.1) the L1 on "semantic_seg" back-propagates into the trainable `_semantics` group (gaussian_pointcloud.py:372-378) and
into the geometry.  Here
the view is preprocessed, binned and sorted ONCE and each extra image is one additional blend pass over the same lists
(`dqo_rast_blend_extra`), bit-identical to what the extra full call returns; while autograd is recording and any input of
an extra image requires grad, that pass is differentiable (`rasterizer.blend_extra_colors_grad`: one reverse blend + the
per-Gaussian chain, `dqo_rast_blend_extra_backward`), its gradients reach the colour tensor and the geometry like those
of the reference's second full call and are added to the main render's by autograd."""
import torch

from . import rasterizer


def render(raster_settings, gaussian_data, tile_mask=None):
    """gaussian_data: dict with xyz, opacity, scales, rotations, shs, normal and optional semantics_color / instance
    ([P,3] each), already activated, as `Renderer.render` receives it (render.py:180-186).  Returns the reference's result
    dict (render.py:218-266).  Gradients flow through "render", "depth" and -- via the differentiable extra blend, see the
    module docstring -- "semantic_seg" / "instance", as in the reference."""
    rs = raster_settings
    means3D = gaussian_data["xyz"]
    dev = means3D.device
    if tile_mask is None:  # render.py:191-198
        tile_mask = torch.ones(((int(rs.image_height) + 15) // 16, (int(rs.image_width) + 15) // 16), dtype=torch.int32,
                               device=dev)
    rast = rasterizer.GaussianRasterizer(rs)
    out = rast(means3D=means3D, opacities=gaussian_data["opacity"], shs=gaussian_data["shs"], colors_precomp=None,
               scales=gaussian_data["scales"], rotations=gaussian_data["rotations"], cov3D_precomp=None,
               normal_w=gaussian_data.get("normal"), tile_mask=tile_mask)
    state = rasterizer._RasterizeGaussians.last_state
    image, depth, color_index, depth_index, color_w, depth_w, T_map, n_touched = out[:8]
    results = {"render": image, "depth": depth, "color_index_map": color_index, "depth_index_map": depth_index,
               "color_hit_weight": color_w, "depth_hit_weight": depth_w, "T_map": T_map}
    normal = gaussian_data.get("normal")
    if normal is not None:  # render.py:212-216
        render_normal = torch.zeros_like(image)
        sel = depth_index[0] > -1
        render_normal[:, sel] = normal[depth_index[depth_index > -1].long()].permute(1, 0)
        results["normal"] = render_normal
    for key, name in (("semantics_color", "semantic_seg"), ("instance", "instance")):  # render.py:227-262
        c = gaussian_data.get(key)
        if c is not None and c.numel() > 1:
            needs_grad = torch.is_grad_enabled() and any(
                t is not None and t.requires_grad
                for t in (c, means3D, gaussian_data["opacity"], gaussian_data["scales"], gaussian_data["rotations"]))
            if needs_grad:
                results[name] = rasterizer.blend_extra_colors_grad(state, rs, c, means3D, gaussian_data["opacity"],
                                                                   gaussian_data["scales"], gaussian_data["rotations"])
            else:
                with torch.no_grad():
                    results[name] = rasterizer.blend_extra_colors(state, c, rs.bg)
        else:
            results[name] = None
    if n_touched is not None:
        results["n_touched"] = n_touched
    return results
