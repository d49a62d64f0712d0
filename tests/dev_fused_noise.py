"""Dev tool: run-to-run and path-to-path spread of the 200-iteration mapping loop (deg-3 scene of
test_fused_c_step_loop_quality): stock torch loop vs fused C step, several repetitions in one process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import test_gpu_mapping as T
from dqo_map_b200 import rasterizer

ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 200
DEG = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gt, cam, settings, raw, gt_color, gt_depth, render_mask = T._scene(deg=DEG)
H, W = cam.image_height, cam.image_width
args = (raw, settings, gt["tile_mask"], gt_color, gt_depth, render_mask)
for rep in range(3):
    r = T._loop(*args, ITERS, "ours_torch")
    print("stock psnr %.4f dl1 %.6f" % (r[0], r[1]))
for rep in range(3):
    params_c, _, _ = T._fused_c_loop(*args, ITERS, W, H)
    with torch.no_grad():
        out = rasterizer.GaussianRasterizer(settings())(
            means3D=params_c["xyz"], opacities=torch.sigmoid(params_c["opacity"]),
            shs=torch.cat((params_c["f_dc"], params_c["f_rest"]), dim=1), scales=torch.exp(params_c["scaling"]),
            rotations=torch.nn.functional.normalize(params_c["rotation"]), tile_mask=gt["tile_mask"])
    hit = (out[3] != -1) & (gt_depth.permute(2, 0, 1) > 0)
    print("fused psnr %.4f dl1 %.6f" % (T.psnr(out[0], gt_color.permute(2, 0, 1)),
                                          float((out[1] - gt_depth.permute(2, 0, 1)).abs()[hit].mean())))
r = T._loop(*args, ITERS, "fused")
print("opfused psnr %.4f dl1 %.6f" % (r[0], r[1]))
