import sys, os
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import torch
import test_gpu_mapping as t
from dqo_map_b200 import mapping
gt, cam, settings, raw, gt_color, gt_depth, render_mask = t._scene(P=6000, deg=3)
H, W = cam.image_height, cam.image_width
def fused():
    params = {k: v.clone().contiguous() for k, v in raw.items()}
    st = mapping.FusedMappingStep(params, t.LRS, W, H, 0.8, 1.0, 0.1)
    rs = settings()
    return [float(st(rs, gt["tile_mask"], gt_color, gt_depth, render_mask)[0]) for _ in range(8)]
def oper():
    pt = {k: torch.nn.Parameter(v.clone()) for k, v in raw.items()}
    ms = mapping.MappingStep(pt, t.LRS, settings, 0.8, 1.0, 0.1, optimizer="fused")
    return [float(ms(None, gt["tile_mask"], gt_color, gt_depth, render_mask)[0]) for _ in range(8)]
a = [fused() for _ in range(3)]; b = [oper() for _ in range(3)]
for i in range(8):
    print(i, " ".join("%.9f" % x[i] for x in a), "|", " ".join("%.9f" % x[i] for x in b))
