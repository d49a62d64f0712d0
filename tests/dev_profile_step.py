"""Dev tool: a few fused mapping steps of the c2 workload for ncu (never a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import refharness as rh
import bench
from dqo_map_b200 import rasterizer, mapping

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
inp, cam, settings = bench.build_workload(cfg, dev, 0)
P, H, W = inp["xyz"].shape[0], cam.image_height, cam.image_width
rs = settings(rasterizer.GaussianRasterizationSettings)
gt_color, gt_depth, mask = bench.make_keyframe(inp, cam, settings, rasterizer)
R, front, back = rasterizer.plan_binning(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"],
                                         shs=inp["shs"])
if len(sys.argv) > 3 and sys.argv[3] == "single":
    front = back = 0
fback = int(back * 1.3) + 65536 if front else 0
print("R", R, "front", front, "back", fback)
fparams = {k: v.contiguous() for k, v in bench.raw_params(inp).items()}
fstep = mapping.FusedMappingStep(fparams, bench.LRS, W, H, 0.8, 1.0, 0.1, confidence=torch.zeros(P, 1, device=dev),
                                 capacity=(front + fback) if front else int(R * 1.3) + 4096, front_instances=front,
                                 back_instances=fback)
for _ in range(iters):
    t = fstep(rs, inp["tile_mask"], gt_color, gt_depth, mask)
torch.cuda.synchronize()
print(float(t[0]), fstep.check())
