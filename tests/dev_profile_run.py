"""Dev tool: a few forward+backward iterations of the c2 workload for ncu (never a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import refharness as rh
import bench
from dqo_map_b200 import rasterizer

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
inp, cam, settings = bench.build_workload(cfg, dev, 0)
P, H, W, M = inp["xyz"].shape[0], cam.image_height, cam.image_width, inp["shs"].shape[1]
gc, gd = rh.make_pixel_grads(H, W, dev)
rs = settings(rasterizer.GaussianRasterizationSettings)
R, front, back = rasterizer.plan_binning(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"],
                                         shs=inp["shs"])
if len(sys.argv) > 3 and sys.argv[3] == "single":
    front = back = 0
print("R", R, "front", front, "back", back)
pipe = rasterizer.RasterPipeline(P, M, W, H, (front + back) if front else int(R * 1.05) + 4096, dev, front, back)
for _ in range(iters):
    pipe.forward(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"], shs=inp["shs"])
    pipe.backward(gc, gd)
torch.cuda.synchronize()
print(pipe.check())
