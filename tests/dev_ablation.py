"""Dev tool: ablations of the c2 forward+backward (never a bench number): n_touched on/off, single- vs two-phase binning."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import refharness as rh
import bench
from dqo_map_b200 import rasterizer

dev = torch.device("cuda:0")
inp, cam, settings = bench.build_workload("c2", dev, 0)
P, H, W, M = inp["xyz"].shape[0], cam.image_height, cam.image_width, inp["shs"].shape[1]
gc, gd = rh.make_pixel_grads(H, W, dev)
rs = settings(rasterizer.GaussianRasterizationSettings)
R, front, back = rasterizer.plan_binning(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"], shs=inp["shs"])


def run(pipe, nt):
    def step():
        pipe.forward(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"], shs=inp["shs"], need_n_touched=nt)
        pipe.backward(gc, gd)
    return bench.timed(step, 30, 5, False) / 30


two = rasterizer.RasterPipeline(P, M, W, H, front + back, dev, front, back)
one = rasterizer.RasterPipeline(P, M, W, H, int(R * 1.05) + 4096, dev)
for name, pipe in (("two-phase", two), ("single-phase", one)):
    for nt in (True, False):
        print("%-13s need_n_touched=%-5s %.4f ms" % (name, nt, run(pipe, nt)))
