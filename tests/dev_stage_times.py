"""Dev tool: per-stage CUDA-event times of the c2 forward+backward over the keyframe window (never a bench number).
    python tests/dev_stage_times.py [cfg] [views] [reps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import refharness as rh
import bench
from dqo_map_b200 import _lib, rasterizer

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
L = _lib.lib()
inp, views = bench.make_views(cfg, dev, 0, n_views)
cam0 = views[0]["cam"]
P, H, W, M = inp["xyz"].shape[0], cam0.image_height, cam0.image_width, inp["shs"].shape[1]
gc, gd = rh.make_pixel_grads(H, W, dev)
plans = []
for v in views:
    v["rs"] = v["settings"](rasterizer.GaussianRasterizationSettings)
    plans.append(rasterizer.plan_binning(v["rs"], inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"],
                                         inp["tile_mask"], shs=inp["shs"]))
two_phase = all(p[1] > 0 for p in plans)
front = max(p[1] for p in plans) if two_phase else 0
back = max(max(p[2] for p in plans) * 2, 1 << 19) if two_phase else 0
cap = front + back if two_phase else int(max(p[0] for p in plans) * 1.1) + 4096
pipe = rasterizer.RasterPipeline(P, M, W, H, cap, dev, front, back)
k = [0]


def step():
    v = views[k[0] % len(views)]
    k[0] += 1
    pipe.forward(v["rs"], inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"], shs=inp["shs"])
    pipe.backward(gc, gd)


ms = bench.timed(step, 40, 10, False) / 40
L.dqo_profile_enable(1)
acc = np.zeros(16)
n = reps * len(views)
for r in range(n):
    step()
    buf = (ctypes.c_float * 16)()
    L.dqo_profile_read(buf, 16)
    acc += np.array(list(buf))
L.dqo_profile_enable(0)
acc /= n
names = ["", "preprocess", "depth_sort", "", "emit", "tile_sort", "ranges", "render_front", "back_binning",
         "compact", "render_fwd2", "", "render_bwd", "gaussian_bwd"]
print("fwd+bwd %.4f ms  (front %d back %d)  status %s" % (ms, front, back, pipe.check()))
print("  ".join("%s %.3f" % (nm, acc[i]) for i, nm in enumerate(names) if nm))
