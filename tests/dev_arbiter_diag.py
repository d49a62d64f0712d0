"""Dev diagnostic: where do ours / the reference differ from the float64 arbiter (per Gaussian)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import refharness as rh
import test_gpu_rasterizer as T
from oracle import f64_arbiter
cfg = sys.argv[1] if len(sys.argv) > 1 else "deg1"
DEV = "cuda:0"
inp = rh.make_inputs(cfg, torch.device(DEV), mask="ones")
cam = inp["cam"]
gc, gd = rh.make_pixel_grads(cam.image_height, cam.image_width, DEV)
o, ex, bw = T._run_ours(inp, gc, gd)
scene, lists = T._arbiter_inputs(inp, o, ex)
exact = f64_arbiter.gradients(scene, lists, gc, gd, device=DEV)
ours = dict(zip(T.GRADS, bw))
C = rh.load_reference()[1]
fwd = C.rasterize_gaussians(*rh.raster_args(inp))
refs = [dict(zip(T.GRADS, C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd)))) for _ in range(4)]
ours2 = [dict(zip(T.GRADS, T._run_ours(inp, gc, gd)[2])) for _ in range(3)]
for name, e in exact.items():
    if e.numel() <= 1: continue
    e = e.double(); n = float(e.norm())
    eo = [float((x[name].double().reshape(e.shape) - e).norm()) / n for x in [ours] + ours2]
    er = [float((x[name].double().reshape(e.shape) - e).norm()) / n for x in refs]
    print("%-14s ours %s   ref %s" % (name, ["%.2e" % v for v in eo], ["%.2e" % v for v in er]))
e = exact["dL_drotations"].double()
do = (ours["dL_drotations"].double() - e).norm(dim=1)
dr = (refs[0]["dL_drotations"].double() - e).norm(dim=1)
top = torch.argsort(do, descending=True)[:8]
print("top ours-err Gaussians:", [(int(i), "%.2e" % float(do[i]), "ref %.2e" % float(dr[i]), "|g| %.2e" % float(e[i].norm())) for i in top])
print("share of squared error in top 8 (ours): %.2f  (ref, its own top 8): %.2f" % (
    float((do[top] ** 2).sum() / (do ** 2).sum()), float((torch.sort(dr, descending=True).values[:8] ** 2).sum() / (dr ** 2).sum())))
