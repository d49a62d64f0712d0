"""Dev tool: the config-2 mapping loop (5-keyframe window) run by the fused step, the stock torch loop around this
library's rasterizer and the stock torch loop around the reference rasterizer; final PSNR / depth L1 per view.
    python tests/dev_c2_loop.py [iters] [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import refharness as rh
import bench
from dqo_map_b200 import mapping, rasterizer
from test_gpu_mapping import psnr

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
ref_pkg = rh.load_reference()[0]
inp, views = bench.make_views("c2", dev, 0, 5)
cam0 = views[0]["cam"]
P, H, W = inp["xyz"].shape[0], cam0.image_height, cam0.image_width
for v in views:
    v["rs"] = v["settings"](rasterizer.GaussianRasterizationSettings)
    v["rs_ref"] = v["settings"](ref_pkg.GaussianRasterizationSettings)
    v["kf"] = bench.make_keyframe(inp, v["settings"], rasterizer)
R = max(rasterizer.plan_binning(v["rs"], inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"],
                                shs=inp["shs"])[0] for v in views)


def quality(params, v):
    with torch.no_grad():
        out = rasterizer.GaussianRasterizer(v["rs"])(
            means3D=params["xyz"], opacities=torch.sigmoid(params["opacity"]),
            shs=torch.cat((params["f_dc"], params["f_rest"]), dim=1), scales=torch.exp(params["scaling"]),
            rotations=torch.nn.functional.normalize(params["rotation"]), tile_mask=inp["tile_mask"])
    gt_color, gt_depth = v["kf"][0].permute(2, 0, 1), v["kf"][1].permute(2, 0, 1)
    hit = (out[3] != -1) & (gt_depth > 0)
    return psnr(out[0], gt_color), float((out[1] - gt_depth).abs()[hit].mean())


def fused():
    fparams = {k: t.contiguous() for k, t in bench.raw_params(inp).items()}
    st = mapping.FusedMappingStep(fparams, bench.LRS, W, H, 0.8, 1.0, 0.1, confidence=torch.zeros(P, 1, device=dev),
                                  capacity=int(R * 1.3) + 4096)
    st.begin_window(attach=True)
    losses = []
    for k in range(iters):
        v = views[k % len(views)]
        t = st(v["rs"], inp["tile_mask"], *v["kf"])
        if k < 12 or k % 50 == 0:
            losses.append(round(float(t[0]), 6))
    st.check()
    return fparams, losses


def stock(Rast, key):
    rparams = {k: torch.nn.Parameter(t) for k, t in bench.raw_params(inp).items()}
    init = {k: rparams[k].detach().clone() for k in ("xyz", "scaling", "rotation", "opacity")}
    conf = torch.zeros(P, 1, device=dev)
    opt = torch.optim.Adam([{"params": [rparams[k]], "lr": bench.LRS[k], "name": k} for k in bench.ORDER], lr=0.0, eps=1e-15)
    losses = []
    for k in range(iters):
        v = views[k % len(views)]
        t = bench.torch_mapping_iteration(rparams, init, opt, conf, Rast, v[key], inp["tile_mask"], *v["kf"])
        if k < 12 or k % 50 == 0:
            losses.append(round(float(t), 6))
    return {k: t.detach() for k, t in rparams.items()}, losses


print("start", [quality(bench.raw_params(inp), v) for v in views[:2]])
for r in range(reps):
    for name, fn in (("fused", fused), ("stock+ours", lambda: stock(rasterizer.GaussianRasterizer, "rs")),
                     ("stock+ref", lambda: stock(ref_pkg.GaussianRasterizer, "rs_ref"))):
        p, losses = fn()
        print("%-11s run %d: %s" % (name, r, " ".join("%.3f/%.5f" % quality(p, v) for v in views)))
        print("            losses", losses)
