"""Dev check: at config-2 size (programmatic launches active) a CUDA-graph replay of the fused step gives the same loss
sequence as eager calls (both are reproducible: fp64 gradient accumulators).  python tests/dev_graph_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dqo_map_b200 import rasterizer, mapping

dev = torch.device("cuda:0")
inp, views = bench.make_views("c2", dev, 0, 1)
v = views[0]
cam = v["cam"]
P, H, W = inp["xyz"].shape[0], cam.image_height, cam.image_width
rs = v["settings"](rasterizer.GaussianRasterizationSettings)
kf = bench.make_keyframe(inp, v["settings"], rasterizer)
R, front, back = rasterizer.plan_binning(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"],
                                         shs=inp["shs"])
back = max(back * 2, 1 << 19)


def run(use_graph, iters=16):
    params = {k: t.contiguous() for k, t in bench.raw_params(inp).items()}
    st = mapping.FusedMappingStep(params, bench.LRS, W, H, 0.8, 1.0, 0.1, confidence=torch.zeros(P, 1, device=dev),
                                  capacity=front + back, front_instances=front, back_instances=back)
    st.begin_window(attach=True)
    losses = []
    g = None
    for k in range(iters):
        if use_graph and k == 4:
            torch.cuda.synchronize()
            g = st.graph(rs, inp["tile_mask"], *kf, warmup=False)
        if g is not None:
            g.replay()
            losses.append(float(st.loss[0]))
        else:
            losses.append(float(st(rs, inp["tile_mask"], *kf)[0]))
    st.check()
    return losses, params["xyz"].clone()


a, xa = run(False)
b, xb = run(True)
print("eager", ["%.8f" % x for x in a[-4:]])
print("graph", ["%.8f" % x for x in b[-4:]])
print("identical losses:", a == b, " identical positions:", bool(torch.equal(xa, xb)))
