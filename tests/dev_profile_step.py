"""Dev tool: a few fused mapping steps of the c2 workload for ncu (never a bench number).
    python tests/dev_profile_step.py [cfg] [iters] [single|terms]
`terms`: after the plain steps, steps with the optional loss terms -- the masked step with the semantic colour term and the
mask-less global pass with the SSIM term -- so that a launch list shows their kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dqo_map_b200 import rasterizer, mapping

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
inp, views = bench.make_views(cfg, dev, 0, 1)
v = views[0]
cam = v["cam"]
P, H, W = inp["xyz"].shape[0], cam.image_height, cam.image_width
rs = v["settings"](rasterizer.GaussianRasterizationSettings)
gt_color, gt_depth, mask = bench.make_keyframe(inp, v["settings"], rasterizer)
R, front, back = rasterizer.plan_binning(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"],
                                         shs=inp["shs"])
if len(sys.argv) > 3 and sys.argv[3] == "single":
    front = back = 0
fback = max(back * 2, 1 << 19) if front else 0
print("R", R, "front", front, "back", fback)
fparams = {k: t.contiguous() for k, t in bench.raw_params(inp).items()}
fstep = mapping.FusedMappingStep(fparams, bench.LRS, W, H, 0.8, 1.0, 0.1, confidence=torch.zeros(P, 1, device=dev),
                                 capacity=(front + fback) if front else int(R * 1.3) + 4096, front_instances=front,
                                 back_instances=fback)
fstep.begin_window(attach=True)
for _ in range(iters):
    t = fstep(rs, inp["tile_mask"], gt_color, gt_depth, mask)
torch.cuda.synchronize()
if iters >= 20:  # device time of the fused step alone (keyframe resident, no read-back): CUDA events over the next `iters` steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        t = fstep(rs, inp["tile_mask"], gt_color, gt_depth, mask)
    e1.record()
    torch.cuda.synchronize()
    print("fused step %.4f ms" % (e0.elapsed_time(e1) / iters))
print(float(t[0]), fstep.check())
if len(sys.argv) > 3 and sys.argv[3] == "terms":
    g = torch.Generator(device="cpu").manual_seed(7)
    params = {k: t.contiguous() for k, t in bench.raw_params(inp).items()}
    params["semantics"] = torch.rand(P, 3, generator=g).to(dev)
    gt_sem = torch.rand(H, W, 3, generator=g).to(dev)
    st = mapping.FusedMappingStep(params, dict(bench.LRS, semantics=5e-4), W, H, 0.8, 1.0, 0.1,
                                  capacity=(front + fback) if front else int(R * 1.3) + 4096, front_instances=front,
                                  back_instances=fback, ssim_weight=0.2, semantic_weight=0.1)
    st.begin_window(attach=True)
    for _ in range(iters):
        st(rs, inp["tile_mask"], gt_color, gt_depth, mask, gt_semantic=gt_sem)
    torch.cuda.synchronize()
    print("semantic step", [round(float(x), 6) for x in st.loss], st.check()[:4])
    for _ in range(iters):
        st(rs, inp["tile_mask"], gt_color, gt_depth, None)
    torch.cuda.synchronize()
    print("global pass with SSIM", [round(float(x), 6) for x in st.loss], st.check()[:4])
