#!/usr/bin/env python
"""Secondary benchmark: BASELINE config 3, "Cube-Diorama room-shaped multi-object mapping" -- ~20 objects + background,
each with its own Gaussians, per-object masked colour/depth loss, objects sharded across the ranks by Gaussian count
(dqo_map_b200.sharding.assign_objects, LPT bin packing), no gradient exchange.  Metric: mapping objects*iters/s.

    python bench_objects.py [--impl ours|reference] [--gpus N --steps K --warmup W] [--objects 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_objects.py --gpus N

A "step" is one mapping iteration of every object the rank owns (render of the object's Gaussians with its tile mask ->
masked L1 colour + depth loss -> backward -> Adam).  `value`: keyframes resident in HBM; `e2e`: every object's keyframe
(colour, depth, mask) is copied from pinned host memory each step and every object's loss is read back.  The total
number of objects is fixed as N grows (strong scaling); `config.load_imbalance` = max / mean Gaussians per rank.
bench.py (the driver's contract, BASELINE config 2, weak scaling) is the headline; this script is not run by the driver.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (shared helpers: timing protocol, clock sampler, stock mapping iteration)

METRIC, UNIT = "mapping objects*iters/s", "objects*iters/s"
CAM = "c1"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=20)
    ap.add_argument("--streams", type=int, default=8,
                    help="ours: objects are independent, their steps are enqueued round-robin on this many CUDA streams")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import refharness as rh
    if a.impl == "reference" and not rh.reference_available():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from dqo_map_b200 import map_utils, mapping, rasterizer, sharding, synthetic

    counts = synthetic.object_counts(a.objects)
    owner, load = sharding.assign_objects(counts, world)
    mine = sharding.local_objects(owner, rank)
    rows = max(sum(1 for o in owner if owner[o] == r) for r in range(world))
    cam = synthetic.make_camera(CAM).to(dev)
    H, W = cam.image_height, cam.image_width
    rd = synthetic.RENDER_DEFAULTS
    bg = torch.zeros(3, device=dev)
    rasterizer.set_binning_mode("single")

    def settings(Sett, deg):
        return Sett(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
                    viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=deg,
                    campos=cam.camera_center, opaque_threshold=rd["opaque_threshold"],
                    normal_threshold=rd["normal_threshold"], depth_threshold=rd["depth_threshold"], prefiltered=False,
                    debug=False, cx=cam.cx, cy=cam.cy)

    if a.impl == "ours":
        Rast, Sett = rasterizer.GaussianRasterizer, rasterizer.GaussianRasterizationSettings
    else:
        pkg = rh.load_reference()[0]
        Rast, Sett = pkg.GaussianRasterizer, pkg.GaussianRasterizationSettings

    objs = []
    ones = torch.ones(((H + 15) // 16, (W + 15) // 16), dtype=torch.int32, device=dev)
    for o in mine:
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic.make_object(o, counts[o], CAM).items()}
        d["bg"], d["tile_mask"] = bg, ones
        rs = settings(rasterizer.GaussianRasterizationSettings, d["sh_degree"])
        # keyframe of this object = render of a perturbed copy; its mask and tile mask as Mapping.evaluate_render_range
        # builds them (mapper.py:983-987)
        gt_color, gt_depth, _ = bench.make_keyframe(d, lambda S: settings(S, d["sh_degree"]), rasterizer)
        with torch.no_grad():
            cur = rasterizer.GaussianRasterizer(rs)(means3D=d["xyz"], opacities=d["opacity"], shs=d["shs"],
                                                    scales=d["scales"], rotations=d["rotations"], tile_mask=ones)
        render_mask, tile_mask, _ = map_utils.evaluate_render_range({"T_map": cur[6]})
        R = int(rasterizer._RasterizeGaussians.last_state.status_host[0])
        raw = {k: v.contiguous() for k, v in bench.raw_params(d).items()}
        host = [t.cpu().pin_memory() for t in (gt_color, gt_depth, render_mask)]
        e = {"id": o, "P": counts[o], "rs_ours": rs, "rs": settings(Sett, d["sh_degree"]), "tile_mask": tile_mask.contiguous(),
             "kf": [gt_color, gt_depth, render_mask.contiguous()], "host": host, "slot": [torch.empty_like(t, device=dev) for t in host]}
        if a.impl == "ours":
            e["step"] = mapping.FusedMappingStep(raw, bench.LRS, W, H, 0.8, 1.0, 0.1,
                                                 confidence=torch.zeros(counts[o], 1, device=dev), capacity=int(R * 1.5) + 65536)
            e["step"].begin_window(attach=True)
        else:
            e["params"] = {k: torch.nn.Parameter(v) for k, v in raw.items()}
            e["init"] = {k: e["params"][k].detach().clone() for k in ("xyz", "scaling", "rotation", "opacity")}
            groups = [{"params": [e["params"][k]], "lr": bench.LRS[k], "name": k} for k in mapping.FusedMappingStep.ORDER]
            e["opt"] = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
            e["conf"] = torch.zeros(counts[o], 1, device=dev)
        objs.append(e)
    table = torch.full((rows, 12), -1.0, device=dev)
    for i, e in enumerate(objs):
        table[i, 0], table[i, 1] = e["id"], e["P"]

    def one(e, kf):
        if a.impl == "ours":
            return e["step"](e["rs_ours"], e["tile_mask"], kf[0], kf[1], kf[2])[0]
        return bench.torch_mapping_iteration(e["params"], e["init"], e["opt"], e["conf"], Rast, e["rs"], e["tile_mask"], kf[0],
                                             kf[1], kf[2])

    # An object covers a few dozen tiles with deep lists: one object's blend kernels keep only a fraction of the 148 SMs
    # busy.  Objects are independent, so their steps go round-robin onto several streams and overlap on the device (the
    # reference cannot do this: its forward blocks the host twice per call).
    n_streams = max(1, a.streams) if a.impl == "ours" else 1
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else [torch.cuda.current_stream()]
    main_stream = torch.cuda.current_stream()

    def fan_out(fn):
        if n_streams == 1:
            return [fn(e) for e in objs]
        for st in streams:
            st.wait_stream(main_stream)
        res = []
        for i, e in enumerate(objs):
            with torch.cuda.stream(streams[i % n_streams]):
                res.append(fn(e))
        for st in streams:
            main_stream.wait_stream(st)
        return res

    def kernel_step():
        fan_out(lambda e: one(e, e["kf"]))

    def e2e_object(e):
        for dst, src in zip(e["slot"], e["host"]):
            dst.copy_(src, non_blocking=True)
        return one(e, e["slot"])

    def e2e_step():
        losses = fan_out(e2e_object)
        return [float(x) for x in losses]  # D2H read of every object's loss

    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    ms = bench.timed(kernel_step, a.steps, a.warmup, dist_on, sampler)
    ms_e2e = bench.timed(e2e_step, a.steps, a.warmup, dist_on, sampler)
    if a.impl == "ours":
        for e in objs:
            e["step"].check()
    sampler.stop_flag = True
    n_obj = len(counts)
    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for e in objs for t in e["host"])
        out = {"metric": METRIC, "value": n_obj * a.steps / (ms / 1000.0), "unit": UNIT, "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "c3: %d objects + background, %d Gaussians in total, SH degree 3, %dx%d RGB-D keyframe, "
                                      "per-object masked loss" % (a.objects, sum(counts), W, H),
                          "parallelism": "object-sharded x%d (LPT by Gaussian count)" % world,
                          "objects_per_rank": [sum(1 for o in owner if owner[o] == r) for r in range(world)],
                          "gaussians_per_rank": load, "load_imbalance": max(load) / (sum(load) / world),
                          "streams_per_rank": n_streams},
               "e2e": {"value": n_obj * a.steps / (ms_e2e / 1000.0), "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 * len(objs),
                       "what": "rank 0's bytes; one mapping iteration per owned object: H2D keyframe + mask, fused step "
                               "(ours) / stock torch loop around the reference rasterizer (reference), D2H loss"},
               "clocks": sampler.summary()}
        if a.impl == "reference":
            out["impl"] = "reference"
        print(json.dumps(out))
    if dist_on:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
