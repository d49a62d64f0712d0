"""Secondary measurements (BASELINE config 4 "object initialisation burst" and the mapper's small image passes):
distCUDA2 at 500k points, accumulate_gaussian_error, batched quadric init / project / refine, mask builders.
Each op is timed with CUDA events (warm-up 3, 20 repetitions) next to what the reference runs for it on the same GPU:
the unmodified extension (oracle/_ref) where one exists, otherwise the reference's own torch code restated in
oracle/ (maps_oracle / quadric_oracle) executed with torch on the GPU where possible.  Prints one JSON line.

    python tests/dev_bench_ops.py > profiles/r01_bench_secondary_ops.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import refharness as rh
from dqo_map_b200 import knn, map_utils, quadric
from oracle import maps_oracle as mo

DEV = torch.device("cuda:0")


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {}
    g = torch.Generator().manual_seed(2024)
    ref_ok = rh.reference_available()
    ref = rh.load_reference() if ref_ok else None

    # --- distCUDA2, 500k new Gaussians (GPC:538) -----------------------------------------------------------------
    pts = (torch.rand(500_000, 3, generator=g) * torch.tensor([6.0, 3.0, 6.0])).to(DEV)
    ours = timed(lambda: knn.distCUDA2(pts))
    row = {"points": 500_000, "ours_ms": ours, "alg_bytes": 500_000 * (12 * 3 + 72 + 16)}
    if ref_ok:
        ref_knn = ref[2]
        row["reference_ms"] = timed(lambda: ref_knn.distCUDA2(pts), reps=5)
        a, b = knn.distCUDA2(pts), ref_knn.distCUDA2(pts)
        row["identical"] = bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))
    out["distCUDA2_500k"] = row

    # --- accumulate_gaussian_error on a 1200x680 frame, 1M Gaussians ----------------------------------------------
    H, W, P = 680, 1200, 1_000_000
    ce, de, ne = (torch.rand(H, W, 1, generator=g).to(DEV) for _ in range(3))
    ci = torch.randint(-1, P, (H, W, 1), generator=g, dtype=torch.int32).to(DEV)
    di = torch.randint(-1, P, (H, W, 1), generator=g, dtype=torch.int32).to(DEV)
    row = {"ours_ms": timed(lambda: map_utils.accumulate_gaussian_error(H, W, P, ce, de, ne, ci, di, 0.1, 0.05, 0.3, True))}
    if ref_ok:
        ref_cu = ref[3]
        row["reference_ms"] = timed(lambda: ref_cu.accumulate_gaussian_error(H, W, P, ce, de, ne, ci, di, 0.1, 0.05, 0.3, True))
    out["accumulate_gaussian_error_1200x680_1M"] = row

    # --- mask builders at 1200x680 (reference: the torch op chains of SLAM/utils.py on the same GPU) -----------------
    T_map = torch.where(torch.rand(1, H, W, generator=g) > 0.3, torch.rand(1, H, W, generator=g), torch.ones(1, H, W)).to(DEV)
    render = (torch.rand(3, H, W, generator=g).to(DEV)) * (T_map != 1)
    gt = torch.rand(3, H, W, generator=g).to(DEV)
    ro = {"T_map": T_map, "render": render, "semantic_seg": None}
    out["evaluate_render_range_local"] = {"ours_ms": timed(lambda: map_utils.evaluate_render_range(ro)),
                                          "reference_torch_ms": timed(lambda: mo.evaluate_render_range(ro))}

    def ref_topk():
        # colorerror2tilemask builds its mask with zeros(dtype int32) on the error's device in the reference
        err = mo.color_error_map(render, gt)
        down = mo.tile_means(err, 16)
        _, idx = torch.topk(down.reshape(-1), k=int(down.numel() * 0.4))
        m = torch.zeros(down.numel(), dtype=torch.int32, device=DEV)
        m[idx] = 1
        return m

    out["evaluate_render_range_global_topk"] = {
        "ours_ms": timed(lambda: map_utils.evaluate_render_range(ro, gt_image=gt, global_opt=True, sample_ratio=0.4)),
        "reference_torch_ms": timed(ref_topk)}

    # --- quadric batch: 64 objects x 8 views, the 20-iteration refinement (QUAD:2234-2298) ---------------------------
    gq = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quadric.npz"))
    n0, V = gq["obs_bboxes"].shape[:2]
    reps = (64 + n0 - 1) // n0
    tile = lambda a: np.concatenate([a] * reps, axis=0)[:64]
    ob, Ps = torch.tensor(tile(gq["obs_bboxes"])).to(DEV), torch.tensor(tile(gq["Ps"])).to(DEV)
    ax, R, c, vc = tile(gq["init_axes"]), tile(gq["init_R"]), tile(gq["init_center"]), tile(gq["view_choice"])
    nv = np.full(64, V, np.int32)
    row = {"objects": 64, "views": int(V), "iters": 20,
           "ours_ms": timed(lambda: quadric.quadric_refine(ax, R, c, ob, Ps, nv, vc), reps=10)}
    # the reference optimises one object at a time with torch autograd on the GPU: ~20 iterations x ~60 small kernels
    # each; its cost is measured through the fp32 torch restatement (oracle/quadric_oracle.py) on the host for ONE object
    from oracle import quadric_oracle as qo
    import time
    t0 = time.perf_counter()
    qo.refine(gq["init_axes"][0], gq["init_R"][0], gq["init_center"][0], gq["obs_bboxes"][0], gq["Ps"][0], gq["view_choice"][0])
    row["torch_restatement_cpu_ms_per_object"] = (time.perf_counter() - t0) * 1e3
    out["quadric_refine_64_objects"] = row

    # --- tracker front-end at 1200x680: pyramids + 3 x 5 ICP iterations (icp.py:424-441) -----------------------------
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_icp_golden import synth_depth
    from dqo_map_b200 import icp
    from oracle import icp_oracle as io
    Kn = np.array([[600.0, 0, 599.5], [0, 600.0, 339.5], [0, 0, 1.0]])
    p1 = np.eye(4)
    p1[:3, 3] = [0.02, -0.01, 0.015]
    d0 = torch.from_numpy(synth_depth(680, 1200, Kn, np.eye(4))).to(DEV).view(680, 1200, 1)
    d1 = torch.from_numpy(synth_depth(680, 1200, Kn, p1)).to(DEV).view(680, 1200, 1)
    Kt = torch.from_numpy(Kn).float().to(DEV)
    out["tracker_predict_pose_1200x680"] = {
        "ours_ms": timed(lambda: icp.predict_pose(d0, d1, Kt), reps=10),
        "reference_torch_ms": timed(lambda: io.predict_pose(d0, d1, Kt.clone()), reps=5),
        "what": "depth max-pyramid, vertex + normal maps of both frames, 3 levels x 5 Gauss-Newton iterations; the torch "
                "figure is oracle/icp_oracle.py (the reference's op sequence, host-side 6x6 inverse included) on the same GPU"}

    print(json.dumps(out))


if __name__ == "__main__":
    main()
