"""Dev tool: run-to-run and implementation-to-implementation spread of the mapping-loop metrics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refharness as rh
import test_gpu_mapping as tm
rast_pkg = rh.load_reference()[0] if rh.reference_available() else None
for iters in (50, 200):
    gt, cam, settings, raw, gt_color, gt_depth, render_mask = tm._scene()
    for mode in ["fused", "fused", "ours_torch", "ours_torch", "reference", "reference"]:
        if mode == "reference" and rast_pkg is None:
            continue
        p, d, _, _ = tm._loop(raw, settings, gt["tile_mask"], gt_color, gt_depth, render_mask, iters, mode, rast_pkg)
        print("iters %3d %-10s psnr %.4f depthL1 %.6f" % (iters, mode, p, d), flush=True)
