#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference operators (oracle/_ref, built by
oracle/build_ref.py from /root/reference) on seeded synthetic inputs.  Must run on a GPU box:

    gpurun -- python tests/golden/make_golden.py --out gpurun_out/golden [--compare]

The fixtures hold the inputs and every reference output / decoded private buffer needed to pin the CPU oracle
(oracle/) and the CUDA path: binning artefacts (keys, sorted ids, ranges, n_contrib, radii, tiles_touched) and
float images / gradients.  `--compare` additionally prints a differential report of this repository's CUDA path.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness as rh  # noqa: E402

CASES = [
    # name, cfg, P, sh_degree, mask, precomp
    ("tiny_sh0_full", "tiny", 2000, 0, "ones", False),
    ("tiny_sh3_half", "tiny", 3000, 3, "half", False),
    ("tiny_precomp", "tiny", 1500, 0, "ones", True),
]


def t2n(t):
    return t.detach().cpu().numpy()


def run_reference(inp):
    rast, C, _, _ = rh.load_reference()
    args = rh.raster_args(inp)
    fwd = C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    cam = inp["cam"]
    H, W = cam.image_height, cam.image_width
    gc, gd = rh.make_pixel_grads(H, W, inp["xyz"].device)
    bwd = C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd))
    torch.cuda.synchronize()
    return fwd, bwd, gc, gd


def pack_case(inp, fwd, bwd, gc, gd):
    (rendered, tile_num, color, depth, hit_color, hit_depth, hcw, hdw, T_map, radii, geom, binning, img, tile_indices,
     n_touched) = fwd
    cam = inp["cam"]
    H, W = cam.image_height, cam.image_width
    P = inp["xyz"].shape[0]
    dec = rh.decode_ref_buffers(geom, binning, img, P, rendered, W, H)
    vis = t2n(radii) > 0
    d = {
        "xyz": t2n(inp["xyz"]), "scales": t2n(inp["scales"]), "rotations": t2n(inp["rotations"]),
        "opacity": t2n(inp["opacity"]), "shs": t2n(inp["shs"]), "rgb_in": t2n(inp["rgb"]),
        "sh_degree": np.int32(inp["sh_degree"]), "precomp": np.int32(inp["precomp"]), "bg": t2n(inp["bg"]),
        "tile_mask": t2n(inp["tile_mask"]), "viewmatrix": t2n(cam.world_view_transform),
        "projmatrix": t2n(cam.full_proj_transform), "campos": t2n(cam.camera_center),
        "W": np.int32(W), "H": np.int32(H), "tanfovx": np.float64(cam.tanfovx), "tanfovy": np.float64(cam.tanfovy),
        "cx": np.float64(cam.cx), "cy": np.float64(cam.cy),
        "num_rendered": np.int64(rendered), "tile_num": np.int64(tile_num),
        "color": t2n(color), "depth": t2n(depth), "hit_color": t2n(hit_color), "hit_depth": t2n(hit_depth),
        "hit_color_weight": t2n(hcw), "hit_depth_weight": t2n(hdw), "T_map": t2n(T_map), "radii": t2n(radii),
        "n_touched": t2n(n_touched), "tile_indices": t2n(tile_indices)[: max(int(tile_num), 1)],
        "keys_sorted": dec["keys_sorted"], "point_list": dec["point_list"], "ranges": dec["ranges"],
        "n_contrib": dec["n_contrib"], "accum_alpha": dec["accum_alpha"], "tiles_touched": dec["tiles_touched"],
        # per-Gaussian float state is only defined where radii > 0 (uninitialised elsewhere): zero the rest
        "means2D": np.where(vis[:, None], dec["means2D"], 0), "depths": np.where(vis, dec["depths"], 0),
        "conic_opacity": np.where(vis[:, None], dec["conic_opacity"], 0),
        "rgb": np.where(vis[:, None], dec["rgb"], 0), "cov3D": np.where(vis[:, None], dec["cov3D"], 0),
        "clamped": np.where(vis[:, None], dec["clamped"], 0),
        "grad_color": t2n(gc), "grad_depth": t2n(gd),
    }
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for n, t in zip(names, bwd):
        d[n] = t2n(t)
    return d


def compare(inp, fwd, bwd, gc, gd, name):
    from dqo_map_b200 import rasterizer as ours
    (rendered, tile_num, color, depth, hit_color, hit_depth, hcw, hdw, T_map, radii, geom, binning, img, tile_indices,
     n_touched) = fwd
    cam = inp["cam"]
    H, W = cam.image_height, cam.image_width
    P = inp["xyz"].shape[0]
    ref = rh.decode_ref_buffers(geom, binning, img, P, rendered, W, H)
    o = ours.rasterize_gaussians(*rh.raster_args(inp))
    torch.cuda.synchronize()
    st = o[10]._dqo_state
    mine = rh.export_ours(st, P, W, H)
    print("== %s: P=%d R(ref)=%d R(ours)=%d tiles ref=%d ours=%d" % (name, P, rendered, o[0], tile_num, o[1]))
    vis = t2n(radii) > 0

    def cmp_int(label, a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            print("  %-18s SHAPE %s vs %s" % (label, a.shape, b.shape))
            return
        print("  %-18s mismatches %d / %d" % (label, int((a != b).sum()), a.size))

    def cmp_f(label, a, b, mask=None):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if mask is not None:
            a, b = a[mask], b[mask]
        if a.size == 0:
            print("  %-18s empty" % label)
            return
        bit = int((a.astype(np.float32).view(np.uint32) != b.astype(np.float32).view(np.uint32)).sum())
        print("  %-18s max|d| %.3e  rel %.3e  bit-mismatch %d / %d" % (
            label, np.abs(a - b).max(), np.abs(a - b).max() / (np.abs(a).max() + 1e-30), bit, a.size))

    cmp_int("radii", t2n(radii), t2n(o[9]))
    cmp_int("tiles_touched", ref["tiles_touched"], mine["tiles_touched"])
    cmp_f("depths", ref["depths"], mine["depths"], vis)
    cmp_f("means2D", ref["means2D"], mine["means2D"], vis)
    cmp_f("conic_opacity", ref["conic_opacity"], mine["conic_opacity"], vis)
    cmp_f("rgb", ref["rgb"], mine["rgb"], vis)
    cmp_int("keys_sorted", ref["keys_sorted"], mine["keys_sorted"])
    cmp_int("point_list", ref["point_list"], mine["point_list"])
    cmp_int("ranges", ref["ranges"], mine["ranges"])
    cmp_int("tile_indices", t2n(tile_indices)[:tile_num], t2n(o[13])[: o[1]])
    # n_contrib / final_T are only defined inside rendered tiles
    rendered_px = t2n(T_map)[0] != 1.0
    nc_ref = np.where(rendered_px, ref["n_contrib"], 0)
    nc_our = np.where(rendered_px, mine["n_contrib"], 0)
    cmp_int("n_contrib", nc_ref, nc_our)
    cmp_int("hit_depth idx", t2n(hit_depth), t2n(o[5]))
    cmp_int("hit_color idx", t2n(hit_color), t2n(o[4]))
    cmp_int("n_touched", t2n(n_touched), t2n(o[14]))
    cmp_f("color", t2n(color), t2n(o[2]))
    cmp_f("depth", t2n(depth), t2n(o[3]))
    cmp_f("T_map", t2n(T_map), t2n(o[8]))
    cmp_f("hit_color_weight", t2n(hcw), t2n(o[6]))
    cmp_f("hit_depth_weight", t2n(hdw), t2n(o[7]))
    ob = ours.rasterize_gaussians_backward(*rh.backward_args(inp, o, gc, gd))
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    # the reference's own run-to-run noise (float atomics in unspecified order) is the floor for any comparison
    _, C, _, _ = rh.load_reference()
    bwd2 = C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd))
    torch.cuda.synchronize()
    for n, a, b, a2 in zip(names, bwd, ob, bwd2):
        a, b = t2n(a).astype(np.float64), t2n(b).astype(np.float64).reshape(t2n(a).shape)
        a2 = t2n(a2).astype(np.float64)
        if a.size == 0:
            continue
        nrm = np.linalg.norm(a - b) / (np.linalg.norm(a) + 1e-30)
        floor = np.linalg.norm(a - a2) / (np.linalg.norm(a) + 1e-30)
        print("  %-18s rel-norm %.3e (ref-vs-ref %.3e)  max|d| %.3e  max|ref| %.3e" % (
            n, nrm, floor, np.abs(a - b).max(), np.abs(a).max()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(rh.ROOT, "gpurun_out", "golden"))
    ap.add_argument("--compare", action="store_true")
    ap.add_argument("--big", action="store_true", help="also compare (not store) the c1/c2 configurations")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, cfg, P, deg, mask, precomp in CASES:
        inp = rh.make_inputs(cfg, dev, P=P, sh_degree=deg, mask=mask, precomp=precomp)
        fwd, bwd, gc, gd = run_reference(inp)
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **pack_case(inp, fwd, bwd, gc, gd))
        print("wrote", name, "R =", fwd[0], "tiles =", fwd[1])
        if a.compare:
            compare(inp, fwd, bwd, gc, gd, name)
    # kNN and error-scatter fixtures
    _, _, knn_C, cu_C = rh.load_reference()
    g = torch.Generator(device="cpu").manual_seed(5)
    pts = torch.cat([torch.rand(3000, 3, generator=g) * 4 - 1, torch.rand(1000, 3, generator=g) * 0.1 + 2.0])
    pts[100:110] = pts[100]  # exact duplicates: distance ties
    md, ki = knn_C.distCUDA2(pts.to(dev))
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(a.out, "knn_4000.npz"), points=t2n(pts), mean_dist2=t2n(md), knn_idx=t2n(ki))
    H, W, P = 48, 64, 500
    ce, de, ne = (torch.rand(H, W, 1, generator=g) for _ in range(3))
    ci = torch.randint(-1, P + 5, (H, W, 1), generator=g, dtype=torch.int32)
    di = torch.randint(-1, P, (H, W, 1), generator=g, dtype=torch.int32)
    fix = {"ce": t2n(ce), "de": t2n(de), "ne": t2n(ne), "ci": t2n(ci), "di": t2n(di), "P": np.int32(P)}
    for cm in (True, False):
        outs = cu_C.accumulate_gaussian_error(H, W, P, ce.to(dev), de.to(dev), ne.to(dev), ci.to(dev), di.to(dev), 0.5,
                                              0.6, 0.7, cm)
        torch.cuda.synchronize()
        for k, t in zip(["color", "depth", "normal", "rescale"], outs):
            fix["%s_%s" % (k, "max" if cm else "mean")] = t2n(t)
    np.savez_compressed(os.path.join(a.out, "accum_error.npz"), **fix)
    print("wrote knn_4000, accum_error")
    if a.big:
        for cfg in ("c1", "c2"):
            inp = rh.make_inputs(cfg, dev)
            fwd, bwd, gc, gd = run_reference(inp)
            compare(inp, fwd, bwd, gc, gd, cfg)


if __name__ == "__main__":
    main()
