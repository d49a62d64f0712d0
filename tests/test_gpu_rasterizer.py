"""GPU parity tests of the rasterizer hot path, all through the C-ABI (dqo_rast_forward / dqo_rast_backward).

Gates (BASELINE.json north_star): sort keys, sorted ids, tile ranges, tile list, per-pixel contributor counts, radii and
index maps bit-exact; colour / depth / T within max-abs 1e-4; gradients within 1e-3 relative.
Three arbiters: (1) golden fixtures produced by the reference on a B200, (2) the CPU oracle on fresh seeded inputs,
(3) the live reference extension (oracle/_ref) when its .so travelled to the box.
"""
import os

import numpy as np
import pytest
import torch

import refharness as rh
from dqo_map_b200 import _lib, rasterizer, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRADS = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]


def t2n(t):
    return t.detach().cpu().numpy()


def _inputs_from_golden(g):
    dev = torch.device(DEV)
    d = {k: torch.tensor(g[k]).to(dev) for k in ["xyz", "scales", "rotations", "opacity", "shs", "bg", "tile_mask"]}
    d["rgb"] = torch.tensor(g["rgb_in"]).to(dev)
    d["sh_degree"] = int(g["sh_degree"])
    d["precomp"] = bool(int(g["precomp"]))
    cam = synthetic.make_camera("tiny").to(dev)
    assert np.array_equal(t2n(cam.world_view_transform), g["viewmatrix"])
    assert np.array_equal(t2n(cam.full_proj_transform), g["projmatrix"])
    d["cam"] = cam
    return d


def _run_ours(inp, gc=None, gd=None, binning=("single",)):
    rasterizer.set_binning_mode(*binning)
    o = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
    cam = inp["cam"]
    st = o[10]._dqo_state
    ex = rh.export_ours(st, inp["xyz"].shape[0], cam.image_width, cam.image_height)
    bw = None
    if gc is not None:
        bw = rasterizer.rasterize_gaussians_backward(*rh.backward_args(inp, o, gc, gd))
        torch.cuda.synchronize()
    return o, ex, bw


def _check_forward_exact(o, ex, ref, P):
    """ref: dict in golden-fixture format."""
    (rendered, tile_num, color, depth, hit_color, hit_depth, hcw, hdw, T_map, radii, *_rest) = o
    tile_indices, n_touched = o[13], o[14]
    assert rendered == int(ref["num_rendered"]) and tile_num == int(ref["tile_num"])
    assert np.array_equal(t2n(radii), ref["radii"])
    assert np.array_equal(ex["tiles_touched"], ref["tiles_touched"])
    if ex["keys_sorted"] is not None:  # single-phase binning: the full sorted list exists
        assert np.array_equal(ex["keys_sorted"], ref["keys_sorted"])
        assert np.array_equal(ex["point_list"], ref["point_list"])
        assert np.array_equal(ex["ranges"], ref["ranges"])
    assert np.array_equal(t2n(tile_indices)[:tile_num], ref["tile_indices"][:tile_num])
    rendered_px = ref["T_map"][0] != 1.0
    assert np.array_equal(np.where(rendered_px, ex["n_contrib"], 0), np.where(rendered_px, ref["n_contrib"], 0))
    assert np.array_equal(t2n(hit_depth), ref["hit_depth"])
    assert np.array_equal(t2n(hit_color), ref["hit_color"])
    assert np.array_equal(t2n(n_touched), ref["n_touched"])
    vis = ref["radii"] > 0
    for k in ["means2D", "depths", "conic_opacity"]:
        assert np.array_equal(ex[k][vis].view(np.uint32), np.ascontiguousarray(ref[k][vis]).view(np.uint32)), k
    for name, t in [("color", color), ("depth", depth), ("T_map", T_map), ("hit_color_weight", hcw),
                    ("hit_depth_weight", hdw)]:
        assert np.abs(t2n(t) - ref[name]).max() <= 1e-4, name


def _check_grads(bw, ref, tol=1e-3, ref2=None):
    """Norm-wise relative gate of 1e-3 (north_star).  `ref2` = a second run of the same reference: its float atomics
    land in unspecified order and the conic->cov3D chain amplifies the last-bit differences, so at 1M Gaussians two
    reference runs differ from EACH OTHER by up to ~7e-4 in dL_drotations (profiles/r01_gradient_parity_vs_reference.log).
    Where that floor is known the gate is max(1e-3, 3 x floor); the float64 arbiter test below is the sharper gate."""
    for i, (name, t) in enumerate(zip(GRADS, bw)):
        b = np.asarray(ref[name], dtype=np.float64)
        if b.size == 0:
            continue
        a = t2n(t).astype(np.float64).reshape(b.shape)
        rel = np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
        gate = tol
        if ref2 is not None:
            # the noise is heavy-tailed (one badly conditioned Gaussian can dominate a run): take the largest of several
            # repeated reference runs as the floor
            runs = ref2 if isinstance(ref2, (list, tuple)) else [ref2]
            floor = max(np.linalg.norm(np.asarray(r[name], dtype=np.float64) - b) / (np.linalg.norm(b) + 1e-30) for r in runs)
            gate = max(tol, 3.0 * floor)
        assert rel <= gate, (name, rel, gate)


@pytest.mark.parametrize("case", ["tiny_sh0_full", "tiny_sh3_half", "tiny_precomp"])
def test_against_reference_golden(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    inp = _inputs_from_golden(g)
    gc, gd = torch.tensor(g["grad_color"]).to(DEV), torch.tensor(g["grad_depth"]).to(DEV)
    o, ex, bw = _run_ours(inp, gc, gd)
    _check_forward_exact(o, ex, g, inp["xyz"].shape[0])
    # against the reference on the GPU the images are in fact bit-identical (same expf)
    for name, idx in [("color", 2), ("depth", 3), ("T_map", 8)]:
        assert np.array_equal(t2n(o[idx]).view(np.uint32), g[name].view(np.uint32)), name
    _check_grads(bw, g)


def _oracle_as_ref(inp):
    cam = inp["cam"]
    sc = oracle.Scene(t2n(inp["xyz"]), t2n(inp["scales"]), t2n(inp["rotations"]), t2n(inp["opacity"]),
                      t2n(cam.world_view_transform), t2n(cam.full_proj_transform), t2n(cam.camera_center),
                      cam.image_width, cam.image_height, cam.tanfovx, cam.tanfovy, cam.cx, cam.cy, t2n(inp["bg"]),
                      t2n(inp["tile_mask"]), shs=None if inp["precomp"] else t2n(inp["shs"]),
                      sh_degree=0 if inp["precomp"] else inp["sh_degree"],
                      colors_precomp=t2n(inp["rgb"]) if inp["precomp"] else None,
                      normal_threshold=synthetic.RENDER_DEFAULTS["normal_threshold"])
    oracle.set_threads(os.cpu_count() or 1)
    pre, bn, img = oracle.forward(sc)
    ref = dict(num_rendered=bn["num_rendered"], tile_num=bn["tile_num"], radii=pre["radii"],
               tiles_touched=pre["tiles_touched"], keys_sorted=bn["keys_sorted"], point_list=bn["point_list"],
               ranges=bn["ranges"], tile_indices=bn["tile_indices"], n_contrib=img["n_contrib"], T_map=img["T_map"],
               hit_depth=img["hit_depth"], hit_color=img["hit_color"], n_touched=img["n_touched"],
               means2D=pre["means2D"], depths=pre["depths"], conic_opacity=pre["conic_opacity"], color=img["color"],
               depth=img["depth"], hit_color_weight=img["hit_color_weight"], hit_depth_weight=img["hit_depth_weight"])
    return sc, pre, bn, img, ref


@pytest.mark.parametrize("cfg,P,deg,mask", [("small", 20000, 3, "ones"), ("small", 12000, 1, "half"), ("c1", 30000, 0, "ones"),
                                            ("ragged", 15000, 2, "ones"), ("ragged", 9000, 2, "half")])
def test_against_cpu_oracle(cfg, P, deg, mask):
    inp = rh.make_inputs(cfg, torch.device(DEV), seed=77, P=P, sh_degree=deg, mask=mask)
    cam = inp["cam"]
    gc, gd = rh.make_pixel_grads(cam.image_height, cam.image_width, DEV, seed=5)
    o, ex, bw = _run_ours(inp, gc, gd)
    sc, pre, bn, img, ref = _oracle_as_ref(inp)
    # the CPU cannot reproduce MUFU.EX2 bit-exactly: allow the (unobserved so far) alpha-threshold flips on a
    # vanishing fraction of pixels for the quantities downstream of expf; everything upstream must be exact
    assert o[0] == ref["num_rendered"] and o[1] == ref["tile_num"]
    assert np.array_equal(t2n(o[9]), ref["radii"])
    assert np.array_equal(ex["keys_sorted"], ref["keys_sorted"])
    assert np.array_equal(ex["point_list"], ref["point_list"])
    assert np.array_equal(ex["ranges"], ref["ranges"])
    rendered_px = img["T_map"][0] != 1.0
    mism = (np.where(rendered_px, ex["n_contrib"], 0) != np.where(rendered_px, img["n_contrib"], 0)).mean()
    assert mism <= 1e-4
    assert (t2n(o[5]) != img["hit_depth"]).mean() <= 1e-4
    good = (t2n(o[5]) == img["hit_depth"])[0] & (np.where(rendered_px, ex["n_contrib"], 0) == np.where(rendered_px, img["n_contrib"], 0))
    assert np.abs(t2n(o[2]) - img["color"])[:, good].max() <= 1e-4
    assert np.abs(t2n(o[3]) - img["depth"])[:, good].max() <= 1e-4
    assert np.abs(t2n(o[8]) - img["T_map"])[:, good].max() <= 1e-4
    gr = oracle.backward(sc, pre, bn, img, t2n(gc), t2n(gd))
    _check_grads(bw, gr)


@pytest.mark.skipif(not rh.reference_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg,mask,precomp", [("c1", "ones", False), ("c1", "half", True), ("c2", "ones", False),
                                              ("ragged", "half", False), ("ragged", "ones", False), ("deg1", "ones", False),
                                              ("c5", "ones", False)])
def test_against_live_reference(cfg, mask, precomp):
    """Full BASELINE sizes (config 1, 2 and 5: 3 M Gaussians at 1920x1080, ~10^8 instances) against the unmodified
    reference extension on the same GPU."""
    _, C, _, _ = rh.load_reference()
    inp = rh.make_inputs(cfg, torch.device(DEV), mask=mask, precomp=precomp)
    cam = inp["cam"]
    H, W, P = cam.image_height, cam.image_width, inp["xyz"].shape[0]
    gc, gd = rh.make_pixel_grads(H, W, DEV)
    fwd = C.rasterize_gaussians(*rh.raster_args(inp))
    bwd = C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd))
    torch.cuda.synchronize()
    dec = rh.decode_ref_buffers(fwd[10], fwd[11], fwd[12], P, fwd[0], W, H)
    ref = dict(num_rendered=fwd[0], tile_num=fwd[1], radii=t2n(fwd[9]), tile_indices=t2n(fwd[13]), color=t2n(fwd[2]),
               depth=t2n(fwd[3]), hit_color=t2n(fwd[4]), hit_depth=t2n(fwd[5]), hit_color_weight=t2n(fwd[6]),
               hit_depth_weight=t2n(fwd[7]), T_map=t2n(fwd[8]), n_touched=t2n(fwd[14]), **dec)
    for n, t in zip(GRADS, bwd):
        ref[n] = t2n(t)
    ref2 = [{n: t2n(t) for n, t in zip(GRADS, C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd)))}
            for _ in range(3)]
    o, ex, bw = _run_ours(inp, gc, gd)
    _check_forward_exact(o, ex, ref, P)
    assert np.array_equal(t2n(o[2]).view(np.uint32), ref["color"].view(np.uint32))
    _check_grads(bw, ref, ref2=ref2)


@pytest.mark.skipif(not rh.reference_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg,mask,front,back", [("c2", "ones", 2_000_128, 4_000_000), ("c1", "half", 65_536, 1_200_000),
                                                 ("c5", "ones", 6_000_128, 30_000_000)])
def test_two_phase_binning_against_live_reference(cfg, mask, front, back):
    """Occlusion-aware two-phase binning (front_instances > 0) against the unmodified reference: every per-pixel and
    per-Gaussian output the reference returns must still be bit-identical although most instances are never binned."""
    _, C, _, _ = rh.load_reference()
    inp = rh.make_inputs(cfg, torch.device(DEV), mask=mask)
    cam = inp["cam"]
    H, W, P = cam.image_height, cam.image_width, inp["xyz"].shape[0]
    gc, gd = rh.make_pixel_grads(H, W, DEV)
    fwd = C.rasterize_gaussians(*rh.raster_args(inp))
    bwd = C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd))
    torch.cuda.synchronize()
    dec = rh.decode_ref_buffers(fwd[10], fwd[11], fwd[12], P, fwd[0], W, H)
    ref = dict(num_rendered=fwd[0], tile_num=fwd[1], radii=t2n(fwd[9]), tile_indices=t2n(fwd[13]), color=t2n(fwd[2]),
               depth=t2n(fwd[3]), hit_color=t2n(fwd[4]), hit_depth=t2n(fwd[5]), hit_color_weight=t2n(fwd[6]),
               hit_depth_weight=t2n(fwd[7]), T_map=t2n(fwd[8]), n_touched=t2n(fwd[14]), **dec)
    for n, t in zip(GRADS, bwd):
        ref[n] = t2n(t)
    ref2 = [{n: t2n(t) for n, t in zip(GRADS, C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd)))}
            for _ in range(3)]
    o, ex, bw = _run_ours(inp, gc, gd, binning=("fixed", front, back))
    st = o[10]._dqo_state.status_host
    assert 0 < st[_lib.ST_R_FRONT] <= front and st[_lib.ST_R_FRONT] + st[_lib.ST_R_BACK] < fwd[0]
    _check_forward_exact(o, ex, ref, P)
    for name, idx in [("color", 2), ("depth", 3), ("T_map", 8), ("hit_color_weight", 6), ("hit_depth_weight", 7)]:
        assert np.array_equal(t2n(o[idx]).view(np.uint32), ref[name].view(np.uint32)), name
    _check_grads(bw, ref, ref2=ref2)


@pytest.mark.parametrize("cfg,mask,precomp", [("c2", "ones", False), ("c1", "half", True), ("small", "ones", False)])
def test_two_phase_binning_equals_single_phase(cfg, mask, precomp):
    """Front regions from 'almost nothing' (every tile unfinished, whole lists in the back phase) to 'almost everything'
    (back phase empty) give bit-identical forward outputs and gradients equal up to atomic ordering."""
    inp = rh.make_inputs(cfg, torch.device(DEV), mask=mask, precomp=precomp)
    cam = inp["cam"]
    H, W = cam.image_height, cam.image_width
    gc, gd = rh.make_pixel_grads(H, W, DEV)
    o1, ex1, bw1 = _run_ours(inp, gc, gd)
    # run-to-run noise of the float atomics (heavy-tailed, see _check_grads): the largest of three repeated runs
    floor = {n: 0.0 for n in GRADS}
    for _ in range(3):
        _, _, bwr = _run_ours(inp, gc, gd)
        for n, a, b in zip(GRADS, bw1, bwr):
            if a.numel():
                floor[n] = max(floor[n], float((a - b).norm() / (a.norm() + 1e-30)))
    R = o1[0]
    tested = 0
    for frac in (0.002, 0.05, 0.3, 0.98):
        front = max(256, int(R * frac) // 256 * 256)
        o2, ex2, bw2 = _run_ours(inp, gc, gd, binning=("fixed", front, R + 1024))
        st = o2[10]._dqo_state.status_host
        assert st[_lib.ST_R_FRONT] <= front
        assert o2[0] == o1[0] and o2[1] == o1[1]
        for i in (2, 3, 4, 5, 6, 7, 8, 9, 13, 14):
            assert torch.equal(o1[i], o2[i]), (frac, i)
        assert np.array_equal(ex1["n_contrib"], ex2["n_contrib"])
        assert np.array_equal(ex1["accum_alpha"].view(np.uint32), ex2["accum_alpha"].view(np.uint32))
        for name, a, b in zip(GRADS, bw1, bw2):
            if a.numel() == 0:
                continue
            rel = float((a - b).norm() / (a.norm() + 1e-30))
            assert rel <= max(1e-3, 3.0 * floor[name]), (frac, name, rel, floor[name])  # same kernels, other atomic order
        tested += 1
    assert tested == 4
    rasterizer.set_binning_mode("single")


def test_auto_binning_policy_converges_and_keeps_results():
    """Default mode on a heavily occluded scene: first call single-phase, then two-phase with a safe back region, then a
    tight one; outputs never change."""
    inp = rh.make_inputs("c2", torch.device(DEV))
    rasterizer.set_binning_mode("auto")
    outs, stats = [], []
    for _ in range(4):
        o = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
        st = o[10]._dqo_state
        outs.append(o)
        stats.append((st.settings.front_instances, st.settings.back_instances, list(st.status_host)))
    assert stats[0][0] == 0 and stats[0][2][_lib.ST_WALKED] < 0.35 * stats[0][2][_lib.ST_NUM_RENDERED]
    assert stats[1][0] > 0 and stats[1][0] + stats[1][1] >= stats[0][2][_lib.ST_NUM_RENDERED]
    assert stats[3][0] > 0 and stats[3][0] + stats[3][1] < 0.5 * stats[0][2][_lib.ST_NUM_RENDERED]
    for o in outs[1:]:
        assert o[0] == outs[0][0] and o[1] == outs[0][1]
        for i in (2, 3, 4, 5, 6, 7, 8, 9, 13, 14):
            assert torch.equal(outs[0][i], o[i]), i
    rasterizer.set_binning_mode("single")


@pytest.mark.skipif(not rh.reference_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("binning", ["single", "two_phase"])
def test_more_than_65535_tiles_uses_32bit_keys(binning):
    """Images with >= 65535 tiles switch the tile-id sort keys from u16 to u32 (both binning modes)."""
    _, C, _, _ = rh.load_reference()
    inp = rh.make_inputs("huge", torch.device(DEV))
    inp["scales"] = inp["scales"] * 3.0
    cam = inp["cam"]
    H, W, P = cam.image_height, cam.image_width, inp["xyz"].shape[0]
    assert ((H + 15) // 16) * ((W + 15) // 16) >= 65535
    gc, gd = rh.make_pixel_grads(H, W, DEV)
    fwd = C.rasterize_gaussians(*rh.raster_args(inp))
    bwd = C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd))
    bwd2 = C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd))
    mode = ("single",) if binning == "single" else ("fixed", max(256, (fwd[0] // 8) // 256 * 256), fwd[0] + 1024)
    o, ex, bw = _run_ours(inp, gc, gd, binning=mode)
    assert o[0] == fwd[0] and o[1] == fwd[1]
    for i in (2, 3, 4, 5, 6, 7, 8, 9, 14):
        assert torch.equal(o[i], fwd[i]), i
    assert torch.equal(o[13][:o[1]], fwd[13][:fwd[1]])
    # this scene (3x scales at f = 2000) makes the conic -> cov3D chain so ill-conditioned that two runs of the reference
    # differ by > 1e-3 in dL_dcov3D / dL_dscales / dL_drotations, with a heavy tail: gate the well-conditioned gradients
    # at 1e-3 and the chain at 5 x the reference's own noise
    ref = {n: t2n(t) for n, t in zip(GRADS, bwd)}
    ref2 = {n: t2n(t) for n, t in zip(GRADS, bwd2)}
    for name, t in zip(GRADS, bw):
        b = ref[name].astype(np.float64)
        if b.size == 0:
            continue
        a = t2n(t).astype(np.float64).reshape(b.shape)
        rel = np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
        floor = np.linalg.norm(ref2[name].astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30)
        chain = name in ("dL_dcov3D", "dL_dscales", "dL_drotations")
        assert rel <= (max(1e-3, 5.0 * floor) if chain else 1e-3), (name, rel, floor)


def test_full_size_properties():
    """Size-independent properties at BASELINE config 2 (1M Gaussians, 1200x680)."""
    inp = rh.make_inputs("c2", torch.device(DEV))
    cam = inp["cam"]
    H, W, P = cam.image_height, cam.image_width, inp["xyz"].shape[0]
    o1, ex1, _ = _run_ours(inp)
    o2, ex2, _ = _run_ours(inp)
    R = o1[0]
    assert R == int(ex1["tiles_touched"].sum())                       # every counted tile emits exactly one instance
    k = ex1["keys_sorted"]
    assert np.all(k[1:] >= k[:-1])                                     # sortedness
    same = k[1:] == k[:-1]
    assert np.all(ex1["point_list"][1:][same] > ex1["point_list"][:-1][same])  # stability (ties in index order)
    rg = ex1["ranges"].astype(np.int64)
    ne = rg[:, 0] != rg[:, 1]
    assert int((rg[ne, 1] - rg[ne, 0]).sum()) == R                     # ranges partition the list
    assert np.array_equal((k >> np.uint64(32)).astype(np.int64)[rg[ne, 0]], np.nonzero(ne)[0])
    assert o1[1] == int(ne.sum())
    # idempotence / determinism of everything but float atomics
    for i in (2, 3, 4, 5, 6, 7, 8, 9, 14):
        assert torch.equal(o1[i], o2[i])
    assert np.array_equal(ex1["point_list"], ex2["point_list"])
    # transmittance stays in (0, 1]; pixels outside rendered tiles keep the fill values (N1)
    T = t2n(o1[8])
    assert T.max() <= 1.0 and T.min() > 0.0
    hit = t2n(o1[5])
    assert hit.max() < P and hit.min() >= -1


def test_edge_cases():
    dev = torch.device(DEV)
    cam = synthetic.make_camera("tiny").to(dev)
    H, W = cam.image_height, cam.image_width
    th, tw = (H + 15) // 16, (W + 15) // 16

    def settings(deg=0):
        return rasterizer.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3, device=dev),
            scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=deg,
            campos=cam.camera_center, opaque_threshold=0.6, normal_threshold=0.5, depth_threshold=1.0,
            prefiltered=False, debug=True, cx=cam.cx, cy=cam.cy)

    ones = torch.ones(th, tw, dtype=torch.int32, device=dev)
    # P == 0 (rasterize_points.cu:103): pre-filled outputs
    r = rasterizer.GaussianRasterizer(settings())
    out = r(means3D=torch.zeros(0, 3, device=dev), opacities=torch.zeros(0, 1, device=dev),
            colors_precomp=torch.zeros(0, 3, device=dev), scales=torch.zeros(0, 3, device=dev),
            rotations=torch.zeros(0, 4, device=dev), tile_mask=ones)
    assert out[0].shape == (3, H, W) and float(out[0].abs().max()) == 0 and float(out[6].min()) == 1.0
    assert int(out[3].abs().max()) == 0 and out[8].numel() == 0
    # all Gaussians culled
    g = synthetic.make_gaussians("tiny", P=500)
    xyz = g["xyz"].to(dev).clone()
    far = xyz + 1000.0
    out = r(means3D=far, opacities=g["opacity"].to(dev), colors_precomp=g["rgb"].to(dev), scales=g["scales"].to(dev),
            rotations=g["rotations"].to(dev), tile_mask=ones)
    assert int(out[8].max()) == 0 and float(out[6].min()) == 1.0
    # tile mask all zeros: radii are still reported (N7) but nothing is rendered
    out = r(means3D=xyz, opacities=g["opacity"].to(dev), colors_precomp=g["rgb"].to(dev), scales=g["scales"].to(dev),
            rotations=g["rotations"].to(dev), tile_mask=torch.zeros_like(ones))
    assert int(out[8].max()) > 0 and float(out[6].min()) == 1.0 and float(out[0].abs().max()) == 0


def test_capacity_overflow_is_detected_and_recovered():
    dev = torch.device(DEV)
    L = _lib.lib()
    inp = rh.make_inputs("small", dev, P=16000, sh_degree=0)
    inp["scales"] = inp["scales"] * 4.0
    ref = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
    R = ref[0]
    assert R > 70000, R  # larger than the wrapper's minimum capacity so that the first attempt overflows
    rasterizer._capacity_hint.clear()
    again = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
    assert again[0] == R and torch.equal(again[2], ref[2]) and torch.equal(again[5], ref[5])
    # raw C-ABI call with a too-small capacity: overflow flag set, nothing rendered, no crash
    st = ref[10]._dqo_state
    cap = 1024
    binning = torch.empty((L.dqo_rast_binning_bytes(cap),), dtype=torch.uint8, device=dev)
    status = torch.zeros(8, dtype=torch.int32, device=dev)
    args = rh.raster_args(inp)
    p = _lib.ptr
    cam = inp["cam"]
    H, W, P = cam.image_height, cam.image_width, inp["xyz"].shape[0]
    f = lambda *s: torch.empty(s, device=dev)
    i = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
    color, depth, hd, hc, hcw, hdw, T = f(3, H, W), f(H, W), i(H, W), i(H, W), f(H, W), f(H, W), f(H, W)
    radii, nt, tidx = i(P), i(P), i(((H + 15) // 16) * ((W + 15) // 16))
    code = L.dqo_rast_forward(st.settings, p(inp["bg"]), p(inp["xyz"]), p(inp["shs"]), None, p(inp["opacity"]),
                              p(inp["scales"]), p(inp["rotations"]), None, p(cam.world_view_transform),
                              p(cam.full_proj_transform), p(cam.camera_center), p(inp["tile_mask"]), p(st.geom),
                              p(binning), cap, p(st.image), p(tidx), p(color), p(depth), p(hd), p(hc), p(hcw), p(hdw),
                              p(T), p(radii), p(nt), p(status), torch.cuda.current_stream().cuda_stream)
    assert code == 0
    s = status.tolist()
    assert s[_lib.ST_OVERFLOW] == 1 and s[_lib.ST_NUM_RENDERED] == R and s[_lib.ST_TILE_NUM] == 0
    assert float(T.min()) == 1.0


def test_autograd_api_matches_pybind_path():
    dev = torch.device(DEV)
    inp = rh.make_inputs("tiny", dev, P=2500, sh_degree=3)
    cam = inp["cam"]
    rd = synthetic.RENDER_DEFAULTS
    s = rasterizer.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=inp["bg"], scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=3, campos=cam.camera_center, opaque_threshold=rd["opaque_threshold"],
        normal_threshold=rd["normal_threshold"], depth_threshold=rd["depth_threshold"], prefiltered=False, debug=False,
        cx=cam.cx, cy=cam.cy)
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ["xyz", "shs", "opacity", "scales", "rotations"]}
    out = rasterizer.GaussianRasterizer(s)(means3D=leaves["xyz"], opacities=leaves["opacity"], shs=leaves["shs"],
                                           scales=leaves["scales"], rotations=leaves["rotations"],
                                           tile_mask=inp["tile_mask"], normal_w=None)
    assert len(out) == 9
    gc, gd = rh.make_pixel_grads(cam.image_height, cam.image_width, DEV)
    (out[0] * gc).sum().add((out[1] * gd).sum()).backward()
    o = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
    bw = rasterizer.rasterize_gaussians_backward(*rh.backward_args(inp, o, gc, gd))
    assert torch.equal(out[0], o[2]) and torch.equal(out[3], o[5]) and torch.equal(out[8], o[9])
    for name, g in [("xyz", bw[3]), ("shs", bw[5]), ("opacity", bw[2]), ("scales", bw[6]), ("rotations", bw[7])]:
        a, b = leaves[name].grad, g.reshape(leaves[name].shape)
        assert float((a - b).norm() / (b.norm() + 1e-30)) <= 1e-4, name
    vis = rasterizer.GaussianRasterizer(s).markVisible(inp["xyz"])
    assert vis.dtype == torch.bool and bool(vis[o[9] > 0].all())
    assert np.array_equal(t2n(vis), oracle.mark_visible(t2n(inp["xyz"]), t2n(cam.world_view_transform),
                                                        t2n(cam.full_proj_transform)))


@pytest.mark.parametrize("cfg,mask,binning", [("small", "half", ("single",)), ("c1", "ones", ("single",)),
                                              ("c2", "ones", ("fixed", 2_000_128, 1_000_000))])
def test_extra_colour_blend_equals_second_full_render(cfg, mask, binning):
    """the semantic / instance images of Renderer.render (render.py:227-262) from the lists of the main
    render: bit-identical to a full second rasterizer call with colors_precomp (ours, and the reference when present)."""
    from dqo_map_b200 import render as render_mod
    inp = rh.make_inputs(cfg, torch.device(DEV), mask=mask)
    cam = inp["cam"]
    P = inp["xyz"].shape[0]
    g = torch.Generator().manual_seed(3)
    sem, inst = torch.rand(P, 3, generator=g).to(DEV), (torch.randint(0, 20, (P, 1), generator=g) / 255.0).repeat(1, 3).to(DEV)
    normal = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1).to(DEV)
    rd = synthetic.RENDER_DEFAULTS
    rs = rasterizer.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.tensor([0.1, 0.2, 0.3], device=DEV), scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=inp["sh_degree"], campos=cam.camera_center,
        opaque_threshold=rd["opaque_threshold"], normal_threshold=rd["normal_threshold"],
        depth_threshold=rd["depth_threshold"], prefiltered=False, debug=False, cx=cam.cx, cy=cam.cy)
    data = dict(xyz=inp["xyz"], opacity=inp["opacity"], scales=inp["scales"], rotations=inp["rotations"], shs=inp["shs"],
                normal=normal, semantics_color=sem, instance=inst)
    rasterizer.set_binning_mode(*binning)
    res = render_mod.render(rs, data, inp["tile_mask"])
    rasterizer.set_binning_mode("single")
    for key, colors in (("semantic_seg", sem), ("instance", inst)):
        full = rasterizer.GaussianRasterizer(rs)(means3D=inp["xyz"], opacities=inp["opacity"], colors_precomp=colors,
                                                 scales=inp["scales"], rotations=inp["rotations"], tile_mask=inp["tile_mask"])
        assert torch.equal(res[key], full[0]), key
        if rh.reference_available():
            ref_pkg = rh.load_reference()[0]
            rs_ref = ref_pkg.GaussianRasterizationSettings(**rs._asdict())
            ref_img = ref_pkg.GaussianRasterizer(rs_ref)(means3D=inp["xyz"], opacities=inp["opacity"], colors_precomp=colors,
                                                         scales=inp["scales"], rotations=inp["rotations"],
                                                         tile_mask=inp["tile_mask"])[0]
            assert torch.equal(res[key], ref_img), key
    # the normal map is the reference's gather of per-Gaussian normals through the depth index map (render.py:212-216)
    di = res["depth_index_map"][0]
    want = torch.zeros(3, cam.image_height, cam.image_width, device=DEV)
    want[:, di > -1] = normal[di[di > -1].long()].T
    assert torch.equal(res["normal"], want)


def test_render_semantic_image_is_differentiable_like_the_reference():
    """Renderer.render's semantic / instance images are full differentiable rasterizer calls in the reference
    (SLAM/render.py:227-262), and loss_update's semantic L1 (mapper.py:876-879) trains `_semantics` and the geometry through
    them.  render() must hand back the same graph: gradients reach the semantic colours and the positions and equal those
    of an explicit second GaussianRasterizer call."""
    from dqo_map_b200 import render as render_mod
    inp = rh.make_inputs("small", torch.device(DEV), mask="ones")
    cam = inp["cam"]
    P = inp["xyz"].shape[0]
    g = torch.Generator().manual_seed(4)
    rd = synthetic.RENDER_DEFAULTS
    rs = rasterizer.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.zeros(3, device=DEV), scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=inp["sh_degree"], campos=cam.camera_center,
        opaque_threshold=rd["opaque_threshold"], normal_threshold=rd["normal_threshold"],
        depth_threshold=rd["depth_threshold"], prefiltered=False, debug=False, cx=cam.cx, cy=cam.cy)
    gt_sem = torch.rand(3, cam.image_height, cam.image_width, generator=g).to(DEV)
    sem0 = torch.rand(P, 3, generator=g).to(DEV)
    normal = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1).to(DEV)

    def leaves():
        return sem0.clone().requires_grad_(True), inp["xyz"].clone().requires_grad_(True)

    sem_a, xyz_a = leaves()
    data = dict(xyz=xyz_a, opacity=inp["opacity"], scales=inp["scales"], rotations=inp["rotations"], shs=inp["shs"],
                normal=normal, semantics_color=sem_a, instance=None)
    res = render_mod.render(rs, data, inp["tile_mask"])
    assert res["semantic_seg"].requires_grad and res["instance"] is None
    (0.1 * (res["semantic_seg"] - gt_sem).abs().mean()).backward()
    sem_b, xyz_b = leaves()
    full = rasterizer.GaussianRasterizer(rs)(means3D=xyz_b, opacities=inp["opacity"], colors_precomp=sem_b,
                                             scales=inp["scales"], rotations=inp["rotations"], tile_mask=inp["tile_mask"])[0]
    assert torch.equal(res["semantic_seg"].detach(), full.detach())
    (0.1 * (full - gt_sem).abs().mean()).backward()
    assert float(sem_a.grad.abs().sum()) > 0 and float(xyz_a.grad.abs().sum()) > 0
    assert float((sem_a.grad - sem_b.grad).norm() / sem_b.grad.norm()) <= 1e-4
    assert float((xyz_a.grad - xyz_b.grad).norm() / xyz_b.grad.norm()) <= 1e-3
    # evaluation renders (no grad) take the shared-binning blend and return the same image
    with torch.no_grad():
        res_ng = render_mod.render(rs, dict(data, xyz=inp["xyz"], semantics_color=sem0), inp["tile_mask"])
    assert torch.equal(res_ng["semantic_seg"], full.detach())


@pytest.mark.parametrize("two_phase", [False, True])
def test_differentiable_extra_blend_gradients_for_every_input(two_phase):
    """f3 with gradients (`rasterizer.blend_extra_colors_grad`, `dqo_rast_blend_extra_backward`): a loss on the main image,
    the depth AND the semantic image; the gradients of every leaf (semantic colours, positions, opacities, scales,
    rotations) equal those of the reference's graph -- main call plus a second full rasterizer call with colors_precomp
    (SLAM/render.py:187-246) -- for single-phase and two-phase binning."""
    from dqo_map_b200 import render as render_mod
    inp = rh.make_inputs("small", torch.device(DEV), mask="half")
    cam = inp["cam"]
    P, H, W = inp["xyz"].shape[0], cam.image_height, cam.image_width
    g = torch.Generator().manual_seed(8)
    rd = synthetic.RENDER_DEFAULTS
    rs = rasterizer.GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.tensor([0.05, 0.1, 0.0], device=DEV), scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=inp["sh_degree"], campos=cam.camera_center,
        opaque_threshold=rd["opaque_threshold"], normal_threshold=rd["normal_threshold"],
        depth_threshold=rd["depth_threshold"], prefiltered=False, debug=False, cx=cam.cx, cy=cam.cy)
    w_img, w_sem = torch.randn(3, H, W, generator=g).to(DEV), torch.randn(3, H, W, generator=g).to(DEV)
    w_dep = (0.1 * torch.randn(1, H, W, generator=g)).to(DEV)
    sem0 = torch.rand(P, 3, generator=g).to(DEV)
    names = ("sem", "xyz", "opacity", "scales", "rotations")

    def leaves():
        return {"sem": sem0.clone().requires_grad_(True), "xyz": inp["xyz"].clone().requires_grad_(True),
                "opacity": inp["opacity"].clone().requires_grad_(True), "scales": inp["scales"].clone().requires_grad_(True),
                "rotations": inp["rotations"].clone().requires_grad_(True)}

    binning = ("single",)
    if two_phase:
        rasterizer.set_binning_mode("single")
        R = rasterizer.rasterize_gaussians(*rh.raster_args(inp))[0]
        binning = ("fixed", max(256, (R // 6) // 256 * 256), R + 1024)
    rasterizer.set_binning_mode(*binning)
    try:
        a = leaves()
        res = render_mod.render(rs, dict(xyz=a["xyz"], opacity=a["opacity"], scales=a["scales"], rotations=a["rotations"],
                                         shs=inp["shs"], normal=None, semantics_color=a["sem"], instance=None),
                                inp["tile_mask"])
        ((res["render"] * w_img).sum() + (res["depth"] * w_dep).sum() + (res["semantic_seg"] * w_sem).sum()).backward()
        b = leaves()
        rast = rasterizer.GaussianRasterizer(rs)
        kw = dict(means3D=b["xyz"], opacities=b["opacity"], scales=b["scales"], rotations=b["rotations"],
                  tile_mask=inp["tile_mask"])
        main = rast(shs=inp["shs"], **kw)
        sem_img = rast(colors_precomp=b["sem"], **kw)[0]
        ((main[0] * w_img).sum() + (main[1] * w_dep).sum() + (sem_img * w_sem).sum()).backward()
    finally:
        rasterizer.set_binning_mode("single")
    assert torch.equal(res["semantic_seg"].detach(), sem_img.detach()) and torch.equal(res["render"].detach(), main[0].detach())
    for n in names:
        ga, gb = a[n].grad, b[n].grad
        assert ga is not None and float(gb.abs().sum()) > 0, n
        rel = float((ga - gb).norm() / gb.norm())
        assert rel <= 1e-5, (n, rel)      # the same kernels and the same fp64 accumulation on both sides


def _arbiter_inputs(inp, o, ex):
    cam = inp["cam"]
    rd = synthetic.RENDER_DEFAULTS
    scene = dict(xyz=inp["xyz"], scales=inp["scales"], rotations=inp["rotations"], opacity=inp["opacity"],
                 shs=None if inp["precomp"] else inp["shs"], rgb=inp["rgb"] if inp["precomp"] else None,
                 view=cam.world_view_transform, proj=cam.full_proj_transform, campos=cam.camera_center, bg=inp["bg"],
                 W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, cx=cam.cx, cy=cam.cy,
                 sh_degree=0 if inp["precomp"] else inp["sh_degree"], scale_modifier=rd["scale_modifier"],
                 opaque_threshold=rd["opaque_threshold"], depth_threshold=rd["depth_threshold"],
                 normal_threshold=rd["normal_threshold"])
    lists = dict(point_list=torch.from_numpy(ex["point_list"].astype(np.int64)),
                 ranges=torch.from_numpy(ex["ranges"].astype(np.int64)),
                 n_contrib=torch.from_numpy(ex["n_contrib"].astype(np.int64)), hit=o[5][0].long(), radii=o[9])
    return scene, lists


@pytest.mark.parametrize("cfg,mask,precomp", [("tiny", "ones", False), ("small", "half", False), ("deg1", "ones", False),
                                              ("tiny", "ones", True)])
def test_gradients_against_float64_arbiter(cfg, mask, precomp):
    """VERDICT r1 / SURVEY 7: the 1e-3 gradient gate with a float64 evaluation of the same function as arbiter
    (oracle/f64_arbiter.py: torch float64 + autograd, independent of every hand-written backward):
        |ours - f64| <= max(1e-3 |f64|, 1.5 |reference - f64|)   per gradient tensor, norm-wise.
    No multiple-of-self-noise floor and no re-draw: both float32 implementations are measured against the exact value.
    Both errors are medians of four runs: the float atomics of either implementation land in unspecified order and the
    norm is dominated by ONE ill-conditioned Gaussian per scene (tests/dev_arbiter_diag.py: > 99 % of the squared error
    of ours AND of the reference sits in the same Gaussian; the reference's own four runs range 4.8e-4 .. 9.1e-4, ours
    5e-4 .. 1.1e-3)."""
    from oracle import f64_arbiter
    inp = rh.make_inputs(cfg, torch.device(DEV), mask=mask, precomp=precomp)
    cam = inp["cam"]
    gc, gd = rh.make_pixel_grads(cam.image_height, cam.image_width, DEV)
    o, ex, bw = _run_ours(inp, gc, gd)
    scene, lists = _arbiter_inputs(inp, o, ex)
    exact = f64_arbiter.gradients(scene, lists, gc, gd, device=DEV)
    # both float32 implementations accumulate with float atomics in unspecified order: four runs each, medians compared
    ours = [dict(zip(GRADS, bw))] + [dict(zip(GRADS, rasterizer.rasterize_gaussians_backward(*rh.backward_args(inp, o, gc, gd))))
                                     for _ in range(3)]
    refs = []
    if rh.reference_available():
        C = rh.load_reference()[1]
        fwd = C.rasterize_gaussians(*rh.raster_args(inp))
        refs = [dict(zip(GRADS, C.rasterize_gaussians_backward(*rh.backward_args(inp, fwd, gc, gd)))) for _ in range(4)]
    checked = 0
    for name, e in exact.items():
        if e.numel() <= 1:
            continue
        e = e.double()
        nrm = float(e.norm())
        assert nrm > 0, name
        err = float(np.median([float((a[name].double().reshape(e.shape) - e).norm()) for a in ours]))
        gate = 1e-3 * nrm
        if refs:
            gate = max(gate, 1.5 * float(np.median([float((r[name].double().reshape(e.shape) - e).norm()) for r in refs])))
        assert err <= gate, (name, err / nrm, gate / nrm)
        checked += 1
    assert checked == 5


@pytest.mark.parametrize("cfg,binning", [("small", ("single",)), ("c1", ("single",)), ("c2", ("fixed", 2_000_128, 4_000_000))])
def test_backward_is_reproducible_run_to_run(cfg, binning):
    """The backward blend adds each warp's fp32 partial sums into fp64 accumulators (common.cuh: DQO_GACC_FLOATS), so the
    order in which warps and tiles arrive does not show in the result: repeated runs return bit-identical gradients (the
    reference's float atomics make its own runs differ by up to 1e-3 in dL_drotations, profiles/r01_gradient_parity_*)."""
    inp = rh.make_inputs(cfg, torch.device(DEV))
    cam = inp["cam"]
    gc, gd = rh.make_pixel_grads(cam.image_height, cam.image_width, DEV)
    o, _, bw = _run_ours(inp, gc, gd, binning=binning)
    first = [t.clone() for t in bw]
    for _ in range(3):
        again = rasterizer.rasterize_gaussians_backward(*rh.backward_args(inp, o, gc, gd))
        for name, a, b in zip(GRADS, first, again):
            assert torch.equal(a, b), name
    rasterizer.set_binning_mode("single")


@pytest.mark.parametrize("cfg,binning", [("c1", None), ("c2", (2_000_128, 4_000_000))])
def test_raster_pipeline_matches_operator_path_over_a_window(cfg, binning):
    """`RasterPipeline` (pre-allocated workspaces, geom_clean = 2: self-cleaning accumulators, gradient rows that stay zero are
    not rewritten) against the allocate-per-call operator path, cycling over three keyframes so that rows turn non-zero,
    zero and non-zero again: every output image and every gradient tensor identical, bit for bit, at every iteration."""
    import bench
    dev = torch.device(DEV)
    inp, views = bench.make_views(cfg, dev, 0, 3)
    cam0 = views[0]["cam"]
    P, H, W, M = inp["xyz"].shape[0], cam0.image_height, cam0.image_width, inp["shs"].shape[1]
    gc, gd = rh.make_pixel_grads(H, W, DEV)
    front, back = binning if binning else (0, 0)
    pipe = rasterizer.RasterPipeline(P, M, W, H, (front + back) if front else 16 * P, dev, front, back)
    rasterizer.set_binning_mode(*(("fixed", front, back) if front else ("single",)))
    try:
        for k in range(5):
            v = views[k % 3]
            rs = v["settings"](rasterizer.GaussianRasterizationSettings)
            pipe.forward(rs, inp["xyz"], inp["opacity"], inp["scales"], inp["rotations"], inp["tile_mask"], shs=inp["shs"])
            pipe.backward(gc, gd)
            pipe.check()
            vin = dict(inp)
            vin["cam"] = v["cam"]
            o = rasterizer.rasterize_gaussians(*rh.raster_args(vin))
            bw = rasterizer.rasterize_gaussians_backward(*rh.backward_args(vin, o, gc, gd))
            assert torch.equal(pipe.color, o[2]) and torch.equal(pipe.depth, o[3]) and torch.equal(pipe.hit_depth, o[5])
            mine = dict(dL_dmeans2D=pipe.g_means2D, dL_dcolors=pipe.g_colors, dL_dopacity=pipe.g_opacity,
                        dL_dmeans3D=pipe.g_means3D, dL_dcov3D=pipe.g_cov3D, dL_dsh=pipe.g_sh, dL_dscales=pipe.g_scales,
                        dL_drotations=pipe.g_rot)
            for name, t in zip(GRADS, bw):
                if name in mine and t.numel():
                    m = mine[name].reshape(t.shape)
                    if not torch.equal(m, t):
                        d = (m - t).abs()
                        rows = (d.reshape(d.shape[0], -1) > 0).any(dim=1)
                        raise AssertionError((k, name, "rows differing", int(rows.sum()), "max abs", float(d.max()),
                                              "max |ref|", float(t.abs().max()), "worst rel",
                                              float((d / (t.abs() + 1e-30))[d > 0].max()),
                                              "rows zero in mine", int(((m.reshape(m.shape[0], -1) == 0).all(1) & rows).sum())))
    finally:
        rasterizer.set_binning_mode("single")
