"""CPU: the oracle (oracle/dqo_oracle.c) against the fixtures produced by the reference itself on a B200
(tests/golden/make_golden.py).  Integer / index artefacts and every float that does not pass through expf are
bit-exact; floats downstream of expf agree to ~1 ulp (MUFU.EX2 cannot be reproduced on a CPU)."""
import os

import numpy as np
import pytest

from oracle import oracle

CASES = ["tiny_sh0_full", "tiny_sh3_half", "tiny_precomp"]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module", params=CASES)
def case(request, golden_dir):
    g = np.load(os.path.join(golden_dir, request.param + ".npz"))
    sc = oracle.Scene.from_golden(g)
    pre, bn, img = oracle.forward(sc)
    return g, sc, pre, bn, img


def test_preprocess_bit_exact(case):
    g, sc, pre, bn, img = case
    assert np.array_equal(pre["radii"], g["radii"])
    assert np.array_equal(pre["tiles_touched"], g["tiles_touched"])
    vis = g["radii"] > 0
    for k in ["means2D", "depths", "conic_opacity", "cov3D"]:
        assert np.array_equal(_bits(pre[k][vis]), _bits(g[k][vis])), k
    if not int(g["precomp"]):
        # every SH degree mirrors the compiled reference's operation order (decoded from its SASS): bit-exact colours
        assert np.array_equal(_bits(pre["rgb"][vis]), _bits(g["rgb"][vis]))
        assert np.array_equal(pre["clamped"][vis], g["clamped"][vis])


def test_binning_bit_exact(case):
    g, sc, pre, bn, img = case
    assert bn["num_rendered"] == int(g["num_rendered"])
    assert bn["tile_num"] == int(g["tile_num"])
    assert np.array_equal(bn["keys_sorted"], g["keys_sorted"])
    assert np.array_equal(bn["point_list"], g["point_list"])
    assert np.array_equal(bn["ranges"], g["ranges"])
    assert np.array_equal(bn["tile_indices"][: bn["tile_num"]], g["tile_indices"][: bn["tile_num"]])
    # keys are sorted by (tile, depth) and stable in the Gaussian index
    k = bn["keys_sorted"]
    assert np.all(k[1:] >= k[:-1])
    same = k[1:] == k[:-1]
    assert np.all(bn["point_list"][1:][same] > bn["point_list"][:-1][same])


def test_render_forward(case):
    g, sc, pre, bn, img = case
    rendered = g["T_map"][0] != 1.0
    # index maps / counters: exact up to alpha-threshold flips caused by the 1-ulp expf difference (none observed)
    assert np.array_equal(np.where(rendered, img["n_contrib"], 0), np.where(rendered, g["n_contrib"], 0))
    for k in ["hit_depth", "hit_color", "n_touched"]:
        assert np.array_equal(img[k], g[k]), k
    assert np.array_equal(_bits(img["depth"]), _bits(g["depth"]))  # depth does not depend on expf
    for k in ["color", "T_map", "hit_color_weight", "hit_depth_weight"]:
        assert np.abs(img[k] - g[k]).max() <= 1e-6, k  # north_star gate is 1e-4


def test_backward(case):
    g, sc, pre, bn, img = case
    gr = oracle.backward(sc, pre, bn, img, g["grad_color"], g["grad_depth"])
    for k in ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations"]:
        b = g[k].astype(np.float64)
        if b.size == 0:
            continue
        a = gr[k].astype(np.float64).reshape(b.shape)
        rel = np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
        assert rel <= 1e-4, (k, rel)  # north_star gate is 1e-3 relative


def test_higher_msb_is_bit_length():
    for n in list(range(1, 5000)) + [8160, 65535, 65536, 1 << 20]:
        assert oracle.higher_msb(n) == n.bit_length()


def test_mark_visible_matches_radii_support(case):
    g, sc, pre, bn, img = case
    vis = oracle.mark_visible(g["xyz"], g["viewmatrix"], g["projmatrix"])
    assert np.all(vis[g["radii"] > 0])  # everything rendered passed the frustum test


def test_knn_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "knn_4000.npz"))
    md, ki = oracle.knn(g["points"])
    assert np.array_equal(ki, g["knn_idx"])
    assert np.array_equal(_bits(md), _bits(g["mean_dist2"]))


def test_accumulate_error_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "accum_error.npz"))
    H, W = g["ce"].shape[:2]
    P = int(g["P"])
    for cm, tag in ((True, "max"), (False, "mean")):
        outs = oracle.accumulate_error(H, W, P, g["ce"], g["de"], g["ne"], g["ci"], g["di"], 0.5, 0.6, 0.7, cm)
        for k, o in zip(["color", "depth", "normal", "rescale"], outs):
            ref = g["%s_%s" % (k, tag)]
            if cm or k == "rescale":
                assert np.array_equal(o, ref), (k, tag)
            else:
                np.testing.assert_allclose(o, ref, rtol=1e-6, atol=1e-7)


def test_oracle_edge_cases():
    # all Gaussians behind the camera -> nothing rendered, fill values kept
    rng = np.random.RandomState(0)
    P, W, H = 50, 64, 48
    xyz = rng.rand(P, 3).astype(np.float32)
    xyz[:, 2] = -1.0
    view = np.eye(4, dtype=np.float32)
    proj = np.eye(4, dtype=np.float32)
    proj[2, 3] = 1.0
    sc = oracle.Scene(xyz, np.full((P, 3), 0.01, np.float32), np.tile([1, 0, 0, 0], (P, 1)), np.full(P, 0.9), view, proj,
                      np.zeros(3), W, H, 0.5, 0.4, W / 2, H / 2, np.zeros(3), np.ones((3, 4), np.int32),
                      colors_precomp=rng.rand(P, 3))
    pre, bn, img = oracle.forward(sc)
    assert bn["num_rendered"] == 0 and bn["tile_num"] == 0
    assert np.all(pre["radii"] == 0)
    assert np.all(img["T_map"] == 1.0) and np.all(img["hit_depth"] == 0) and np.all(img["color"] == 0)
