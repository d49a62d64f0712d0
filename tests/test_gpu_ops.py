"""GPU parity tests of the operators around the rasterizer: distCUDA2, accumulate_gaussian_error, masked L1 loss,
fused Adam, quadric kernels — all through the C-ABI, against golden fixtures and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import refharness as rh
from dqo_map_b200 import knn, map_utils, mapping, quadric
from oracle import oracle
from oracle import quadric_oracle as qo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t2n(t):
    return t.detach().cpu().numpy()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_knn_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "knn_4000.npz"))
    md, ki = knn.distCUDA2(torch.tensor(g["points"]).to(DEV))
    assert np.array_equal(t2n(ki), g["knn_idx"])
    assert np.array_equal(_bits(t2n(md)), _bits(g["mean_dist2"]))


@pytest.mark.parametrize("P", [1, 2, 3, 4, 33, 1024, 1025, 50000])
def test_knn_against_oracle(P):
    gen = torch.Generator(device="cpu").manual_seed(P)
    pts = torch.randn(P, 3, generator=gen) * 2.0
    if P >= 33:  # clustered + duplicated points: distance ties and empty Morton cells
        pts[: P // 3] = pts[: P // 3] * 0.01 + 1.0
        pts[5:9] = pts[5]
    md, ki = knn.distCUDA2(pts.to(DEV))
    omd, oki = oracle.knn(pts.numpy())
    assert np.array_equal(t2n(ki), oki)
    assert np.array_equal(_bits(t2n(md)), _bits(omd))
    if P >= 4:
        # property: reported neighbours are the true 3 nearest (brute force on a sample)
        idx = np.random.RandomState(0).choice(P, size=min(P, 200), replace=False)
        x = pts.numpy().astype(np.float64)
        for i in idx:
            d = ((x - x[i]) ** 2).sum(1)
            d[i] = np.inf
            best = np.sort(d)[:3]
            got = ((x[t2n(ki)[i]] - x[i]) ** 2).sum(1)
            np.testing.assert_allclose(np.sort(got), best, rtol=1e-5, atol=1e-12)


@pytest.mark.skipif(not rh.reference_available(), reason="oracle/_ref not built")
def test_knn_against_live_reference_500k():
    """BASELINE config 4: distCUDA2 for 500k new points (+50k existing neighbours)."""
    _, _, knn_C, _ = rh.load_reference()
    gen = torch.Generator(device="cpu").manual_seed(4)
    pts = torch.cat([torch.rand(500000, 3, generator=gen) * torch.tensor([6.0, 3.0, 6.0]),
                     torch.rand(50000, 3, generator=gen) * 0.5 + 2.0]).to(DEV)
    md_r, ki_r = knn_C.distCUDA2(pts)
    md, ki = knn.distCUDA2(pts)
    torch.cuda.synchronize()
    assert torch.equal(ki, ki_r)
    assert torch.equal(md.view(torch.int32), md_r.view(torch.int32))


def test_accumulate_error_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "accum_error.npz"))
    H, W = g["ce"].shape[:2]
    P = int(g["P"])
    tens = [torch.tensor(g[k]).to(DEV) for k in ["ce", "de", "ne", "ci", "di"]]
    for cm, tag in ((True, "max"), (False, "mean")):
        outs = map_utils.accumulate_gaussian_error(H, W, P, *tens, 0.5, 0.6, 0.7, cm)
        assert all(o.shape == (P, 1) for o in outs)
        for k, o in zip(["color", "depth", "normal", "rescale"], outs):
            ref = g["%s_%s" % (k, tag)]
            if cm or k == "rescale":
                assert np.array_equal(t2n(o), ref), (k, tag)
            else:
                np.testing.assert_allclose(t2n(o), ref, rtol=1e-5, atol=1e-7)
    # empty image / P == 0
    e = torch.zeros(0, 0, 1, device=DEV)
    outs = map_utils.accumulate_gaussian_error(0, 0, 7, e, e, e, e.int(), e.int(), 0.1, 0.1, 0.1, True)
    assert all(float(o.abs().sum()) == 0 for o in outs)


@pytest.mark.parametrize("use_mask,depth_weight", [(True, 1.0), (False, 1.0), (True, 0.0)])
def test_masked_l1_loss(use_mask, depth_weight):
    gen = torch.Generator(device="cpu").manual_seed(9)
    H, W = 67, 131
    image = torch.rand(3, H, W, generator=gen)
    depth = torch.rand(1, H, W, generator=gen) * 3
    hit = torch.randint(-1, 50, (1, H, W), generator=gen, dtype=torch.int32)
    gt_c = torch.rand(H, W, 3, generator=gen)
    gt_c[:5] = image.permute(1, 2, 0)[:5]  # exact zeros of the error: sign(0) = 0
    gt_d = torch.rand(H, W, 1, generator=gen) * 3
    gt_d[gt_d < 0.3] = 0.0
    mask = (torch.rand(H, W, generator=gen) < 0.6) if use_mask else None
    ref = oracle.masked_l1_loss(image.numpy(), depth.numpy(), hit.numpy(), gt_c.numpy(), gt_d.numpy(),
                                None if mask is None else mask.numpy(), 0.8, depth_weight, 0.1)
    img_d = image.to(DEV).requires_grad_(True)
    dep_d = depth.to(DEV).requires_grad_(True)
    total, lc, ld, counts = mapping.masked_l1_loss(img_d, dep_d, hit.to(DEV), gt_c.to(DEV), gt_d.to(DEV),
                                                   None if mask is None else mask.to(DEV), 0.8, depth_weight, 0.1)
    total.backward()
    assert abs(float(total) - ref[0]) <= 1e-6 * max(1.0, abs(ref[0]))
    assert abs(float(lc) - ref[1]) <= 1e-6 and abs(float(ld) - ref[2]) <= 1e-6
    np.testing.assert_allclose(t2n(img_d.grad), ref[3], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(t2n(dep_d.grad), ref[4], rtol=1e-5, atol=1e-10)
    # empty selection -> NaN loss (torch.mean of an empty tensor), zero gradients
    total2, *_ = mapping.masked_l1_loss(image.to(DEV), depth.to(DEV), hit.to(DEV), gt_c.to(DEV), gt_d.to(DEV),
                                        torch.zeros(H, W, dtype=torch.bool, device=DEV), 0.8, 1.0, 0.1)
    assert bool(torch.isnan(total2))


def test_fused_adam_matches_torch_adam():
    gen = torch.Generator(device="cpu").manual_seed(21)
    shapes = dict(xyz=(5003, 3), f_dc=(5003, 1, 3), f_rest=(5003, 15, 3), opacity=(5003, 1), scaling=(5003, 3), rotation=(5003, 4))
    lrs = dict(xyz=1e-3, f_dc=5e-4, f_rest=2.5e-5, opacity=0.0, scaling=4e-3, rotation=1e-3)
    init = {k: torch.randn(s, generator=gen) for k, s in shapes.items()}
    steps = 7
    grads = []
    for _ in range(steps):
        g = {k: torch.randn(s, generator=gen) * 10 ** float(torch.randint(-6, 1, (1,), generator=gen)) for k, s in shapes.items()}
        g["f_dc"][::3] = 0.0
        grads.append(g)
    names = list(shapes)
    ref = oracle.adam_reference([init[k].numpy() for k in names], [[g[k].numpy() for k in names] for g in grads],
                                [lrs[k] for k in names], eps=1e-15)
    ps = {k: torch.nn.Parameter(init[k].clone().to(DEV)) for k in names}
    conf = torch.zeros(shapes["xyz"][0], 1, device=DEV)
    opt = mapping.FusedAdam([{"params": [ps[k]], "lr": lrs[k], "name": k} for k in names], lr=0.0, eps=1e-15,
                            confidence=conf, confidence_param=ps["f_dc"])
    for g in grads:
        for k in names:
            ps[k].grad = g[k].clone().to(DEV)
        opt.step()
    for k, (p_ref, m_ref, v_ref) in zip(names, ref):
        # CPU torch.optim.Adam (oracle): parameters tight; the moments differ by CPU-vs-GPU lerp rounding on
        # elements that cancelled to ~1e-6 of the tensor scale
        np.testing.assert_allclose(t2n(ps[k]), p_ref, rtol=2e-6, atol=1e-7, err_msg=k)
        np.testing.assert_allclose(t2n(opt.state[ps[k]]["exp_avg"]), m_ref, rtol=1e-5, atol=1e-6 * np.abs(m_ref).max(), err_msg=k)
        np.testing.assert_allclose(t2n(opt.state[ps[k]]["exp_avg_sq"]), v_ref, rtol=1e-5, atol=1e-6 * np.abs(v_ref).max(), err_msg=k)
    # same-device arbiter: stock torch.optim.Adam on the GPU
    qs = {k: torch.nn.Parameter(init[k].clone().to(DEV)) for k in names}
    topt = torch.optim.Adam([{"params": [qs[k]], "lr": lrs[k], "name": k} for k in names], lr=0.0, eps=1e-15)
    for g in grads:
        for k in names:
            qs[k].grad = g[k].clone().to(DEV)
        topt.step()
    for k in names:
        np.testing.assert_allclose(t2n(ps[k]), t2n(qs[k]), rtol=2e-6, atol=1e-7, err_msg=k)
        np.testing.assert_allclose(t2n(opt.state[ps[k]]["exp_avg"]), t2n(topt.state[qs[k]]["exp_avg"]), rtol=1e-5,
                                   atol=1e-6 * float(topt.state[qs[k]]["exp_avg"].abs().max()), err_msg=k)
    expect = sum(((g["f_dc"].abs() != 0).any(dim=-1)).float() for g in grads)  # mapper.py:909-910
    assert torch.equal(conf.cpu(), expect)
    assert torch.equal(ps["opacity"].detach().cpu(), init["opacity"])  # lr 0.0 (configs/replica_base.yaml:20)


def test_quadric_init_and_project(golden_dir):
    g = np.load(os.path.join(golden_dir, "quadric.npz"))
    ax, R, c = quadric.quadric_init(g["init_bbox"], g["init_depth_stats"], g["K"], g["init_Rt"], device=DEV)
    np.testing.assert_allclose(t2n(ax), g["init_axes"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(t2n(R), g["init_R"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(t2n(c), g["init_center"], rtol=1e-12, atol=1e-12)
    bb, ell = quadric.quadric_project(ax, R, c, torch.tensor(g["proj_P"]).to(DEV))
    np.testing.assert_allclose(t2n(bb), g["proj_bbox"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(t2n(ell), g["proj_ellipse"], rtol=1e-7, atol=1e-7)


def test_quadric_refine(golden_dir):
    g = np.load(os.path.join(golden_dir, "quadric.npz"))
    n, V = g["obs_bboxes"].shape[:2]
    ax, R, c, last = quadric.quadric_refine(g["init_axes"], g["init_R"], g["init_center"],
                                            torch.tensor(g["obs_bboxes"]).to(DEV), torch.tensor(g["Ps"]).to(DEV),
                                            np.full(n, V, np.int32), g["view_choice"])
    # the reference's own Object_Optimize_only (fp32 torch autograd + torch.linalg.eig), tests/golden/make_quadric_golden.py
    np.testing.assert_allclose(t2n(ax), g["refined_axes"], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(t2n(R).reshape(n, 3, 3), g["refined_R"], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(t2n(c), g["refined_center"], rtol=2e-3, atol=2e-4)
    # one iteration against the torch-autograd oracle: the analytic gradient itself (Adam's first step is lr*sign(g))
    for i in range(n):
        oax, oR, oc, _ = qo.refine(g["init_axes"][i], g["init_R"][i], g["init_center"][i], g["obs_bboxes"][i], g["Ps"][i],
                                   g["view_choice"][i], iters=3)
        ax3, R3, c3, _ = quadric.quadric_refine(g["init_axes"][i:i + 1], g["init_R"][i:i + 1], g["init_center"][i:i + 1],
                                                torch.tensor(g["obs_bboxes"][i:i + 1]).to(DEV),
                                                torch.tensor(g["Ps"][i:i + 1]).to(DEV), np.full(1, V, np.int32),
                                                g["view_choice"][i:i + 1, :3], iters=3)
        np.testing.assert_allclose(t2n(ax3)[0], oax, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(t2n(c3)[0], oc, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(t2n(R3)[0].reshape(3, 3), oR, rtol=1e-4, atol=1e-5)
