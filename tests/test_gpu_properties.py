"""GPU property tests (hypothesis): size-independent invariants of the CUDA path on drawn inputs -- sortedness and
stability of the radix sort, the loss kernels against torch on arbitrary (ragged, unaligned) image sizes and masks, the
rasterizer's forward outputs being independent of how the binning is split into phases and of Gaussian order-preserving
padding, SSIM against its float64 oracle on arbitrary small shapes."""
import os
import sys

import pytest
import torch
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import refharness as rh  # noqa: E402
from oracle import ssim_oracle as so  # noqa: E402
from dqo_map_b200 import mapping, rasterizer  # noqa: E402
from test_gpu_sort import reference_sort, run_sort  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DRAW = settings(max_examples=20, deadline=None, derandomize=True)


@DRAW
@given(st.integers(1, 200_000), st.integers(1, 32), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_sort_is_a_stable_permutation_for_any_size_and_width(n, bits, seed, few_keys):
    g = torch.Generator(device="cpu").manual_seed(seed)
    hi = min(1 << min(bits, 31), 37) if few_keys else (1 << min(bits + 1, 31))
    keys = torch.randint(0, hi, (n,), generator=g, dtype=torch.int64).to(torch.int32).to(DEV)
    ks, vs, vals_in = run_sort(keys, bits)
    rk, rv = reference_sort(keys, bits, n, vals_in)
    digits = ks[:n].to(torch.int64) & ((1 << bits) - 1)
    assert bool((digits[1:] >= digits[:-1]).all())                         # sorted by the low `bits` bits
    assert torch.equal(ks[:n].to(torch.int64) & 0xFFFFFFFF, rk) and torch.equal(vs[:n], rv)   # and stable


@DRAW
@given(st.integers(1, 70), st.integers(1, 90), st.integers(0, 2 ** 31 - 1), st.sampled_from(["none", "random", "empty", "full"]),
       st.sampled_from([0.0, 1.0]))
def test_masked_l1_loss_matches_torch_for_any_shape_and_mask(H, W, seed, mask_kind, depth_weight):
    g = torch.Generator(device="cpu").manual_seed(seed)
    image, depth = torch.rand(3, H, W, generator=g).to(DEV), (torch.rand(1, H, W, generator=g) * 3).to(DEV)
    hit = torch.randint(-1, 9, (1, H, W), generator=g, dtype=torch.int32).to(DEV)
    gt_c, gt_d = torch.rand(H, W, 3, generator=g).to(DEV), (torch.rand(H, W, 1, generator=g) * 3).to(DEV)
    gt_d[gt_d < 0.4] = 0.0
    mask = {"none": None, "random": (torch.rand(H, W, generator=g) < 0.5).to(DEV),
            "empty": torch.zeros(H, W, dtype=torch.bool, device=DEV), "full": torch.ones(H, W, dtype=torch.bool, device=DEV)}[mask_kind]
    x, d = image.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    total, lc, ld, counts = mapping.masked_l1_loss(x, d, hit, gt_c, gt_d, mask, 0.8, depth_weight, 0.1)
    # loss_update's own expressions (mapper.py:847-857)
    xr, dr = image.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    m = torch.ones(H, W, dtype=torch.bool, device=DEV) if mask is None else mask
    color_loss = torch.abs(xr.permute(1, 2, 0)[m] - gt_c[m]).mean()
    err = dr.permute(1, 2, 0) - gt_d
    valid = (hit.permute(1, 2, 0) != -1).squeeze(-1) & (gt_d > 0).squeeze(-1) & (err < 0.1).squeeze(-1) & m
    depth_loss = torch.abs(err[valid]).mean() if depth_weight > 0 else torch.zeros((), device=DEV)
    ref = depth_weight * depth_loss + 0.8 * color_loss
    assert int(counts[0]) == int(m.sum()) and (depth_weight == 0 or int(counts[1]) == int(valid.sum()))
    if bool(torch.isnan(ref)):
        assert bool(torch.isnan(total))                                     # empty selection: NaN like torch.mean
        return
    assert abs(float(total.detach()) - float(ref.detach())) <= 2e-6 * max(1.0, abs(float(ref.detach())))
    total.backward()
    ref.backward()
    assert float((x.grad - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max()) + 1e-12
    if depth_weight > 0:
        assert float((d.grad - dr.grad).abs().max()) <= 1e-5 * float(dr.grad.abs().max()) + 1e-12


@DRAW
@given(st.integers(1, 60), st.integers(1, 60), st.integers(0, 2 ** 31 - 1))
def test_ssim_loss_matches_the_float64_oracle_for_any_shape(H, W, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    img = (0.6 * gt + 0.4 * torch.rand(3, H, W, generator=g)).clamp(0, 1)
    gt_hwc = gt.permute(1, 2, 0).contiguous()
    l64, g64 = so.ssim_loss_and_grad(img, gt_hwc)
    x = img.to(DEV).requires_grad_(True)
    loss = mapping.ssim_loss(x, gt_hwc.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - l64) <= 2e-6
    scale = float(g64.abs().max())
    assert float((x.grad.cpu().double() - g64).abs().max()) <= 5e-5 * scale + 1e-12
    # identical images: ssim = 1 exactly up to rounding, loss ~ 0
    same = mapping.ssim_loss(gt.to(DEV), gt_hwc.to(DEV))
    assert abs(float(same)) <= 2e-6


FWD_EXACT = (2, 3, 4, 5, 6, 7, 8, 9, 13, 14)   # colour, depth, index maps, hit weights, T, radii, tile list, n_touched


@settings(max_examples=8, deadline=None, derandomize=True)
@given(st.integers(300, 6000), st.integers(0, 2 ** 31 - 1), st.floats(0.01, 0.95), st.sampled_from(["ones", "half"]))
def test_forward_outputs_do_not_depend_on_the_binning_split(P, seed, front_frac, mask):
    """Any front / back split of the occlusion-aware binning reproduces the single-phase forward bit for bit (the front
    list followed by the back list IS the reference's list), on drawn scenes, sizes and split points."""
    inp = rh.make_inputs("small", torch.device(DEV), seed=seed % 100000, P=P, sh_degree=1, mask=mask)
    try:
        rasterizer.set_binning_mode("single")
        o1 = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
        R = o1[0]
        if R < 512:
            return
        front = max(256, int(R * front_frac) // 256 * 256)
        rasterizer.set_binning_mode("fixed", front, R + 1024)
        o2 = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
        assert o2[0] == o1[0] and o2[1] == o1[1]
        for i in FWD_EXACT:
            assert torch.equal(o1[i], o2[i]), i
    finally:
        rasterizer.set_binning_mode("single")


@settings(max_examples=6, deadline=None, derandomize=True)
@given(st.integers(300, 4000), st.integers(0, 2 ** 31 - 1), st.integers(1, 500))
def test_culled_gaussians_appended_to_the_cloud_change_nothing(P, seed, n_extra):
    """Gaussians behind the camera are culled in the preprocess (forward.cu:in_frustum): appending any number of them
    leaves every image bit-identical and gives them radius 0 and zero touches."""
    inp = rh.make_inputs("small", torch.device(DEV), seed=seed % 100000, P=P, sh_degree=1)
    rasterizer.set_binning_mode("single")
    o1 = rasterizer.rasterize_gaussians(*rh.raster_args(inp))
    cam = inp["cam"]
    back = cam.camera_center[None] - 5.0 * cam.world_view_transform[:3, 2][None]    # 5 m behind the camera
    ext = dict(inp)
    ext["xyz"] = torch.cat([inp["xyz"], back.repeat(n_extra, 1)]).contiguous()
    for k in ("opacity", "scales", "rotations", "shs"):
        ext[k] = torch.cat([inp[k], inp[k][:1].repeat(n_extra, *([1] * (inp[k].dim() - 1)))]).contiguous()
    o2 = rasterizer.rasterize_gaussians(*rh.raster_args(ext))
    assert o2[0] == o1[0] and o2[1] == o1[1]
    for i in (2, 3, 4, 5, 6, 7, 8, 13):
        assert torch.equal(o1[i], o2[i]), i
    assert torch.equal(o2[9][:P], o1[9]) and int(o2[9][P:].abs().sum()) == 0
    assert torch.equal(o2[14][:P], o1[14]) and int(o2[14][P:].abs().sum()) == 0
