"""CPU: host-side mirror of the reference interface (names, argument rules, error behaviour) and sharding logic."""
import os
import subprocess
import sys

import pytest
import torch

from dqo_map_b200 import rasterizer, sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _settings():
    cam = synthetic.make_camera("tiny")
    return rasterizer.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.zeros(3), scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=0, campos=cam.camera_center, opaque_threshold=0.6, normal_threshold=0.5, depth_threshold=1.0,
        prefiltered=False, debug=False, cx=cam.cx, cy=cam.cy)


def test_settings_field_order_matches_reference():
    # RAST/diff_gaussian_rasterization_depth/__init__.py:288-307 — field order is part of the API
    assert rasterizer.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "opaque_threshold", "normal_threshold", "depth_threshold", "prefiltered", "debug", "cx",
        "cy", "color_sigma", "T_threshold")
    s = _settings()
    assert s.color_sigma == 3.0 and s.T_threshold == 0.0001


def test_rasterizer_argument_rules():
    r = rasterizer.GaussianRasterizer(_settings())
    x = torch.zeros(4, 3)
    o = torch.zeros(4, 1)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=x, opacities=o, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=x, opacities=o, shs=torch.zeros(4, 1, 3), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=x, opacities=o, colors_precomp=x)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=x, opacities=o, colors_precomp=x, scales=x, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))
    # CPU tensors are refused loudly: there is no CPU path
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(means3D=x, opacities=o, colors_precomp=x, scales=x, rotations=torch.zeros(4, 4),
          tile_mask=torch.ones(6, 10, dtype=torch.int32))


def test_camera_conventions():
    cam = synthetic.make_camera("tiny")
    # transposed (column-major) matrices: last column of the stored matrix is (0,0,0,1)^T of W2C^T
    assert torch.allclose(cam.world_view_transform[:, 3], torch.tensor([0.0, 0.0, 0.0, 1.0]))
    full = cam.world_view_transform @ cam.projection_matrix
    assert torch.allclose(full, cam.full_proj_transform)
    # camera centre maps to the origin of the view frame
    c = torch.cat([cam.camera_center, torch.ones(1)])
    assert torch.allclose((c @ cam.world_view_transform)[:3], torch.zeros(3), atol=1e-5)


def test_dropin_packages_expose_reference_names():
    code = (
        "import sys; sys.path.insert(0, %r);"
        "from diff_gaussian_rasterization_depth import GaussianRasterizationSettings, GaussianRasterizer;"
        "from diff_gaussian_rasterization_depth import _C_depth;"
        "assert all(hasattr(_C_depth, n) for n in ('rasterize_gaussians','rasterize_gaussians_backward','mark_visible'));"
        "from simple_knn._C import distCUDA2;"
        "from cuda_utils._C import accumulate_gaussian_error;"
        "print('ok')" % os.path.join(ROOT, "dqo-map_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert "ok" in out.stdout


def test_assign_objects_is_balanced_and_deterministic():
    counts = {i: c for i, c in enumerate([900, 500, 400, 300, 300, 200, 100, 50, 50, 10])}
    owner, load = sharding.assign_objects(counts, 4)
    assert sorted(owner) == sorted(counts)
    assert sum(load) == sum(counts.values())
    assert max(load) <= 900  # the largest object bounds the makespan here
    assert sharding.assign_objects(counts, 4) == (owner, load)
    owner1, load1 = sharding.assign_objects(counts, 1)
    assert set(owner1.values()) == {0} and load1 == [sum(counts.values())]
    assert sharding.local_objects(owner, 0) == sorted(o for o, r in owner.items() if r == 0)


def test_fused_adam_accepts_reference_param_groups():
    from dqo_map_b200.mapping import FusedAdam
    ps = {k: torch.nn.Parameter(torch.zeros(s)) for k, s in
          dict(xyz=(5, 3), f_dc=(5, 1, 3), f_rest=(5, 15, 3), opacity=(5, 1), scaling=(5, 3), rotation=(5, 4)).items()}
    groups = [{"params": [ps[k]], "lr": lr, "name": k} for k, lr in
              dict(xyz=1e-3, f_dc=5e-4, f_rest=2.5e-5, opacity=0.0, scaling=4e-3, rotation=1e-3).items()]
    opt = FusedAdam(groups, lr=0.0, eps=1e-15)
    assert [g["name"] for g in opt.param_groups] == ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
    assert opt.step() is None  # no gradients -> nothing to do, no GPU touched


# ---------------------------------------------------------------------------------------------------------------------
# binning policy (pure host logic)
# ---------------------------------------------------------------------------------------------------------------------
def _status(R=0, overflow=0, r_front=0, r_back=0, walked=0, unfinished=0):
    return [R, 100, overflow, 1000, r_front, r_back, walked, unfinished]


def test_binning_policy_stays_single_for_small_or_fully_walked_scenes():
    from dqo_map_b200.binning_policy import BinningPolicy
    p = BinningPolicy()
    assert p.plan(1 << 30) == (0, 0)
    assert p.update(_status(R=500_000, walked=10_000), 0, 0) is False     # too small to bother
    assert p.plan(1 << 30) == (0, 0)
    assert p.update(_status(R=15_000_000, walked=12_000_000), 0, 0) is False  # nearly everything is read
    assert p.plan(1 << 30) == (0, 0)
    assert p.update(_status(R=15_000_000, overflow=1), 0, 0) is True


def test_binning_policy_enters_tightens_and_leaves_two_phase():
    from dqo_map_b200.binning_policy import BinningPolicy
    p = BinningPolicy()
    R = 15_000_000
    p.update(_status(R=R, walked=1_200_000), 0, 0)
    front, back = p.plan(1 << 30)
    assert front % 256 == 0 and 2_000_000 <= front <= R // 2
    assert front + back >= R  # first two-phase call cannot overflow
    # observed: few unfinished tiles -> the back region shrinks to what was needed plus headroom
    assert p.update(_status(R=R, r_front=front - 100, r_back=300_000, unfinished=40), front, back) is False
    f2, b2 = p.plan(1 << 30)
    assert f2 == front and 300_000 < b2 < back
    # a harder keyframe overflows the back region: repeat with room for what it reported
    assert p.update(_status(R=R, overflow=1, r_front=front - 100, r_back=2 * b2), f2, b2) is True
    f3, b3 = p.plan(1 << 30)
    assert b3 > 2 * b2
    # a scene where nothing finishes early: give up and do not probe again immediately
    assert p.update(_status(R=R, r_front=f3, r_back=R - f3, unfinished=3000), f3, b3) is False
    assert p.plan(1 << 30) == (0, 0)
    p.update(_status(R=R, walked=1_000_000), 0, 0)
    assert p.plan(1 << 30) == (0, 0)  # cooling down
    # capacity too small for the plan -> clipped or single phase, never an invalid pair
    p2 = BinningPolicy()
    p2.update(_status(R=R, walked=1_200_000), 0, 0)
    f, b = p2.plan(1 << 20)
    assert (f, b) == (0, 0) or (f % 256 == 0 and f + b <= 1 << 20)


def test_object_scene_table_is_deterministic_and_packs_evenly():
    """The multi-object workload (bench_objects.py): every rank derives the same object table and the same owner map."""
    from dqo_map_b200 import sharding, synthetic
    c1, c2 = synthetic.object_counts(20), synthetic.object_counts(20)
    assert c1 == c2 and len(c1) == 21 and c1[0] == 200_000 and all(20_000 <= c <= 80_000 for c in c1[1:])
    for world in (1, 2, 4, 8):
        owner, load = sharding.assign_objects(c1, world)
        assert sorted(owner) == list(range(21)) and sum(load) == sum(c1)
        assert sorted(o for r in range(world) for o in sharding.local_objects(owner, r)) == list(range(21))
        # LPT: no rank exceeds the mean by more than the largest object
        assert max(load) <= sum(c1) / world + max(c1)
    obj = synthetic.make_object(3, 500)
    assert obj["xyz"].shape == (500, 3) and obj["shs"].shape == (500, 16, 3) and obj["obj_id"] == 3
    assert torch.equal(obj["xyz"], synthetic.make_object(3, 500)["xyz"])


def test_integration_snippets_are_valid_python():
    """The binding stubs shown in INTEGRATION.md / README.md must at least parse."""
    import ast
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 0
    for doc in ("INTEGRATION.md",):
        text = open(os.path.join(root, doc)).read()
        for block in re.findall(r"```python\n(.*?)```", text, flags=re.S):
            ast.parse(block)
            n += 1
    assert n >= 3
