#!/usr/bin/env python
"""Generates tests/golden/scale_init.npz by importing the reference's own SLAM/gaussian_pointcloud.py and SLAM/utils.py
(unmodified, from /root/reference) in THIS container and running

  * bbox_filter                          (SLAM/utils.py:801-808)
  * GaussianPointCloud.get_radius        (SLAM/gaussian_pointcloud.py:739-743)
  * GaussianPointCloud.update_geometry   (SLAM/gaussian_pointcloud.py:519-570)

on seeded inputs.  The class is instantiated without its constructor (which wants a full config object) and only the
attributes those methods read are set; `delete` is intercepted to record the mask.  update_geometry calls
`distCUDA2(total_xyz.float().cuda())` -- the CUDA extension cannot run in this GPU-less container, so
`simple_knn._C.distCUDA2` is served by oracle.knn (the C restatement of submodules/simple-knn that tests/golden/knn_4000.npz
pins bit-exact to the compiled reference) and `Tensor.cuda()` is the identity.  Packages the module imports but these
methods never touch (plyfile, open3d, pytorch3d, cv2, matplotlib ...) get empty stand-ins; device="cuda" tensor factories are
routed to the CPU.  No reference source is modified or copied.  Run:  python tests/golden/make_scale_init_golden.py"""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("DQO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "scale_init.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda k: mock.MagicMock()
    m.__path__ = []
    sys.modules[name] = m
    return m


def load_reference():
    from oracle import oracle

    def dist_cuda2(points):
        md, ki = oracle.knn(points.detach().cpu().numpy().astype(np.float32))
        return torch.from_numpy(np.asarray(md)), torch.from_numpy(np.asarray(ki))

    for n in ["cv2", "open3d", "plyfile", "pytorch3d", "pytorch3d.loss", "pytorch3d.ops", "skimage", "skimage.color",
              "skimage.filters", "PIL", "yaml", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits",
              "mpl_toolkits.mplot3d", "imgviz", "torchmetrics"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    sys.modules["plyfile"].PlyData = mock.MagicMock()
    sys.modules["plyfile"].PlyElement = mock.MagicMock()
    _stub("simple_knn")
    _stub("simple_knn._C", distCUDA2=dist_cuda2)

    def cpuify(fn):
        def wrapped(*a, **k):
            if k.get("device", None) == "cuda":
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    saved = (torch.tensor, torch.eye, torch.zeros, torch.Tensor.cuda)
    torch.tensor, torch.eye, torch.zeros = cpuify(torch.tensor), cpuify(torch.eye), cpuify(torch.zeros)
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    try:
        spec = importlib.util.spec_from_file_location("ref_gpc", os.path.join(REF, "SLAM/gaussian_pointcloud.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        torch.tensor, torch.eye, torch.zeros = saved[:3]
    # Tensor.cuda stays patched while the golden vectors are produced (update_geometry calls it)
    return mod, saved[3]


def make_case(seed, n_new, n_extra, spacing, smax=0.02):
    """New points on a noisy plane patch (like back-projected depth samples), existing Gaussians around and beyond it."""
    g = torch.Generator().manual_seed(seed)
    side = int(np.ceil(np.sqrt(n_new)))
    u, v = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
    pts = torch.stack([u.reshape(-1), v.reshape(-1)], 1)[:n_new].float() * spacing
    xyz = torch.cat([pts + 0.3 * spacing * torch.rand(n_new, 2, generator=g), 2.0 + 0.02 * torch.randn(n_new, 1, generator=g)], 1)
    log_scaling = torch.log(0.1 * smax + smax * torch.rand(n_new, 3, generator=g))
    ext = side * spacing
    extra_xyz = torch.cat([(torch.rand(n_extra, 2, generator=g) * 1.6 - 0.3) * ext,
                           2.0 + 0.08 * torch.randn(n_extra, 1, generator=g)], 1)
    extra_radius = 0.05 * smax + 0.5 * smax * torch.rand(n_extra, generator=g)
    return xyz.contiguous(), log_scaling.contiguous(), extra_xyz.contiguous(), extra_radius.contiguous()


def main():
    mod, real_cuda = load_reference()
    try:
        out = {}
        cases = {"a": (11, 3000, 1500, 0.05, 0.013), "b": (12, 5000, 0, 0.05, 0.02), "c": (13, 800, 4000, 0.004, 0.02),
                 "d": (14, 40, 10, 0.5, 0.02)}
        for name, (seed, n_new, n_extra, spacing, smax) in cases.items():
            xyz, log_scaling, extra_xyz, extra_radius = make_case(seed, n_new, n_extra, spacing, smax)
            pc = object.__new__(mod.GaussianPointCloud)
            pc.setup_functions()
            pc._xyz, pc._scaling = xyz.clone(), log_scaling.clone()
            pc.min_radius, pc.max_radius, pc.scale_factor = 0.001, 0.05, 1.0
            pc.xyz_factor = torch.tensor([1.0, 1.0, 0.1])
            deleted = {}
            pc.delete = lambda mask, _d=deleted: _d.__setitem__("mask", mask.clone())
            radius = pc.get_radius.clone()
            pc.update_geometry(extra_xyz.clone(), extra_radius.clone())
            changed = not torch.equal(pc._scaling, log_scaling)
            out[name + "_xyz"], out[name + "_log_scaling"] = xyz.numpy(), log_scaling.numpy()
            out[name + "_extra_xyz"], out[name + "_extra_radius"] = extra_xyz.numpy(), extra_radius.numpy()
            out[name + "_radius"] = radius.numpy()
            out[name + "_inbbox"] = (mod.bbox_filter(xyz, extra_xyz).numpy() if n_extra else np.zeros(0, bool))
            out[name + "_invalid"] = deleted["mask"].numpy()
            out[name + "_scaling_updated"] = np.array(changed)
            out[name + "_new_scaling"] = pc._scaling.numpy()
            print(name, "n_new", n_new, "n_extra", n_extra, "in bbox", int(out[name + "_inbbox"].sum()),
                  "invalid", int(deleted["mask"].sum()), "scaling updated", changed)
        np.savez_compressed(OUT, **out)
        print("wrote", OUT)
    finally:
        torch.Tensor.cuda = real_cuda


if __name__ == "__main__":
    main()
