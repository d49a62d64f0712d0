#!/usr/bin/env python
"""Generates tests/golden/maps.npz by importing the reference's own SLAM/utils.py (unmodified, from /root/reference) in
THIS container and running transmission2tilemask (utils.py:752-763) and colorerror2tilemask (utils.py:765-796) on seeded
inputs.  The module imports packages that are not installed here (open3d, pytorch3d, plyfile, skimage, cv2) and
utils/general_utils.py creates its dtype sentinels on "cuda"; the script substitutes empty stand-in modules for the
former and routes device="cuda" tensor factories to the CPU.  No reference source is modified or copied.
Run:  python tests/golden/make_maps_golden.py"""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = os.environ.get("DQO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps.npz")


def _stub(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda k: mock.MagicMock()
    m.__path__ = []
    sys.modules[name] = m


def load_reference_utils():
    for n in ["cv2", "open3d", "plyfile", "pytorch3d", "pytorch3d.loss", "pytorch3d.ops", "skimage", "skimage.color",
              "skimage.filters", "PIL", "yaml"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    real_tensor = torch.tensor

    def cpu_tensor(*a, **k):
        if k.get("device", None) == "cuda":
            k["device"] = "cpu"
        return real_tensor(*a, **k)

    torch.tensor = cpu_tensor
    sys.path.insert(0, REF)
    try:
        spec = importlib.util.spec_from_file_location("ref_slam_utils", os.path.join(REF, "SLAM/utils.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        torch.tensor = real_tensor
        sys.path.remove(REF)
    return mod


def main():
    ref = load_reference_utils()
    g = torch.Generator().manual_seed(2024)
    out = {}
    for name, (H, W) in {"a": (48, 64), "b": (37, 53), "c": (200, 330)}.items():
        # transmission-style mask: blobs of rendered pixels
        base = torch.rand((H + 15) // 16, (W + 15) // 16, generator=g)
        up = torch.nn.functional.interpolate(base[None, None], size=(H, W), mode="bilinear")[0, 0]
        pixelmask = (up + 0.25 * torch.rand(H, W, generator=g)) > 0.6
        for ratio in (0.5, 0.3):
            out["%s_tm_%02d" % (name, int(ratio * 10))] = ref.transmission2tilemask(pixelmask, 16, ratio).numpy()
        out[name + "_pixelmask"] = pixelmask.numpy()
        err = torch.rand(H, W, generator=g) * up
        err[~pixelmask] = 0
        for ratio in (0.4, 0.1):
            out["%s_ce_%02d" % (name, int(ratio * 10))] = ref.colorerror2tilemask(err, 16, ratio).numpy()
        out[name + "_error"] = err.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
