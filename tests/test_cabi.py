"""CPU: the C-ABI library loads and exports exactly what include/dqo_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from dqo_map_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dqo_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dqo_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.PROTOTYPES) == declared


def _declared_parameter_counts():
    """name -> number of parameters of every prototype in the header (comments stripped; `(void)` = 0)."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    counts = {}
    for m in re.finditer(r"\b(dqo_[a-z0-9_]+)\s*\(", text):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        params = text[m.end():i - 1].strip()
        if text[i:].lstrip()[:1] != ";":
            continue  # not a prototype
        counts[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return counts


def test_binding_argument_counts_match_the_header():
    counts = _declared_parameter_counts()
    assert sorted(counts) == sorted(_lib.PROTOTYPES)
    for name, (_res, args) in _lib.PROTOTYPES.items():
        assert len(args) == counts[name], (name, len(args), counts[name])


def test_library_exports_every_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(h, name), name
    L = _lib.lib()
    assert L.dqo_abi_version() == _lib.ABI_VERSION
    assert L.dqo_last_error() is not None


def test_struct_layouts():
    assert ctypes.sizeof(_lib.RastSettings) == 21 * 4
    assert ctypes.sizeof(_lib.AdamTensor) == 4 * 8 + 8 + 8 + 4 + 4


def test_struct_layouts_match_the_header_as_a_c_compiler_sees_it(tmp_path):
    """Every struct that crosses the boundary: size and the offset of every field, from a C program compiled against
    include/dqo_b200.h (gcc, plain C: the header must stay C-compatible), against the ctypes mirrors in _lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"dqo_rast_settings": _lib.RastSettings, "dqo_adam_tensor": _lib.AdamTensor, "dqo_map_params": _lib.MapParams,
             "dqo_keyframe": _lib.Keyframe}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dqo_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, fname, value = line.split()
        got[(cname, fname)] = int(value)
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def test_invalid_arguments_are_reported_without_a_gpu():
    L = _lib.lib()
    assert L.dqo_mark_visible(-1, None, None, None, None, None) == -1
    assert b"invalid" in L.dqo_last_error()
    assert L.dqo_knn3(-5, None, None, None, None, 0, None) == -1
    assert L.dqo_adam_step(None, 0, 1, 0.9, 0.999, 1e-15, None, -1, None) == -1
    # P == 0 short-circuits like the reference (rasterize_points.cu:103,208)
    assert L.dqo_knn3(0, None, None, None, None, 0, None) == 0
    assert L.dqo_quadric_refine(0, 20, 1, None, None, None, None, 0.01, 0.001, 0.01, None, None, None, None, None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdqomap_b200.so")
    with pytest.raises(_lib.DqoError, match="no CPU fallback"):
        _lib.lib()


def test_workspace_size_queries_without_a_gpu():
    """Host-only size queries: the optional loss terms enlarge the step workspace by exactly their scratch, and only when
    asked for (DQO_STEP_TERM_*)."""
    L = _lib.lib()
    P, M, W, H, cap = 100_000, 16, 640, 480, 1 << 20
    N = W * H
    base = L.dqo_mapping_step_workspace_bytes(P, M, W, H, cap, 0)
    with_ssim = L.dqo_mapping_step_workspace_bytes(P, M, W, H, cap, _lib.STEP_TERM_SSIM)
    with_sem = L.dqo_mapping_step_workspace_bytes(P, M, W, H, cap, _lib.STEP_TERM_SEMANTIC)
    both = L.dqo_mapping_step_workspace_bytes(P, M, W, H, cap, _lib.STEP_TERM_SSIM | _lib.STEP_TERM_SEMANTIC)
    assert 0 < base < with_ssim < both and base < with_sem < both
    ssim_ws = L.dqo_ssim_workspace_bytes(W, H)
    assert ssim_ws >= 9 * N * 4 and abs((with_ssim - base) - ssim_ws) <= 512          # 256-byte alignment slack
    sem_scratch = 28 * N + 44 * P + L.dqo_loss_workspace_bytes(W, H)
    assert abs((with_sem - base) - sem_scratch) <= 6 * 256
    assert abs((both - base) - (ssim_ws + sem_scratch)) <= 8 * 256
    assert L.dqo_ssim_workspace_bytes(0, H) == 0
