#!/usr/bin/env python
"""Build the UNMODIFIED reference operators into oracle/_ref/ (test infrastructure only).

This is the recipe the task calls "oracle/_ref": the reference's three CUDA extensions
(`submodules/diff-gaussian-rasterizer-depth`, `submodules/simple-knn`, `submodules/cuda_utils`)
are compiled for sm_100a *from the sources where they lie* under /root/reference.  Nothing is
patched or copied into the repository history:

  * the two missing-include problems SURVEY.md §8c found (`<cstdint>` in rasterizer_impl.h,
    `<cfloat>/<climits>` in simple_knn.cu) are solved with nvcc `-include` flags, not source edits;
  * outputs (objects, .so files and an "installed" copy of the reference's own python wrapper
    `diff_gaussian_rasterization_depth/__init__.py`, exactly what `pip install` would place in
    site-packages) go only to oracle/_ref/, which is git-ignored but travels to the GPU box.

The reference has no CPU implementation of this path (it is CUDA only), so oracle/_ref can only be
*executed* on the GPU box.  It is used (a) to generate tests/golden/*.npz (tests/golden/make_golden.py),
(b) as the live differential oracle in `pytest -m gpu`, and (c) by `bench.py --impl reference`.
The product never imports it.

Usage: python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("DQO_REFERENCE_ROOT", "/root/reference")
SUB = os.path.join(REF, "submodules")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _torch_flags():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce

    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    libdirs = ce.library_paths("cuda")
    return inc, libdirs, list(ce.COMMON_NVCC_FLAGS)


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _build_ext(name, modname, srcdir, sources, extra_nvcc, pkg_dir):
    inc, libdirs, common = _torch_flags()
    objdir = os.path.join(OUT, "obj", name)
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(pkg_dir, exist_ok=True)
    defs = ["-DTORCH_EXTENSION_NAME=" + modname, "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1"]
    jobs = []
    objs = []
    for s in sources:
        src = os.path.join(srcdir, s)
        obj = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(obj)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", src, "-o", obj, "-O3", "-std=c++17", "--compiler-options", "-fPIC"] + ARCH + common + defs + inc + extra_nvcc
        else:
            cmd = ["g++", "-c", src, "-o", obj, "-O2", "-std=c++17", "-fPIC"] + defs + inc
        jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        list(ex.map(_run, jobs))
    so = os.path.join(pkg_dir, modname + ".so")
    link = ["g++", "-shared", "-o", so] + objs
    for d in libdirs:
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    _run(link)
    return so


def build(force=False):
    if not os.path.isdir(SUB):
        print("reference sources not present at %s: keeping prebuilt oracle/_ref (if any)" % SUB)
        return False
    rast_dir = os.path.join(SUB, "diff-gaussian-rasterizer-depth")
    knn_dir = os.path.join(SUB, "simple-knn")
    cu_dir = os.path.join(SUB, "cuda_utils")
    rast_pkg = os.path.join(OUT, "diff_gaussian_rasterization_depth")
    knn_pkg = os.path.join(OUT, "simple_knn")
    cu_pkg = os.path.join(OUT, "cuda_utils")
    done = os.path.join(OUT, ".built")
    if os.path.exists(done) and not force:
        print("oracle/_ref already built")
        return True
    os.makedirs(OUT, exist_ok=True)
    _build_ext(
        "rast", "_C_depth", rast_dir,
        ["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu", "cuda_rasterizer/backward.cu",
         "rasterize_points.cu", "ext.cpp"],
        ["-I", os.path.join(rast_dir, "third_party/glm"), "-include", "cstdint", "-lineinfo"], rast_pkg)
    # "install" the reference's own python wrapper next to its extension (what pip install does)
    shutil.copyfile(os.path.join(rast_dir, "diff_gaussian_rasterization_depth", "__init__.py"),
                    os.path.join(rast_pkg, "__init__.py"))
    _build_ext("knn", "_C", knn_dir, ["spatial.cu", "simple_knn.cu", "ext.cpp"],
               ["-include", "cfloat", "-include", "climits"], knn_pkg)
    open(os.path.join(knn_pkg, "__init__.py"), "w").close()
    _build_ext("cu", "_C", cu_dir, ["cuda_utils.cu", "map_process.cu", "ext.cpp"], ["-I", cu_dir], cu_pkg)
    open(os.path.join(cu_pkg, "__init__.py"), "w").close()
    open(done, "w").write("ok\n")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
