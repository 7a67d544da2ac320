"""Build the UNMODIFIED reference pointnet2_ops CUDA extension for sm_100 into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under rfdnet_b200/ may import this.

The reference's own build (external/pointnet2_ops_lib/setup.py:19,
pointnet2_ops/pointnet2_utils.py:23) pins TORCH_CUDA_ARCH_LIST="3.7+PTX;...;7.5",
which nvcc 12.9 rejects and which has no sm_100 target, so its build system is
not run.  Instead the sources are compiled where they lie under
/root/reference (never copied into this repo) with a plain nvcc/g++ recipe and
the result is written only to oracle/_ref/ (git-ignored, travels to the GPU box).

The resulting module `_ref_ext` exposes the reference's 9 pybind functions
(_ext-src/src/bindings.cpp:6-19).  It is the GPU ground truth for the bit-exact
index tests and the "reference kernels recompiled for sm_100" timing baseline
(BASELINE.md §2a).  It needs a GPU to run; there is no CPU path in the reference
(ball_query.cpp:27-29 `AT_ASSERT(false, "CPU not supported")`).
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/external/pointnet2_ops_lib/pointnet2_ops/_ext-src"
OUT = os.path.join(HERE, "_ref")
MODNAME = "_ref_ext"


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("reference build step failed")


REF_PY = "/root/reference/external/pointnet2_ops_lib/pointnet2_ops"
PY_OUT = os.path.join(OUT, "py", "external", "pointnet2_ops_lib", "pointnet2_ops")


def stage_py():
    """Stage the reference's UNMODIFIED pointnet2_utils.py / pointnet2_modules.py under the git-ignored oracle/_ref/py/
    (same package path as in the reference tree) so that the GPU box -- which has no /root/reference -- can execute the
    reference's own Python over `rfdnet_b200.dropin.install()` and over the reference CUDA kernels (tests/test_gpu_dropin.py).
    The files are build artefacts like _ref_ext.so: never committed."""
    if not os.path.isdir(REF_PY):
        return None
    import shutil
    os.makedirs(PY_OUT, exist_ok=True)
    for f in ("pointnet2_utils.py", "pointnet2_modules.py"):
        shutil.copyfile(os.path.join(REF_PY, f), os.path.join(PY_OUT, f))
    return PY_OUT


def load_py(ext, tag):
    """Import the staged reference modules with `pointnet2_ops._ext` bound to `ext` (ours or _ref_ext).
    -> (pointnet2_utils, pointnet2_modules) as fresh module objects named *_<tag>; None if not staged."""
    import importlib.util
    import types
    if not os.path.exists(os.path.join(PY_OUT, "pointnet2_modules.py")):
        return None
    saved = {k: sys.modules.get(k) for k in ("pointnet2_ops", "pointnet2_ops._ext", "external", "external.pointnet2_ops_lib",
                                              "external.pointnet2_ops_lib.pointnet2_ops")}
    try:
        pkg = types.ModuleType("pointnet2_ops")
        pkg.__path__ = []
        pkg._ext = ext
        sys.modules["pointnet2_ops"], sys.modules["pointnet2_ops._ext"] = pkg, ext
        spec = importlib.util.spec_from_file_location("ref_pointnet2_utils_" + tag, os.path.join(PY_OUT, "pointnet2_utils.py"))
        utils = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(utils)       # executes `import pointnet2_ops._ext as _ext` (pointnet2_utils.py:8)
        assert utils._ext is ext
        for name in ("external", "external.pointnet2_ops_lib", "external.pointnet2_ops_lib.pointnet2_ops"):
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        sys.modules["external.pointnet2_ops_lib.pointnet2_ops"].pointnet2_utils = utils
        spec = importlib.util.spec_from_file_location("ref_pointnet2_modules_" + tag, os.path.join(PY_OUT, "pointnet2_modules.py"))
        mods = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods)        # `from external.pointnet2_ops_lib.pointnet2_ops import pointnet2_utils` (:6)
        assert mods.pointnet2_utils is utils
        return utils, mods
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def build(force=False):
    stage_py()
    so = os.path.join(OUT, MODNAME + ".so")
    if os.path.exists(so) and not force:
        return so
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    import torch
    from torch.utils import cpp_extension as ce
    import pybind11

    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    inc = [os.path.join(REF_SRC, "include")] + ce.include_paths() + [
        sysconfig.get_paths()["include"], pybind11.get_include(), "/usr/local/cuda/include"]
    incf = [f"-I{p}" for p in inc]
    defs = [f"-DTORCH_EXTENSION_NAME={MODNAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    srcdir = os.path.join(REF_SRC, "src")
    jobs = []
    objs = []
    for f in sorted(os.listdir(srcdir)):
        src = os.path.join(srcdir, f)
        obj = os.path.join(objdir, f + ".o")
        objs.append(obj)
        if f.endswith(".cu"):
            jobs.append(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100,code=sm_100",
                         "-Xcompiler", "-fPIC", "-c", src, "-o", obj] + incf + defs)
        elif f.endswith(".cpp"):
            jobs.append(["g++", "-O3", "-std=c++17", "-fPIC", "-c", src, "-o", obj] + incf + defs)
    with ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(_run, jobs))
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    _run(["g++", "-shared", "-o", so] + objs + [
        f"-L{tlib}", "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
        "-ltorch", "-ltorch_python", "-lcudart", f"-Wl,-rpath,{tlib}"])
    return so


def load():
    """Import the prebuilt reference module (GPU box or here); None if it was never built."""
    so = os.path.join(OUT, MODNAME + ".so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(MODNAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
