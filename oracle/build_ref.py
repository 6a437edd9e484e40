"""Build the reference's own hash-grid CUDA kernels for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY. The sources are compiled where they lie under
/root/reference/wisp/csrc/ops (hashgrid_interpolate.cpp, hashgrid_interpolate_cuda.cu,
hashgrid_interpolate2d_cuda.cu); nothing is copied into the repo. The only accommodation
for torch 2.11 is the pre-included oracle/ref_compat.h. Output: oracle/_ref/wisp_ref_ops.so
(git-ignored, travels to the GPU box with the gpurun snapshot).

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box)
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OPS = "/root/reference/wisp/csrc/ops"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "wisp_ref_ops.so")
NAME = "wisp_ref_ops"


def _stale(srcs):
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(verbose=False):
    if not os.path.isdir(REF_OPS):
        return OUT if os.path.exists(OUT) else None
    import torch
    from torch.utils import cpp_extension as ce

    srcs_cu = [os.path.join(REF_OPS, f) for f in ("hashgrid_interpolate_cuda.cu", "hashgrid_interpolate2d_cuda.cu")]
    src_cpp = os.path.join(REF_OPS, "hashgrid_interpolate.cpp")
    binding = os.path.join(HERE, "ref_binding.cpp")
    compat = os.path.join(HERE, "ref_compat.h")
    if not _stale(srcs_cu + [src_cpp, binding, compat, os.path.abspath(__file__)]):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = os.path.join("/tmp", "shacira_ref_build_%d" % os.getpid())
    os.makedirs(tmp, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"], "-I", REF_OPS]
    defs = ["-DWITH_CUDA", "-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    objs = []
    for s in srcs_cu:
        o = os.path.join(tmp, os.path.basename(s) + ".o")
        cmd = ["nvcc", "-c", s, "-o", o, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
               "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-include", compat] + inc + defs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    for s in (src_cpp, binding):
        o = os.path.join(tmp, os.path.basename(s) + ".o")
        cmd = ["g++", "-c", s, "-o", o, "-O3", "-std=c++17", "-fPIC"] + inc + defs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", OUT] + objs
    for d in libdirs:
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    return OUT


def load():
    """Import the built module (needs torch imported first). Returns None if absent."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(NAME, OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv)
    print("built" if out else "skipped (no /root/reference)", out)
