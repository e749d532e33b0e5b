"""Build recipe for the GPU-side reference oracle (TEST INFRASTRUCTURE, never shipped).

Compiles the reference's own pointnet2 `_ext` sources *where they lie* under
/root/reference/pointnet2/_ext_src (nothing is copied into this repo) for sm_100a and
writes one artefact: oracle/_ref/pointnet2/_ext.<abi>.so  (git-ignored, travels to the
GPU box via gpurun).  With oracle/_ref on sys.path, `import pointnet2._ext` then yields
the unmodified reference ops (CUDA only: every op raises "CPU not supported" on CPU
tensors, e.g. pointnet2/_ext_src/src/sampling.cpp:87).

We do not run the reference's setup.py; this is a direct nvcc/g++ recipe with the same
flags it passes (-O2 and the include dir, pointnet2/setup.py:24-27).

The reference's PYTHON side of the path (pointnet2/*.py, models/{bdetr,modules,backbone_module,
encoder_decoder_layers}.py) is needed on the GPU box too (full-model parity tests, the R-GPU baseline leg of bench.py),
and /root/reference does not exist there: `build_pyref` byte-compiles those files where they lie into
oracle/_ref/pyref/{pointnet2,models}/*.pyc.bin (sourceless bytecode, git-ignored, travels like the .so; the suffix is
not ".pyc" because the gpurun snapshot drops *.pyc files — oracle/ref_model.py materialises them as .pyc in a temporary
directory before importing).  No reference source text enters the repository.

Only tests/ and bench.py (R-GPU baseline leg / CPU reference arm) load the results.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

REF_ROOT = "/root/reference"
REF_SRC = "/root/reference/pointnet2/_ext_src"
PYREF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")
PYREF_FILES = {
    "pointnet2": ["pointnet2_modules.py", "pointnet2_utils.py", "pytorch_utils.py"],
    "models": ["bdetr.py", "modules.py", "backbone_module.py", "encoder_decoder_layers.py"],
}
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref", "pointnet2")
OBJ_DIR = os.path.join(HERE, "_ref", "obj")


def so_path():
    return os.path.join(OUT_DIR, "_ext" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=True):
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    out = so_path()
    if os.path.exists(out) and not force:
        return out
    from torch.utils import cpp_extension as ce
    import torch

    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [
        f"-I{REF_SRC}/include",
        f"-I{sysconfig.get_paths()['include']}",
    ]
    defs = ["-DTORCH_EXTENSION_NAME=_ext", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    srcs = sorted(os.listdir(f"{REF_SRC}/src"))
    cmds, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, s + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmds.append(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                         "--compiler-options", "-fPIC", "-D__CUDA_NO_HALF_OPERATORS__",
                         "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                         "--expt-relaxed-constexpr", *defs, *inc, "-c", f"{REF_SRC}/src/{s}", "-o", o])
        elif s.endswith(".cpp"):
            cmds.append(["g++", "-O2", "-std=c++17", "-fPIC", *defs, *inc, "-c", f"{REF_SRC}/src/{s}", "-o", o])

    def run(c):
        if verbose:
            print(" ".join(c[:2]), c[-3], flush=True)
        subprocess.run(c, check=True)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, cmds))
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", *objs, "-o", out] + [f"-L{p}" for p in libdirs] + \
           ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"] + \
           [f"-Wl,-rpath,{p}" for p in libdirs]
    run(link)
    return out


def build_pyref(force=False):
    """Byte-compiles the reference's Python modules on the path into oracle/_ref/pyref (see the module docstring).
    Returns the directory, or None when /root/reference is absent (GPU box: the prebuilt files are used)."""
    import py_compile

    if not os.path.isdir(os.path.join(REF_ROOT, "models")):
        return PYREF_DIR if os.path.isdir(PYREF_DIR) else None
    for sub, names in PYREF_FILES.items():
        os.makedirs(os.path.join(PYREF_DIR, sub), exist_ok=True)
        for name in names:
            src = os.path.join(REF_ROOT, sub, name)
            dst = os.path.join(PYREF_DIR, sub, name + "c.bin")
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                # dfile: tracebacks name the reference file; UNCHECKED_HASH: valid without the source next to it
                py_compile.compile(src, cfile=dst, dfile=f"<reference>/{sub}/{name}", doraise=True,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return PYREF_DIR


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("built" if p else "reference sources absent; nothing built", p)
    print("pyref:", build_pyref(force="--force" in sys.argv))
