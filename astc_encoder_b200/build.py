"""In-tree build of the native library (libastc_b200.so) and the astc_cs_enc CLI.

Everything is compiled for sm_100a only; nvcc cross-compiles without a GPU.
The float flags matter for parity: -fmad=false (no implicit contraction; every
FMA in the kernels is an explicit intrinsic), IEEE div/sqrt (nvcc defaults,
never --use_fast_math).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libastc_b200.so"
CLI = PKG / "bin" / "astc_cs_enc"
SOURCES = ("astc_kernels.cu", "astc_capi.cu", "astc_context.cu", "image_io.cpp", "jpeg_io.cpp", "image_formats.cpp")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    # -fwrapv: the image decoders follow stb_image's int arithmetic, which wraps on corrupt input (e.g. the JPEG IDCT
    # on garbage coefficients); wrapping is made defined behaviour instead of undefined
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-fwrapv", "-I", str(INCLUDE), "-I", str(CSRC),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the sm_100a extension cannot be built")


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); the system compiler is the known-good one
    return "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(map(str, cmd)), flush=True)
    res = subprocess.run(list(map(str, cmd)), capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd[:3]))} ...")
    if verbose and (res.stdout.strip() or res.stderr.strip()):
        print((res.stdout + res.stderr).strip())


def build_variant(tag: str, defines, verbose: bool = False) -> Path:
    """Experiment builds (tools/variants.py): libastc_b200_<tag>.so with extra -D flags.
    Only the kernel TU differs; never used by the product path."""
    nvcc = _nvcc()
    objdir = PKG / "_obj"
    objdir.mkdir(exist_ok=True)
    lib = PKG / f"libastc_b200_{tag}.so"
    objs = []
    for src in SOURCES:
        if src.endswith(".cu"):                              # the CUDA translation units see the experiment's defines
            obj = objdir / f"{src.rsplit('.', 1)[0]}_{tag}.o"
            _run([nvcc, "-ccbin", _host_cxx(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", CSRC / src, "-o", obj], verbose)
        else:
            obj = objdir / (src.rsplit(".", 1)[0] + ".o")
            if not obj.exists():
                _run([nvcc, "-ccbin", _host_cxx(), *NVCC_FLAGS, "-c", CSRC / src, "-o", obj], verbose)
        objs.append(obj)
    _run([nvcc, "-ccbin", _host_cxx(), "-shared", "-o", lib, *objs, "-lz"], verbose)
    for obj in objs:
        if obj.name.endswith(f"_{tag}.o"):
            obj.unlink()
    return lib


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    nvcc = _nvcc()
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        list(CSRC.glob("*.inc")) + list(CSRC.glob("*.cpp")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    objdir = PKG / "_obj"
    if force or _stale(LIB, deps):
        objdir.mkdir(exist_ok=True)
        objs = []
        for src in SOURCES:
            obj = objdir / (src.rsplit(".", 1)[0] + ".o")
            if force or _stale(obj, deps):
                extra = ["-Xptxas", "-v"] if (ptxas_info and src.endswith(".cu")) else []
                _run([nvcc, "-ccbin", _host_cxx(), *NVCC_FLAGS, *extra, "-c", CSRC / src, "-o", obj],
                     verbose or ptxas_info)
            objs.append(obj)
        _run([nvcc, "-ccbin", _host_cxx(), "-shared", "-o", LIB, *objs, "-lz"], verbose)
    if force or _stale(CLI, deps + [LIB]):
        CLI.parent.mkdir(exist_ok=True)
        _run([_host_cxx(), "-std=c++17", "-O2", "-I", INCLUDE, CSRC / "astc_cs_enc.cpp", "-o", CLI,
              f"-L{PKG}", "-lastc_b200", "-Wl,-rpath,$ORIGIN/.."], verbose)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print(LIB)
