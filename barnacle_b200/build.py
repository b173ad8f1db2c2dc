"""Builds barnacle_b200/lib/libbarnacle_b200.so (sm_100a only) in-tree.

    python -m barnacle_b200.build [--force]

nvcc cross-compiles without a GPU.  Flags that matter for parity:
  -fmad=false            no implicit FMA contraction in device code
  (nvcc defaults)        -prec-div=true -prec-sqrt=true -ftz=false
  -ffp-contract=off      same for the host-side C++
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libbarnacle_b200.so")
OBJ_DIR = os.path.join(HERE, "lib", "obj")

NVCC = os.environ.get("BN_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "g++"  # /usr/bin/g++ (the CXX the image exports lacks libgomp.spec / is not needed here)

HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-mavx2", "-mfma", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra",
              "-I/usr/local/cuda/include"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "-ccbin", HOST_CXX]

HOST_SOURCES = ["host/error.cpp", "host/scene_host.cpp", "cuda/scene_convert.cpp"]
CUDA_SOURCES = ["cuda/kernels.cu", "cuda/mlt.cu", "cuda/bvh_build.cu", "cuda/multi.cu"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _all_inputs() -> list[str]:
    out = []
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for dp, _, fs in os.walk(root):
            out += [os.path.join(dp, f) for f in fs if f.endswith((".h", ".hpp", ".cuh", ".cu", ".cpp"))]
    out.append(os.path.abspath(__file__))
    return out


def build(force: bool = False, verbose: bool = False, stats: bool = False, defines: list[str] | None = None, out: str | None = None) -> str:
    """stats=True: debug build with traversal phase counters (-DBN_TRAV_STATS); never shipped.
    defines/out: experiment variants (tools/).  A variant (stats, defines or out) NEVER touches the shipped library, its
    objects or the CLI: it gets its own output name (default lib_stats.so / lib_variant.so, selected at run time with
    BN_LIB=...) and its own object directory."""
    defines = defines or []
    variant = stats or bool(defines) or bool(out)
    if not variant:
        return _build(force, verbose, False, [], LIB, OBJ_DIR, True)
    name = out or ("lib_stats.so" if stats and not defines else "lib_variant.so")
    if os.path.abspath(os.path.join(LIB_DIR, name)) == os.path.abspath(LIB):
        raise ValueError("a variant build must not be written over the shipped library")
    return _build(True, verbose, stats, defines, os.path.join(LIB_DIR, name), os.path.join(OBJ_DIR, os.path.splitext(name)[0]), False)


def _build(force: bool, verbose: bool, stats: bool, defines: list[str], LIB: str, OBJ_DIR: str, with_cli: bool) -> str:
    if not force and _newer(LIB, _all_inputs()):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs, cmds = [], []
    dflags = ["-D" + d for d in defines]   # host and device code share switches such as BN_NET9_FMA
    for src in HOST_SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace("/", "_") + ".o")
        cmds.append([HOST_CXX, *HOST_FLAGS, *dflags, "-c", os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    for src in CUDA_SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace("/", "_") + ".o")
        cmds.append([NVCC, *NVCC_FLAGS, *(["-DBN_TRAV_STATS"] if stats else []), *dflags, *(["-Xptxas", "-v"] if verbose else []), "-c", os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    if verbose:
        for cmd in cmds:
            _run(cmd, verbose)
    else:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda c: _run(c, False), cmds))
    cmd = [NVCC, "-shared", "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-cudart", "static"]
    _run(cmd, verbose)
    if not with_cli:
        return LIB
    # C++ CLI mirroring Program.fs (links the C ABI only)
    cli = os.path.join(LIB_DIR, "barnacle_gpu")
    _run([HOST_CXX, "-std=c++17", "-O2", os.path.join(CSRC, "cli", "barnacle_gpu.cpp"), "-o", cli, "-L" + LIB_DIR, "-lbarnacle_b200",
          "-Wl,-rpath,$ORIGIN"], verbose)
    return LIB


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        sys.stdout.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}")


if __name__ == "__main__":
    _defs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--define=")]
    _out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv, stats="--stats" in sys.argv, defines=_defs, out=_out))
