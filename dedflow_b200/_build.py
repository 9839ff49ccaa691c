"""In-tree build of libdedflow_b200.so (nvcc, sm_100a only).

``python -m dedflow_b200._build`` or ``dedflow_b200._build.build()``.  The shared object is written next to the
sources (dedflow_b200/libdedflow_b200.so): it is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libdedflow_b200.so"
SOURCES = ["setup.cu", "color.cu", "assemble.cu", "solve.cu", "pc2.cu", "newton.cu", "compat.cu", "dist.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-Xptxas", "-v", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets"]


def _newer(a: Path, deps) -> bool:
    if not a.exists():
        return False
    t = a.stat().st_mtime
    return all(t >= d.stat().st_mtime for d in deps)


def build_variant(name: str, defines) -> Path:
    """A/B build: the same sources with extra -D flags into dedflow_b200/_obj/<name>/lib<name>.so (load it with DFB_LIB=...).
    Example: build_variant("u4", ["-DUPDATE_REVERSE=0"]) for a forward-sweeping update kernel."""
    out = OBJ / name
    out.mkdir(parents=True, exist_ok=True)
    objs = []
    for s in SOURCES:
        o = out / (Path(s).stem + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, *defines, "-I", str(PKG.parent / "include"), "-c", str(CSRC / s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s} ({name}):\n{r.stderr}")
        objs.append(o)
    lib = out / f"lib{name}.so"
    cmd = [NVCC, *ARCH, "-shared", "-o", str(lib), *map(str, objs), "-Xlinker", "-Bsymbolic",
           "-Xlinker", "--exclude-libs,ALL", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed ({name}):\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    jobs = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or not _newer(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *ARCH, *FLAGS, "-I", str(PKG.parent / "include"), "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (s.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if force or jobs or not LIB.exists():
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-Xlinker", "-Bsymbolic",
               "-Xlinker", "--exclude-libs,ALL", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python -m dedflow_b200._build --variant NAME -DFLAG ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(p)
