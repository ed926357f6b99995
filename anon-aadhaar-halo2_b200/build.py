"""Build libb200zk.so (the C-ABI CUDA library) in-tree for sm_100a.

    python anon-aadhaar-halo2_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Objects go to anon-aadhaar-halo2_b200/build/, the
shared library to anon-aadhaar-halo2_b200/lib/libb200zk.so (git-ignored, shipped to the
GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "lib" / "libb200zk.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_variant(name: str, defines) -> Path:
    """An experimental build with extra -D flags into lib/libb200zk_<name>.so (objects in build/<name>/);
    pick it at run time with B200ZK_LIB_PATH (b200zk/_lib.py).  Used for A/B measurements only."""
    obj_dir = OBJ / name
    obj_dir.mkdir(parents=True, exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    out = LIB.parent / f"libb200zk_{name}.so"
    flags = FLAGS + [f"-D{d}" for d in defines]
    sources = sorted(CSRC.glob("*.cu"))

    def one(src):
        r = subprocess.run([NVCC, *flags, "-c", str(src), "-o", str(obj_dir / (src.stem + ".o"))], capture_output=True, text=True)
        (obj_dir / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(one, sources))
    r = subprocess.run([NVCC, "-shared", "-o", str(out), *[str(obj_dir / (x.stem + ".o")) for x in sources], "-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


WITNESS_LIB = PKG / "lib" / "libb200zk_witness.so"


def build_witness(force: bool = False) -> Path:
    """The host-only witness generator (host/witness_rsa.cpp, include/b200zk_witness.h): g++, no CUDA."""
    src = PKG / "host" / "witness_rsa.cpp"
    hdr = PKG.parent / "include" / "b200zk_witness.h"
    WITNESS_LIB.parent.mkdir(exist_ok=True)
    if force or _stale(WITNESS_LIB, [src, hdr]):
        r = subprocess.run([os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", str(src),
                            "-o", str(WITNESS_LIB)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return WITNESS_LIB


def build(force: bool = False, verbose: bool = False) -> Path:
    build_witness(force)
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "b200zk.h"]
    sources = sorted(CSRC.glob("*.cu"))
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return src.name

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f"compiled {name}", file=sys.stderr)
    objs = [OBJ / (s.stem + ".o") for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:       # python build.py --variant NAME -DX=1 -DY=2
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
