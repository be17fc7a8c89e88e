"""Build libb200mtm.so (the C-ABI library of the mtm path) for sm_100a, in-tree.

nvcc cross-compiles without a GPU.  Objects are cached under build/ (git-ignored) and rebuilt
when a source or header is newer; the shared library lands next to this file so that it travels
to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libb200mtm.so"
SOURCES = ["mtm_api.cu", "mtm_simt_f32.cu", "mtm_simt_f64.cu", "mtm_dmma_f64.cu", "mtm_tf32.cu", "mtm_ffma_tma.cu", "mtv.cu", "trans.cu", "mtm_dmma_tma.cu", "replicate.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCCFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
             "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libb200mtm.so cannot be built (and there is no CPU fallback)")


def _host_compiler_flags() -> list[str]:
    # $CC/$CXX in this image point at a gcc without libgomp specs; /usr/bin/g++ is complete.
    return ["-ccbin", "/usr/bin/g++"] if Path("/usr/bin/g++").exists() else []


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [HERE.parent / "include" / "b200_mtm.h", HERE.parent / "include" / "b200_mtv.h", HERE.parent / "include" / "b200_trans.h", HERE.parent / "include" / "b200_replicate.h",
                                                                      Path(__file__)]

    def compile_one(src: str) -> Path:
        s = CSRC / src
        o = OBJ / (s.stem + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *NVCCFLAGS, *_host_compiler_flags(), "-c", str(s), "-o", str(o)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            (OBJ / (s.stem + ".ptxas.log")).write_text(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
            if verbose:
                print(f"compiled {src}")
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, *_host_compiler_flags(), "-shared", "-cudart", "static", "-o", str(LIB),
               *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
        if verbose:
            print(f"linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
