"""Builds ``libgtb200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gnn_tracking_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libgtb200.so"
SOURCES = ["api.cu", "plan.cu", "mlp_ffma.cu", "mlp_tc.cu", "edge_ws.cu", "node_ws.cu", "enc_ws.cu", "head_ws.cu", "losses.cu", "oc.cu", "grad.cu", "atb_tc.cu", "radius.cu", "dbscan.cu", "cell_list.cu"]
HEADERS = [HERE / "common.cuh", HERE.parents[1] / "include" / "gtb200.h"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libgtb200.so cannot be built")
    return nvcc


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        extra = HEADERS + [HERE / h for h in os.listdir(HERE) if h.endswith(".cuh")]
        if force or _stale(obj, [HERE / src, *extra]):
            cmd = [nvcc, *ARCH, *FLAGS, "-c", str(HERE / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    with ThreadPoolExecutor(max_workers=min(8, len(jobs) or 1)) as ex:
        list(ex.map(run, jobs))
    objs = [str(objdir / (s + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, *ARCH, "-shared", "-o", str(LIB), *objs, "-lcudart"])
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
