"""Builds libb200pmp.so (hand-written sm_100a CUDA + the C ABI of include/b200pmp.h) in-tree.

    python -m approximate_optimal_control_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so that it travels
with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libb200pmp.so")
_OBJ_DIR = os.path.join(_PKG, "build")

SOURCES = ["pmp_capi.cu", "pmp_host.cu", "pmp_eval.cu", "pmp_select.cu", "pmp_gather.cu", "pmp_sens.cu", "pmp_interp.cu", "pmp_forward.cu", "pmp_forward_nn.cu", "pmp_forward_cluster.cu", "pmp_forward_tc.cu", "pmp_forward_umma.cu", "pmp_sys_flatquad.cu", "pmp_sys_flatquad_vxx.cu", "pmp_sys_orbits.cu", "pmp_sys_lqr.cu"]
HEADERS = ["pmp_hostq.h", "pmp_common.cuh", "pmp_systems.cuh", "pmp_tableau.cuh", "pmp_solve.cuh", "pmp_interp.cuh", "pmp_forward.cuh", "pmp_forward_cluster.cuh", "pmp_forward_tc.cuh", "pmp_forward_umma.cuh", "pmp_umma.cuh", "pmp_ustar2d.cuh", "pmp_sens.cuh", "pmp_dual.cuh", "pmp_dual_math.cuh", "pmp_vxx_dual.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
    # IEEE division / sqrt, no flush-to-zero, FMA contraction on (the default): the state path
    # must not silently lose precision; approximate ops are requested explicitly where wanted.
    "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
]


# developer knob: extra nvcc flags (e.g. -DPMP_MINB32=4) for tuning experiments
NVCC_FLAGS += os.environ.get("B200PMP_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libb200pmp.so cannot be built")
    return exe


def _deps():
    deps = [os.path.join(_CSRC, f) for f in SOURCES + HEADERS]
    return deps + [os.path.join(_ROOT, "include", "b200pmp.h"), os.path.abspath(__file__)]


def _source_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, C header, this file with its flags)."""
    import hashlib
    h = hashlib.sha256()
    for d in _deps():
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


_HASH_PATH = LIB_PATH + ".srchash"


def _stale() -> bool:
    """True when the library is missing or was built from other sources.  Compared by CONTENT (the hash build() leaves next
    to the library), not by mtime: the library travels to GPU boxes in snapshots that do not keep file times."""
    if not os.path.exists(LIB_PATH):
        return True
    try:
        return open(_HASH_PATH).read().strip() != _source_hash()
    except OSError:
        return os.path.getmtime(LIB_PATH) < max(os.path.getmtime(d) for d in _deps())


LAST_BUILD = {"mode": None, "units": []}  # what the last build() call did (printed by __graft_entry__.build)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        LAST_BUILD.update(mode="up to date (the content hash of the sources matches the prebuilt library)", units=[])
        return LIB_PATH
    os.makedirs(_OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc wrapper without its spec files; use the system one
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

    def compile_one(src):
        obj = os.path.join(_OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-ccbin", host, "-c", os.path.join(_CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-ccbin", host, "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(_HASH_PATH, "w") as fh:
        fh.write(_source_hash())
    LAST_BUILD.update(mode="compiled with " + " ".join([os.path.basename(nvcc), *NVCC_FLAGS]), units=list(SOURCES))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
