#!/usr/bin/env python
"""Tuning variants of the integrator: recompiles ONE translation unit with extra -D flags and links it with the objects of
the regular build into approximate_optimal_control_b200/build/variants/lib_<tag>.so (select with B200PMP_LIB=<path>).
usage: python scripts/build_variant.py <tag> <unit.cu> [-DPMP_MINB32=4 ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approximate_optimal_control_b200 import build as B

tag, unit, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
out_dir = os.path.join(B._OBJ_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"{tag}_{unit.replace('.cu', '.o')}")
host = "/usr/bin/g++"
r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-Xptxas=-v", "-ccbin", host, "-c", os.path.join(B._CSRC, unit), "-o", obj],
                   capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
print("\n".join(l for l in r.stderr.splitlines() if "spill" in l and " 0 bytes spill stores" not in l)[:2000])
objs = [obj if s == unit else os.path.join(B._OBJ_DIR, s.replace(".cu", ".o")) for s in B.SOURCES]
lib = os.path.join(out_dir, f"lib_{tag}.so")
r = subprocess.run([B._nvcc(), "-shared", "-ccbin", host, "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-ldl"],
                   capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
print(lib)
