#!/usr/bin/env python
"""Host-link probe (one GPU; run under torchrun for all ranks at once): what the copy engines and the SMs ("zero copy")
move between pinned host memory and HBM, one direction and both, as a function of the copy size.
usage: python scripts/pcie_probe.py [--bind] [--short] [> profiles/rNN_pcie_probe.jsonl]
       torchrun --nproc-per-node 8 scripts/pcie_probe.py --short [--bind]   (every rank's line; the aggregate is their sum)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from approximate_optimal_control_b200 import _capi

from approximate_optimal_control_b200 import sharding

rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
BIND, SHORT = "--bind" in sys.argv, "--short" in sys.argv  # --bind: process onto the GPU's NUMA node BEFORE the pinned allocations
binding = sharding.bind_host_to_gpu(rank) if BIND else {"bound": False, "why": "not asked"}
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
L = _capi.lib()
TOTAL = 64 << 20
h_in = torch.empty(TOTAL, dtype=torch.uint8).pin_memory()
h_out = torch.empty(TOTAL, dtype=torch.uint8).pin_memory()
d_a = torch.empty(TOTAL, dtype=torch.uint8, device="cuda")
d_b = torch.empty(TOTAL, dtype=torch.uint8, device="cuda")
s1, s2, s3, s4 = (torch.cuda.Stream() for _ in range(4))


def ce(dst, src, stream, piece):
    with torch.cuda.stream(stream):
        for o in range(0, TOTAL, piece):
            dst[o:o + piece].copy_(src[o:o + piece], non_blocking=True)


def sm(dst, src, stream, piece, blocks):
    for o in range(0, TOTAL, piece):
        _capi.check(L.pmp_host_probe_copy(dst.data_ptr() + o, src.data_ptr() + o, piece, blocks, C.c_void_p(stream.cuda_stream)))


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps + 1):
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for s in (s1, s2, s3, s4):
            s.wait_event(a)
        fn()
        for s in (s1, s2, s3, s4):
            b.wait(torch.cuda.current_stream()) if False else torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def line(name, fn, nbytes, **kw):
    ms = timed(fn)
    d = dict(test=name, rank=rank, world=world, ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1), **kw)
    print(json.dumps(d), flush=True)


print(json.dumps(dict(test="host_binding", rank=rank, world=world, **binding)), flush=True)
if SHORT:  # all ranks at once: what the box's host links give every GPU at the same time
    for piece in (1 << 20, 8 << 20):
        line("ce_h2d", lambda: ce(d_a, h_in, s1, piece), TOTAL, piece_KiB=piece >> 10, bind=BIND)
        line("ce_d2h", lambda: ce(h_out, d_b, s2, piece), TOTAL, piece_KiB=piece >> 10, bind=BIND)
        line("ce_both", lambda: (ce(d_a, h_in, s1, piece), ce(h_out, d_b, s2, piece)), 2 * TOTAL, piece_KiB=piece >> 10, bind=BIND)
    sys.exit(0)
for piece in (512 << 10, 1 << 20, 3 << 20, 8 << 20, 64 << 20):
    line("ce_h2d", lambda: ce(d_a, h_in, s1, piece), TOTAL, piece_KiB=piece >> 10)
    line("ce_d2h", lambda: ce(h_out, d_b, s2, piece), TOTAL, piece_KiB=piece >> 10)
    line("ce_both", lambda: (ce(d_a, h_in, s1, piece), ce(h_out, d_b, s2, piece)), 2 * TOTAL, piece_KiB=piece >> 10)
half = TOTAL // 2
for piece in (1 << 20, 3 << 20):
    # two copy-engine streams per direction
    def two_h2d():
        with torch.cuda.stream(s1):
            for o in range(0, half, piece):
                d_a[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)
        with torch.cuda.stream(s3):
            for o in range(half, TOTAL, piece):
                d_a[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)

    def two_d2h():
        with torch.cuda.stream(s2):
            for o in range(0, half, piece):
                h_out[o:o + piece].copy_(d_b[o:o + piece], non_blocking=True)
        with torch.cuda.stream(s4):
            for o in range(half, TOTAL, piece):
                h_out[o:o + piece].copy_(d_b[o:o + piece], non_blocking=True)
    line("ce_h2d_2streams", two_h2d, TOTAL, piece_KiB=piece >> 10)
    line("ce_both_2x2streams", lambda: (two_h2d(), two_d2h()), 2 * TOTAL, piece_KiB=piece >> 10)
for blocks in (8, 32, 148, 592):
    line("sm_h2d (kernel reads pinned host)", lambda: sm(d_a, h_in, s1, TOTAL, blocks), TOTAL, blocks=blocks)
    line("sm_d2h (kernel writes pinned host)", lambda: sm(h_out, d_b, s2, TOTAL, blocks), TOTAL, blocks=blocks)
line("sm_both", lambda: (sm(d_a, h_in, s1, TOTAL, 64), sm(h_out, d_b, s2, TOTAL, 64)), 2 * TOTAL, blocks=64)
line("ce_h2d + sm_d2h", lambda: (ce(d_a, h_in, s1, 8 << 20), sm(h_out, d_b, s2, TOTAL, 64)), 2 * TOTAL, blocks=64)
line("sm_h2d + ce_d2h", lambda: (sm(d_a, h_in, s1, TOTAL, 64), ce(h_out, d_b, s2, 8 << 20)), 2 * TOTAL, blocks=64)
for piece in (512 << 10, 1 << 20):
    line("sm_d2h small launches", lambda: sm(h_out, d_b, s2, piece, 32), TOTAL, piece_KiB=piece >> 10)
