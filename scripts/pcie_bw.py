"""PCIe copy bandwidth of this box from pinned host memory: host->device alone, device->host alone, both directions at
once on two streams.  The host in -> host out path (pmp_solve_host) is bounded by the last figure."""
import json
import time

import torch

mb = 64
h_in = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return (h2d + d2h) * (mb << 20) / dt / 1e9


for _ in range(2):
    run(True, True, 3)
print(json.dumps({"buffer_MiB": mb, "h2d_GBps": round(run(True, False), 1), "d2h_GBps": round(run(False, True), 1),
                  "both_total_GBps": round(run(True, True), 1)}))
