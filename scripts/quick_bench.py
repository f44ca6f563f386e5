"""Developer probe: kernel time of the flatquad solve + FMA peak (not the driver-facing bench.py)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from common import flatquad_terminal, c_problem
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu

def time_fn(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

L = _capi.lib()
scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
for dtype, name in ((_capi.F32, "f32"), (_capi.F64, "f64")):
    for variant in (0, 1, 2):
        fl = C.c_double()
        def run():
            _capi.check(L.pmp_fma_peak_cuda(dtype, variant, 4096, scratch.data_ptr(), C.byref(fl), None))
        best, med = time_fn(run)
        print(f"fma peak {name} variant {variant}: {fl.value/best/1e9:.1f} TFLOP/s best, {fl.value/med/1e9:.1f} median ({best:.3f} ms)")

prob, _ = c_problem("flatquad")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
y64 = flatquad_terminal(n)
for dtype, tdt, name in ((_capi.F32, torch.float32, "f32"), (_capi.F64, torch.float64, "f64")):
    y = torch.as_tensor(y64, dtype=tdt, device="cuda").contiguous()
    for method, dense, flags in (("tsit5", 0, 0), ("tsit5", 0, 1), ("tsit5", 1, 0), ("dopri5", 0, 0), ("rk4", 0, 0)):
        if method == "rk4":
            opts = _capi.make_opts(method=method, dtype=dtype, adaptive=0, dt0=-1/32, max_steps=128, save_dense=dense, flags=flags)
        else:
            opts = _capi.make_opts(method=method, dtype=dtype, save_dense=dense, flags=flags)
        out = {}
        run = lambda: pu.solve_batch_soa(prob, opts, y, out)
        best, med = time_fn(run, reps=3, warm=1)
        steps = int(out["n_steps"].sum().item())
        print(f"{name} {method} dense={dense} literal_ustar={flags}: {best:.3f} ms best {med:.3f} med | {n/best/1e3:.2f} Mtraj/s | {steps/best/1e6:.2f} Gattempt/s | mean steps {steps/n:.2f}")
        del out
