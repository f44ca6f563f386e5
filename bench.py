#!/usr/bin/env python
"""bench.py -- PMP backward-shooting rollouts: trajectories/s on N B200s, fraction of the FP32 FMA peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload endpoints|dense]
                    [--dtype f32|f64] [--method tsit5|dopri5|rk4] [--log2n 20] [--impl reference]

A "step" is ONE pass of the hot path over one batch of synthetic terminal samples: per GPU,
2^20 flat-quadrotor trajectories (flatquad_landing_experiment.py:340-437 constants, terminal states
on the LQR level set V_f = 1e-3 as in levelsets.py:551-599), Tsit5 + I-controller at
rtol = atol = 1e-5, T = 3, <= 128 step attempts, fp32 -- BASELINE.json's headline configuration
("2^20 flat-quadrotor PMP rollouts", config C4).  With N > 1 every rank solves its own 2^20 shard
(weak scaling, no collective during the solve) and the step ends with the ONE all-gather of the
endpoint records (x, lambda, v, t) the north star names.

JSON keys beyond the base contract:
  roofline      bound "fp32_fma" (the path is not a dense contraction and, endpoints-only, moves
                ~170 B per trajectory): achieved = ALGORITHMIC flops (approximate_optimal_control_b200/
                flops.py, counted on the oracle's op-counting scalar) / CUDA-event time of the solve;
                peak = a register-resident FFMA chain measured in this same process
                (MEASURED_PEAKS.json holds only HBM and bf16 numbers).  --workload dense reports the
                HBM write roofline instead (peak = MEASURED_PEAKS.json hbm_gbs).
  cpu_baseline  the CPU oracle (kind "port": a C++ restatement of the reference; jax/diffrax cannot
                be installed here) on all host threads, on a bounded sample of the same workload.
  e2e           the same metric through the public API (define_backward_solver -> solve_backward)
                from pinned HOST buffers, host->device and device->host copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# --------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="endpoints", choices=["endpoints", "dense"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--method", default="tsit5", choices=["tsit5", "dopri5", "rk4"])
    ap.add_argument("--log2n", type=int, default=20, help="log2 of trajectories per GPU")
    ap.add_argument("--cpu-log2n", type=int, default=20, help="log2 of the cpu_baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall time spent on the cpu_baseline sample")
    ap.add_argument("--gather-chunks", type=int, default=0,
                    help="N>1: pieces of the shard; the all-gather of piece k overlaps the solve of piece k+1")
    ap.add_argument("--no-step-overlap", action="store_true",
                    help="N>1: wait for the endpoint gather of step s before integrating step s+1")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1 endpoint gather transport: peer-to-peer copies over NVLink (symmetric memory) or NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def terminal_batch(n: int, seed: int, dtype):
    """Synthetic terminal samples, [14, n] struct-of-arrays (x, t, v, vx)."""
    from approximate_optimal_control_b200 import pontryagin_utils as pu, systems, terminal
    pp = systems.flatquad_problem_params()
    _, P = pu.get_terminal_lqr(pp)
    st = terminal.sample_terminal_states(P, pp["V_f"], n, seed=seed, dtype=np.float64)
    y = np.concatenate([st["x"].T, st["t"][None], st["v"][None], st["vx"].T], 0)
    return np.ascontiguousarray(y, dtype=dtype), {k: v.astype(dtype) for k, v in st.items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t_begin - 0.3 <= t <= t_end + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    """Host threads the CPU arms use: every core this process may be scheduled on (not OMP_NUM_THREADS, which
    torch.distributed.run sets to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself (jax 0.4.23 +
    diffrax 0.4.1) is not installable offline, so this is the CPU oracle ("port"), all host threads,
    each step a bounded sample (2^17 trajectories) of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    n = 1 << 17
    odt = O.F32 if args.dtype == "f32" else O.F64
    y, _ = terminal_batch(n, 1, np.float32 if args.dtype == "f32" else np.float64)
    method = {"tsit5": O.TSIT5, "dopri5": O.DOPRI5, "rk4": O.RK4}[args.method]
    opts = O.make_opts(method=method, dtype=odt, adaptive=int(args.method != "rk4"),
                       dt0=-0.1 if args.method != "rk4" else -1 / 32, save_dense=int(args.workload == "dense"))
    prob = O.flatquad_problem()
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on explicitly
    for _ in range(args.warmup):
        O.solve_batch(prob, opts, y, n_threads=cores)
    t0 = time.perf_counter()
    attempts = 0
    for _ in range(args.steps):
        out = O.solve_batch(prob, opts, y, n_threads=cores)
        attempts += int(out["n_steps"].sum())
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "PMP trajectories/sec", "value": value, "unit": "trajectories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "rk_step_attempts_per_s": attempts / dt,
        "config": workload_config(args, n),
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x 2^17 flatquad terminal samples, same solver settings "
                                   "(CPU restatement of the reference; jax/diffrax unavailable offline)"},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_per_gpu):
    return {
        "workload": f"flatquad PMP backward shooting, 2^{int(np.log2(n_per_gpu))} terminal samples per GPU, "
                    f"{args.method} {'adaptive rtol=atol=1e-5' if args.method != 'rk4' else 'fixed dt=1/32'}, "
                    f"T=3, max 128 attempts, {'dense SaveAt(steps) output' if args.workload == 'dense' else 'endpoints only'}"
                    + (", all-gather of endpoints" if args.gpus > 1 else ""),
        "system": "flatquad (nx=6, nu=2, box-clipped u*)", "trajectories_per_gpu": n_per_gpu,
        "solver": args.method, "rtol": 1e-5, "atol": 1e-5, "T": 3.0, "max_steps": 128,
        "parallelism": f"shard{args.gpus}" if args.gpus > 1 else "single",
        "l2": ("flushed between timed steps (256 MiB write); per-step working set 172 MB > 126 MB L2" if args.gpus == 1
               else "no flush: back-to-back pipelined steps, per-step working set 172 MB per GPU > 126 MB L2"),
    }


# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from approximate_optimal_control_b200 import _capi, flops, pontryagin_utils as pu, sharding, systems

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the PMP path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    n = 1 << args.log2n
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    ndt = np.float32 if args.dtype == "f32" else np.float64
    y_np, st_np = terminal_batch(n, 1 + rank, ndt)
    pp, ap = systems.flatquad_problem_params(), systems.flatquad_algo_params()
    problem = systems.to_c_problem(pp)
    method = _capi.METHODS[args.method]
    dense = args.workload == "dense"
    opts = _capi.make_opts(method=method, dtype=_capi.F32 if args.dtype == "f32" else _capi.F64,
                           adaptive=int(args.method != "rk4"), dt0=-0.1 if args.method != "rk4" else -1 / 32,
                           save_dense=int(dense))
    L = _capi.lib()

    y_dev = torch.as_tensor(y_np, device=dev).contiguous()
    out = {}
    # N > 1: the shard is integrated in pieces and the endpoint all-gather of piece k overlaps piece k+1
    # (0 = auto: every extra piece costs one more kernel tail, ~50 us; worth it only when the gather is long)
    chunks = args.gather_chunks if args.gather_chunks > 0 else 1
    pg = sharding.PipelinedGather(problem, opts, n, tdt, dev, chunks=chunks, transport=args.gather) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        if world > 1:
            # software pipeline across steps: the gather of step s runs while step s+1 is integrated;
            # every gather completes inside the timed region (pg.wait() before the closing barrier)
            pg.run(y_dev, wait=args.no_step_overlap)
        else:
            pu.solve_batch_soa(problem, opts, y_dev, out)

    def barrier():
        if world > 1:
            pg.wait()
            torch.cuda.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FMA peak (roofline denominator), measured in this process ---------------------------------
    scratch = torch.zeros(16, dtype=torch.float64, device=dev)
    pdt = _capi.F32 if args.dtype == "f32" else _capi.F64
    peak_tflops = 0.0
    for variant in (0, 1):
        fl = C.c_double()
        for it in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _capi.check(L.pmp_fma_peak_cuda(pdt, variant, 4096, scratch.data_ptr(), C.byref(fl),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.record()
            torch.cuda.synchronize()
            if it >= 1:
                peak_tflops = max(peak_tflops, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)

    # ---- warm-up, then K timed steps ----------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    attempts = 0
    t_begin = time.time()
    barrier()
    drained = torch.cuda.Event(enable_timing=True)
    for k in range(args.steps):
        if world == 1:
            flush.fill_(k & 0xFF)  # L2 flush, outside the per-step event pair
        a, m, b = ev[k]
        a.record()
        step()
        m.record()
        b.record()
    if world > 1:
        pg.wait()  # the current stream waits for the last gather: the pipeline is drained
    drained.record()
    barrier()
    t_end = time.time()
    step_ms = [a.elapsed_time(b) for (a, m, b) in ev]
    solve_ms = [a.elapsed_time(m) for (a, m, b) in ev]
    if world > 1:
        # steps are software-pipelined (gather of step s under the integration of step s+1) and run back
        # to back without L2 flushes (the 172 MB per-step working set exceeds the 126 MB L2): the time of
        # the K steps is first start .. pipeline drained
        step_ms = [ev[0][0].elapsed_time(drained) / args.steps] * args.steps
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    total_solve_ms = torch.tensor([sum(solve_ms)], dtype=torch.float64, device=dev)
    my_attempts_t = (pg.attempts() if world > 1 else out["n_steps"].sum()).to(torch.float64).reshape(1)
    attempts_t = my_attempts_t.clone()
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_solve_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(attempts_t, op=dist.ReduceOp.SUM)
    total_s = total_ms.item() * 1e-3
    attempts_per_step = attempts_t.item()  # all ranks, one step
    value = n * world * args.steps / total_s

    # ---- roofline of the dominant kernel (rank 0's own launches) -------------------------------------
    my_attempts = int(my_attempts_t.item())
    if world > 1:
        # kernel-only time of this rank's solve launches: re-time the same pieces without the collective
        t_solo = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for y_in, o_ in zip(pg.inputs, pg.outs[0]):
                pu.solve_batch_soa(problem, opts, y_in, o_)
            b.record()
            torch.cuda.synchronize()
            t_solo.append(a.elapsed_time(b))
        solve_s = min(t_solo) * 1e-3
    else:
        solve_s = float(np.mean(solve_ms)) * 1e-3
    if dense:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = peaks.get("hbm_gbs", 6650.0)
        bytes_alg = flops.dense_bytes(n, 6, opts.max_steps, 4 if args.dtype == "f32" else 8) + 2 * y_dev.numel() * y_dev.element_size()
        roof = {"bound": "hbm", "achieved": bytes_alg / solve_s / 1e9, "peak": peak_gbs, "unit": "GB/s",
                "frac": bytes_alg / solve_s / 1e9 / peak_gbs, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": bytes_alg}
    else:
        fl_alg = flops.solve_flops(problem.system_id, method, my_attempts, n)
        roof = {"bound": "fp32_fma" if args.dtype == "f32" else "fp64_fma", "achieved": fl_alg / solve_s / 1e12,
                "peak": peak_tflops, "unit": "TFLOP/s", "frac": fl_alg / solve_s / 1e12 / peak_tflops, "traffic": None,
                "peak_source": "FMA-chain microbenchmark (pmp_fma_peak_cuda) in this process, burst; "
                               "MEASURED_PEAKS.json has no CUDA-core figure",
                "algorithmic_flops_per_launch": fl_alg,
                "flops_per_attempt": flops.ATTEMPT_FLOPS[(problem.system_id, method)],
                "flops_per_rhs": flops.RHS_FLOPS[problem.system_id],
                "kernel_ms": solve_s * 1e3}
    # measured DRAM traffic of one launch (ncu --set full capture of this same configuration, committed under
    # profiles/); null when no capture exists for the configuration
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file) and args.log2n == 20:
        try:
            roof["traffic"] = json.load(open(traffic_file)).get(f"{args.workload}_{args.dtype}_{args.method}")
        except Exception:
            pass

    # ---- e2e: public API, host buffers ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        solve_backward, _ = pu.define_backward_solver(pp, dict(ap, pontryagin_solver_method=args.method,
                                                               pontryagin_solver_save_steps=dense,
                                                               **({"pontryagin_solver_dt0": 1 / 32} if args.method == "rk4" else {})))
        host_in = {k: torch.as_tensor(v).pin_memory() for k, v in st_np.items()}
        host_out = {k: torch.empty(v.shape, dtype=tdt).pin_memory() for k, v in st_np.items()}
        h2d = sum(v.numel() * v.element_size() for v in host_in.values())
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())
        if not dense:
            d2h += n * 4  # num_accepted_steps, num_steps, result come back with the endpoints, one packed word each

        def e2e_step():
            if dense:  # dense solutions stay on the device (8 GB per step); only the endpoints travel back
                dev_in = {k: v.to(dev, non_blocking=True) for k, v in host_in.items()}
                ye = solve_backward(dev_in).y_end
                for k in host_out:
                    host_out[k].copy_(ye[k], non_blocking=True)
                torch.cuda.synchronize()
                return
            # host in -> host out: the public call with pinned host tensors goes to pmp_solve_host (csrc/pmp_host.cu),
            # which cuts the batch into pieces and overlaps the copies of neighbouring pieces with the integration; the
            # endpoints and per-trajectory counters come back in pinned host memory
            sol = solve_backward(host_in)
            assert not sol.y_end["x"].is_cuda and sol.y_end["x"].shape == host_in["x"].shape

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * args.steps / e_s.item(), "unit": "trajectories/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "pontryagin_utils.define_backward_solver(...)[0](state dict of pinned host tensors) -> endpoints in pinned "
                      "host tensors through the C entry pmp_solve_host (host pointers), which pipelines host->device copy, integration "
                      "and device->host copy over pieces on its own streams"
                      if not dense else "define_backward_solver(...)[0](device state dict) after an explicit host->device copy; "
                      "endpoints copied back to pinned host memory, dense nodes stay on the device"}

    # clocks: the K-step region lasts tens of ms, shorter than nvidia-smi's sampling period, so the
    # sampler runs from the first warm-up step to the end of the e2e loop (GPU busy throughout)
    clocks = sampler.stop(t_begin, time.time()) if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps + e2e loop"

    # ---- cpu baseline (rank 0, N=1 only) --------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        nc = 1 << args.cpu_log2n
        oo = O.make_opts(method={"tsit5": O.TSIT5, "dopri5": O.DOPRI5, "rk4": O.RK4}[args.method],
                         dtype=O.F32 if args.dtype == "f32" else O.F64, adaptive=int(args.method != "rk4"),
                         dt0=-0.1 if args.method != "rk4" else -1 / 32)
        ys = y_np[:, :nc] if nc <= n else y_np
        ncores = host_cores()
        O.solve_batch(O.flatquad_problem(), oo, ys[:, :4096], n_threads=ncores)
        t0 = time.perf_counter()
        passes = 0
        while True:  # bounded sample: whole passes over the batch until ~10 s of wall time on all host threads
            O.solve_batch(O.flatquad_problem(), oo, ys, n_threads=ncores)
            passes += 1
            dtc = time.perf_counter() - t0
            if dtc >= args.cpu_seconds or passes >= 64:
                break
        cpu = {"value": passes * ys.shape[1] / dtc, "unit": "trajectories/s", "cores": ncores, "kind": "port",
               "sample": f"{passes} passes over the first 2^{int(np.log2(ys.shape[1]))} trajectories of the same batch, same "
                         f"solver settings, {dtc:.1f} s wall on {ncores} threads; C++ restatement of the reference "
                         "(jax/diffrax not installable offline)"}

    if rank == 0:
        line = {
            "metric": "PMP trajectories/sec", "value": value, "unit": "trajectories/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms.item() / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, n),
            "rk_step_attempts_per_s": attempts_per_step * args.steps / total_s,
            "mean_attempts_per_trajectory": attempts_per_step / (n * world),
            "solve_ms_per_step": (total_solve_ms.item() / args.steps) if world == 1 else solve_s * 1e3,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": args.steps * L.pmp_solve_launch_count(C.byref(problem), C.byref(opts), n) *
                            (len(pg.slices) if world > 1 else 1),
            "gather": ({"chunks": len(pg.slices), "transport": pg.transport, "bytes_per_rank_per_step": int(y_dev.numel() * y_dev.element_size()),
                        "overlap": ("none across steps" if args.no_step_overlap else
                                    "gather of step s overlaps the integration of step s+1 (double-buffered); "
                                    "all gathers complete inside the timed region")}
                       if world > 1 else None),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
