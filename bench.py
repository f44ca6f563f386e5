#!/usr/bin/env python
"""bench.py -- PMP backward-shooting rollouts: trajectories/s on N B200s, fraction of the FP32 FMA peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload endpoints|dense] [--scaling weak|strong]
                    [--dtype f32|f64] [--method tsit5|dopri5|rk4] [--log2n 20] [--impl reference]

A "step" is ONE pass of the hot path over one batch of synthetic terminal samples: per GPU,
2^20 flat-quadrotor trajectories (flatquad_landing_experiment.py:340-437 constants, terminal states
on the LQR level set V_f = 1e-3 as in levelsets.py:551-599), Tsit5 + I-controller at
rtol = atol = 1e-5, T = 3, <= 128 step attempts, fp32 -- BASELINE.json's headline configuration
("2^20 flat-quadrotor PMP rollouts", config C4).  With N > 1 every rank solves its own shard
(no collective during the solve) and the step ends with the ONE all-gather of the endpoint records
(x, lambda, v, t) the north star names.  --scaling weak (default, what the driver runs): 2^20 per GPU.
--scaling strong: BASELINE.json configs[3] as worded, 2^20 in all, 2^20 / N per GPU.  The weak line also carries
the strong measurement of the same run under "strong".

Timing (the same at every N): W >= 3 warm-up steps, barrier + synchronize, K steps issued back to back on the device
between two CUDA events ("first step starts" .. "last gather has landed"), barrier + synchronize; max over ranks.
No L2 flush: one step touches 172 MB per GPU (terminal states, k1, endpoints), more than the 126 MB L2 (the strong
mode's smaller shards say so in config.l2).  The dominant kernel is timed on its own stream by an event pair around
every launch of the same K steps.

JSON keys beyond the base contract:
  roofline      bound "fp32_fma" (the path is not a dense contraction and, endpoints-only, moves ~170 B per trajectory):
                achieved = ALGORITHMIC flops (approximate_optimal_control_b200/flops.py, counted on the oracle's
                op-counting scalar) / CUDA-event time of the solve; peak = register-resident FMA chains measured in this
                same process (MEASURED_PEAKS.json holds only HBM and bf16 numbers); frac_of_theoretical uses
                148 SM x 128 lanes x 2 x sm_max_mhz.  --workload dense (and the "dense" block of the default line)
                report the HBM write roofline instead (peak = MEASURED_PEAKS.json hbm_gbs).
  parity        accuracy of THIS run's endpoints: a 2^12 subsample against the converged fp64 solution (oracle at
                tol 1e-12), next to the fp32 oracle's accuracy against the same truth, in units of the solver tolerance.
  cpu_baseline  the CPU oracle (kind "port": a C++ restatement of the reference; jax/diffrax cannot be installed here) on
                all host threads, on a bounded sample of the same workload; plus one-core and fp64 figures.
  e2e           the same metric through the reference's own call of the path, sols = vmap(solve_backward_lqr)(xfs)
                (levelsets.py:576-601; here levelsets.define_solve_backward_lqr): pinned HOST terminal samples in, endpoints
                in pinned host memory out, host->device and device->host copies inside the timed region.
  gather_check  N > 1: rank 0 re-solves every rank's shard in one launch and compares it BIT FOR BIT with the slice
                that rank pushed into rank 0's gathered dataset.
  gather        N > 1: transport ("fused": the gather is part of the integrator, pmp_solve_batch_push_cuda -- finished chunks are
                stored into the peers' buffers over NVLink, through the NVSwitch multicast mapping when "nvswitch_multicast" is
                true; --gather p2p / nccl select the copy-engine and NCCL transports), ms per step with and without it.
  host_binding  what sharding.bind_host_to_gpu did for this rank (process onto the CPUs / memory next to its GPU; a no-op on
                boxes that expose no NUMA topology).
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 2^log2n trajectories per GPU; strong: 2^log2n in all (BASELINE.json configs[3])")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--method", default="tsit5", choices=["tsit5", "dopri5", "rk4"])
    ap.add_argument("--log2n", type=int, default=20, help="log2 of trajectories per GPU (weak) / in all (strong)")
    ap.add_argument("--cpu-log2n", type=int, default=20, help="log2 of the cpu_baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall time spent on the cpu_baseline sample")
    ap.add_argument("--gather-chunks", type=int, default=0,
                    help="N>1: pieces of the shard; the all-gather of piece k overlaps the solve of piece k+1")
    ap.add_argument("--no-step-overlap", action="store_true",
                    help="N>1: wait for the endpoint gather of step s before integrating step s+1")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl", "fused"],
                    help="N>1 endpoint gather transport: auto = fused into the integrator (stores to peer memory from the retiring "
                         "warps) where symmetric memory is available; p2p = copy-engine pushes over NVLink; nccl")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bind-host", action="store_true",
                    help="leave the process where the launcher put it (default: onto the CPUs / memory of the GPU's NUMA node)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense (C5) block of the default N=1 line")
    ap.add_argument("--no-forward-nn", action="store_true", help="skip the value-network forward-simulation block of the default N=1 line")
    ap.add_argument("--e2e-per-piece", action="store_true",
                    help="e2e through per-piece launches (state-dict solve_backward) instead of the persistent host-fed launch: the "
                         "form a serialising profiler (ncu) can run -- the persistent kernel waits for copies ncu holds back")
    ap.add_argument("--no-strong", action="store_true", help="N>1, weak: skip the strong-scaling block")
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


def oracle_opts(O, args, dtype=None, dense=False):
    method = {"tsit5": O.TSIT5, "dopri5": O.DOPRI5, "rk4": O.RK4}[args.method]
    odt = {"f32": O.F32, "f64": O.F64}[dtype or args.dtype]
    return O.make_opts(method=method, dtype=odt, adaptive=int(args.method != "rk4"),
                       dt0=-0.1 if args.method != "rk4" else -1 / 32, save_dense=int(dense))


def n_per_gpu(args, world):
    n = 1 << args.log2n
    return n if args.scaling == "weak" else max(1, n // world)


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself (jax 0.4.23 +
    diffrax 0.4.1) is not installable offline, so this is the CPU oracle ("port"), all host threads, on the SAME
    configuration as the b200 arm: every step is one pass over the same 2^log2n-trajectory batch (0.4-2 s of CPU)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    n = n_per_gpu(args, max(1, args.gpus)) if args.scaling == "weak" else (1 << args.log2n)
    y, _ = terminal_batch(n, 1, np.float32 if args.dtype == "f32" else np.float64)
    opts = oracle_opts(O, args, dense=args.workload == "dense")
    prob = O.flatquad_problem()
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on explicitly
    for _ in range(min(args.warmup, 2)):
        O.solve_batch(prob, opts, y[:, : n // 4], n_threads=cores)
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
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "rk_step_attempts_per_s": attempts / dt,
        "config": workload_config(args, n, 1),
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} passes over 2^{int(np.log2(n))} flatquad terminal samples (one GPU's batch of the "
                                   "b200 arm), same solver settings (CPU restatement of the reference; jax/diffrax unavailable offline)"},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_local, world):
    per_step_mb = 3 * 14 * n_local * (4 if args.dtype == "f32" else 8) / 1e6
    return {
        "workload": f"flatquad PMP backward shooting, 2^{np.log2(n_local):g} terminal samples per GPU, "
                    f"{args.method} {'adaptive rtol=atol=1e-5' if args.method != 'rk4' else 'fixed dt=1/32'}, "
                    f"T=3, max 128 attempts, {'dense SaveAt(steps) output' if args.workload == 'dense' else 'endpoints only'}"
                    + (", all-gather of endpoints" if world > 1 else ""),
        "system": "flatquad (nx=6, nu=2, box-clipped u*)", "trajectories_per_gpu": n_local,
        "trajectories_total": n_local * world,
        "solver": args.method, "rtol": 1e-5, "atol": 1e-5, "T": 3.0, "max_steps": 128,
        "parallelism": f"shard{world}" if world > 1 else "single",
        "l2": (f"no flush: steps run back to back at every N; per-step working set {per_step_mb:.0f} MB per GPU "
               + ("> 126 MB L2" if per_step_mb > 126 else "< 126 MB L2 (inputs and k1 of the previous step may still be cached)")),
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
    # before anything allocates pinned host memory: this rank's process onto the CPUs (and so the memory) next to its GPU
    all_cpus = os.sched_getaffinity(0)
    host_binding = {"bound": False, "why": "--no-bind-host"} if args.no_bind_host else sharding.bind_host_to_gpu(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    ndt = np.float32 if args.dtype == "f32" else np.float64
    es = 4 if args.dtype == "f32" else 8
    pp, ap = systems.flatquad_problem_params(), systems.flatquad_algo_params()
    problem = systems.to_c_problem(pp)
    method = _capi.METHODS[args.method]
    dense = args.workload == "dense"
    mk_opts = lambda dn: _capi.make_opts(method=method, dtype=_capi.F32 if args.dtype == "f32" else _capi.F64,
                                         adaptive=int(args.method != "rk4"), dt0=-0.1 if args.method != "rk4" else -1 / 32,
                                         save_dense=int(dn))
    opts = mk_opts(dense)
    L = _capi.lib()
    launches_per_solve = L.pmp_solve_launch_count(C.byref(problem), C.byref(opts), 1 << 10)
    chunks = args.gather_chunks if args.gather_chunks > 0 else 1
    gpu_launches = 0

    def shard_inputs(scaling):
        """(y_dev [14, n_local], st_np leaves of this rank's shard, n_local, seed rule)"""
        if scaling == "weak":
            n_loc = 1 << args.log2n
            y_np, st_np = terminal_batch(n_loc, 1 + rank, ndt)
        else:  # one global batch (seed 1), contiguous shards (iid samples: SURVEY.md 8(e))
            n_all = 1 << args.log2n
            n_loc = n_all // world
            y_all, st_all = terminal_batch(n_all, 1, ndt)
            sl = slice(rank * n_loc, (rank + 1) * n_loc)
            y_np, st_np = np.ascontiguousarray(y_all[:, sl]), {k: np.ascontiguousarray(v[sl]) for k, v in st_all.items()}
        return torch.as_tensor(y_np, device=dev).contiguous(), st_np, y_np, n_loc

    def barrier(pg=None):
        if pg is not None:
            pg.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    def timed_steps(y_dev, n_loc, o, with_gather, steps, warm):
        """K steps back to back between two events (first step starts .. everything, gathers included, has landed);
        per-launch event pairs around every solve give the kernel time.  Returns a dict (ms are max over ranks)."""
        nonlocal gpu_launches
        pg = (sharding.PipelinedGather(problem, o, n_loc, tdt, dev, chunks=chunks, transport=args.gather)
              if (world > 1 and with_gather) else None)
        out = {}

        def step():
            if pg is not None:
                pg.run(y_dev, wait=args.no_step_overlap)
            else:
                pu.solve_batch_soa(problem, o, y_dev, out)

        for _ in range(warm):
            step()
        barrier(pg)
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(pg)
        t0.record()
        for a, b in pairs:
            a.record()
            step()
            b.record()   # end of this step's solve launches on the current stream (the gather runs on the side stream)
        if pg is not None:
            pg.wait()    # the current stream waits for the last gather: the pipeline is drained
        t1.record()
        barrier(pg)
        gpu_launches += steps * launches_per_solve * (len(pg.slices) if pg is not None else 1)
        total_ms = reduce_max(t0.elapsed_time(t1))
        solve_ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
        att = (pg.attempts() if pg is not None else out["n_steps"].sum()).item()
        return {"total_ms": total_ms, "ms_per_step": total_ms / steps, "solve_ms": solve_ms, "attempts": att, "pg": pg,
                "out": out}

    # ---- FMA peak (roofline denominator), measured in this process ---------------------------------
    scratch = torch.zeros(16, dtype=torch.float64, device=dev)
    pdt = _capi.F32 if args.dtype == "f32" else _capi.F64
    peaks_by_variant = {}
    for variant in ((0, 1, 2) if args.dtype == "f32" else (0, 1)):
        fl = C.c_double()
        best = 0.0
        for it in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _capi.check(L.pmp_fma_peak_cuda(pdt, variant, 4096, scratch.data_ptr(), C.byref(fl),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.record()
            torch.cuda.synchronize()
            if it >= 1:
                best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
        peaks_by_variant[{0: "ffma_reg", 1: "ffma_imm", 2: "ffma2_packed"}[variant]] = best
    peak_tflops = max(peaks_by_variant.values())

    # ---- the headline measurement ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_begin = time.time()
    y_dev, st_np, y_np, n = shard_inputs(args.scaling)
    warm = max(args.warmup, 3)
    main_run = timed_steps(y_dev, n, opts, True, args.steps, warm)
    total_s = main_run["total_ms"] * 1e-3
    value = n * world * args.steps / total_s
    attempts_all = reduce_sum(main_run["attempts"])  # all ranks, one step
    pg = main_run["pg"]
    gather = None
    if world > 1:
        solo = timed_steps(y_dev, n, opts, False, args.steps, 2)  # the same K steps without the collective
        gather = {"chunks": len(pg.slices), "transport": pg.transport, "nvswitch_multicast": bool(getattr(pg, "multicast", False)),
                  "bytes_per_rank_per_step": int(y_dev.numel() * y_dev.element_size()),
                  "ms_with_gather": main_run["ms_per_step"], "ms_without_gather": solo["ms_per_step"],
                  "overlap": ("none across steps" if args.no_step_overlap else
                              "fused into the integrator (pmp_solve_batch_push_cuda): the warp that finishes a work-queue chunk of "
                              "32 trajectories stores its endpoint rows into every peer's gather buffer over NVLink while the rest "
                              "of the shard is integrated; three buffer sets; all records have landed inside the timed region"
                              if pg.transport == "fused" else
                              "gather of step s (copy engines, peer-to-peer over NVLink; NCCL all_gather_into_tensor with "
                              "--gather nccl) overlaps the integration of step s+1 (double-buffered); all gathers complete "
                              "inside the timed region")}
    solve_s = main_run["solve_ms"] * 1e-3 if world == 1 else reduce_max(main_run["solve_ms"]) * 1e-3
    my_attempts = int(main_run["attempts"])

    # ---- gather_check: the gathered dataset is what a single launch computes, bit for bit ------------------------
    gather_check = None
    if world > 1:
        pieces = pg.run(y_dev, wait=True)
        torch.cuda.synchronize()
        full = torch.cat(pieces, dim=2)  # [world, NS, n]
        ok = True
        detail = []
        if rank == 0:
            for r in range(world):
                if args.scaling == "weak":
                    y_r = torch.as_tensor(terminal_batch(n, 1 + r, ndt)[0], device=dev).contiguous()
                else:
                    y_r = torch.as_tensor(np.ascontiguousarray(terminal_batch(n * world, 1, ndt)[0][:, r * n:(r + 1) * n]), device=dev)
                ref = pu.solve_batch_soa(problem, opts, y_r)["y_end"]
                torch.cuda.synchronize()
                same = bool(torch.equal(ref, full[r]))
                detail.append(same)
                ok = ok and same
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        dist.broadcast(flag, 0)
        gather_check = "ok" if flag.item() == 1.0 else f"MISMATCH {detail}"
        if rank == 0 and gather_check != "ok":
            print(f"gather_check failed: per-rank bit equality {detail}", file=sys.stderr)

    # ---- roofline of the dominant kernel (this rank's own launches) -----------------------------------
    mpeaks = {}
    try:
        mpeaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def dense_roof(sec, n_loc):
        peak_gbs = mpeaks.get("hbm_gbs", 6650.0)
        bytes_alg = flops.dense_bytes(n_loc, 6, opts.max_steps, es) + 2 * 14 * n_loc * es
        return {"bound": "hbm", "achieved": bytes_alg / sec / 1e9, "peak": peak_gbs, "unit": "GB/s",
                "frac": bytes_alg / sec / 1e9 / peak_gbs, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if mpeaks else "6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": bytes_alg, "kernel_ms": sec * 1e3}

    def load_traffic(roof, workload):
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf) and n == (1 << 20):
            try:
                roof["traffic"] = json.load(open(tf)).get(f"{workload}_{args.dtype}_{args.method}")
            except Exception:
                pass

    if dense:
        roof = dense_roof(solve_s, n)
    else:
        fl_alg = flops.solve_flops(problem.system_id, method, my_attempts, n)
        sm_max = mpeaks.get("sm_max_mhz", 1965.0)
        theo = 148 * 128 * 2 * sm_max * 1e6 / 1e12 / (1 if args.dtype == "f32" else 2)
        roof = {"bound": "fp32_fma" if args.dtype == "f32" else "fp64_fma", "achieved": fl_alg / solve_s / 1e12,
                "peak": peak_tflops, "unit": "TFLOP/s", "frac": fl_alg / solve_s / 1e12 / peak_tflops, "traffic": None,
                "peak_source": "FMA-chain microbenchmark (pmp_fma_peak_cuda) in this process, burst, best of "
                               f"{peaks_by_variant}; MEASURED_PEAKS.json has no CUDA-core figure",
                "peak_theoretical": theo, "frac_of_theoretical": fl_alg / solve_s / 1e12 / theo,
                "algorithmic_flops_per_launch": fl_alg,
                "flops_per_attempt": flops.ATTEMPT_FLOPS[(problem.system_id, method)],
                "flops_per_rhs": flops.RHS_FLOPS[problem.system_id],
                "kernel_ms": solve_s * 1e3}
    load_traffic(roof, args.workload)

    # ---- strong scaling (BASELINE.json configs[3]: 2^20 in all over N GPUs), same run ------------------------------
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong and not dense:
        del pg
        main_run["pg"] = None
        saved = args.scaling
        args.scaling = "strong"
        ys_dev, _, _, ns_loc = shard_inputs("strong")
        args.scaling = saved
        sr = timed_steps(ys_dev, ns_loc, opts, True, args.steps, warm)
        sr_solo = timed_steps(ys_dev, ns_loc, opts, False, args.steps, 2)
        strong = {"trajectories_total": ns_loc * world, "trajectories_per_gpu": ns_loc,
                  "value": ns_loc * world * args.steps / (sr["total_ms"] * 1e-3), "unit": "trajectories/s",
                  "ms_per_step": sr["ms_per_step"], "ms_without_gather": sr_solo["ms_per_step"],
                  "kernel_ms": reduce_max(sr["solve_ms"]),
                  "limiter": "per-launch tail: a shard of 2^20/N trajectories is only a few trajectories per resident lane "
                             "(148 SMs x 384 lanes), so the wait for the slowest lanes of the last wave no longer amortises"}
        sr["pg"] = None

    # ---- dense workload (BASELINE.json configs[4], C5) attached to the default line ------------------------------
    dense_block = None
    if world == 1 and not dense and not args.no_dense and args.log2n <= 20:
        od = mk_opts(True)
        dr = timed_steps(y_dev, n, od, False, max(3, min(args.steps, 10)), 2)
        dense_block = {"workload": "same batch with SaveAt(steps) output: every slot of ts, t, v, x, vx written once",
                       "ms_per_step": dr["ms_per_step"], "value": n / (dr["ms_per_step"] * 1e-3), "unit": "trajectories/s",
                       "roofline": dense_roof(dr["solve_ms"] * 1e-3, n)}
        try:
            dense_block["roofline"]["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                f"dense_{args.dtype}_{args.method}")
        except Exception:
            pass
        dr["out"].clear()
        torch.cuda.empty_cache()

    # ---- forward closed loop under the value-network ensemble costate law attached to the default line -----
    forward_block = None
    if world == 1 and not dense and not args.no_forward_nn and args.dtype == "f32" and args.log2n <= 20:
        try:
            from approximate_optimal_control_b200 import nn_costate as NC
            nf = 1 << 14
            xf = torch.as_tensor(40.0 * y_np[:6, :nf], dtype=torch.float32, device=dev).contiguous()
            ens = NC.NnEnsemble.from_flax(NC.init_flax_params(0, 6, (64, 64, 64), 8), NC.DataNormaliser.identity(6), dev, torch.float32)
            fo = _capi.make_opts(dtype=_capi.F32, t0=0.0, t1=10.0, dt0=0.1, dtmin=0.05, dtmax=0.0, max_steps=256)
            fout, fts = {}, []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record(); pu.forward_batch_soa(problem, fo, xf, out=fout, nn=ens); b.record(); torch.cuda.synchronize()
                fts.append(a.elapsed_time(b))
            gpu_launches += 5
            fms = float(np.median(fts[1:]))
            fatt = int(fout["n_steps"].sum())
            fl = fatt * 6 * 8 * 2 * 2 * (6 * 64 + 64 * 64 * 2 + 64)  # attempts x stages x members x (forward + backward) x 2 flop x weights
            forward_block = {"workload": "forward closed-loop simulation of 2^14 flatquad states, t: 0 -> 10, Tsit5 + PID(dtmin = .05), costate law = "
                                         "gradient of the mean of an 8-member 6-64-64-64-1 softplus value-network ensemble (random flax-layout "
                                         "weights), levelsets.py:838-933",
                             "ms": fms, "value": nf / (fms * 1e-3), "unit": "trajectories/s", "rk_step_attempts_per_s": fatt / (fms * 1e-3),
                             "mlp_tflops_fp32_equivalent": fl / (fms * 1e-3) / 1e12,
                             "kernel": "forward_umma_kernel: tcgen05.mma kind::tf32 in 3xTF32 split precision, clusters of one CTA per member "
                                       "(pmp_forward_umma.cuh)"}
            del fout, xf, ens
            torch.cuda.empty_cache()
        except Exception as e:  # (a secondary block must not take the headline line down with it)
            forward_block = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- e2e: public API, host buffers ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from approximate_optimal_control_b200 import levelsets
        ap_e = dict(ap, pontryagin_solver_method=args.method, pontryagin_solver_save_steps=dense,
                    dtype="float32" if args.dtype == "f32" else "float64",
                    **({"pontryagin_solver_dt0": 1 / 32} if args.method == "rk4" else {}))
        _, P_lqr = pu.get_terminal_lqr(pp)
        # the reference's own call of the path: sols = jax.vmap(solve_backward_lqr, in_axes=(0, None))(xfs, algo_params)
        # (levelsets.py:576-601): host terminal samples xfs in; solve_backward_lqr builds v_f = 1/2 x'Px, vx_f = Px, t = 0
        solve_backward_lqr = levelsets.define_solve_backward_lqr(pp, ap_e, P_lqr)
        solve_backward, _ = pu.define_backward_solver(pp, ap_e)
        host_x = torch.as_tensor(st_np["x"]).pin_memory()
        host_state = {k: torch.as_tensor(v).pin_memory() for k, v in st_np.items()} if args.e2e_per_piece else None
        if args.e2e_per_piece:
            os.environ["B200PMP_HOST_PER_PIECE"] = "1"
        host_out = {k: torch.empty(v.shape, dtype=tdt).pin_memory() for k, v in st_np.items()}
        h2d = host_x.numel() * host_x.element_size()
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())
        if not dense:
            d2h += n * 4  # num_accepted_steps, num_steps, result come back with the endpoints, one packed word each

        def e2e_step():
            if dense:  # dense solutions stay on the device (8 GB per step); only the endpoints travel back
                ye = solve_backward_lqr(host_x.to(dev, non_blocking=True)).y_end
                for k in host_out:
                    host_out[k].copy_(ye[k], non_blocking=True)
                torch.cuda.synchronize()
                return
            # host in -> host out: pinned xfs goes to pmp_solve_host_lqr (csrc/pmp_host.cu): ONE persistent launch of the
            # integrator consumes the pieces as the copy engine delivers them and the endpoints + counters of every finished
            # piece are copied back to pinned host memory while later pieces are still being integrated
            if args.e2e_per_piece:
                sol = solve_backward(host_state)
            else:
                sol = solve_backward_lqr(host_x)
            assert not sol.y_end["x"].is_cuda and sol.y_end["x"].shape == host_x.shape

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e_s = reduce_max(time.perf_counter() - t0)
        if not dense:
            k, c = C.c_int(), C.c_int()
            pcs = L.pmp_solve_host_plan(C.byref(problem), C.byref(opts), n, 0, C.byref(k), C.byref(c))
            gpu_launches += (args.steps + 3) * (k.value if pcs > 0 else 0)
        e2e = {"value": n * world * args.steps / e_s, "unit": "trajectories/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "gather": "none: every rank returns its own shard's endpoints to its host; the NVLink gather is part of `value` only" if world > 1 else None,
               "api": "levelsets.define_solve_backward_lqr(...)(xfs): the reference's vmapped caller of the path (levelsets.py:576-601), "
                      "pinned host terminal samples xfs -> endpoints (x, t, v, vx) + counters in pinned host tensors through the C "
                      "entry pmp_solve_host_lqr (host pointers): one persistent integrator launch fed piece by piece by the "
                      "host->device stream, device->host copies of finished pieces under the integration of later ones; the "
                      "terminal v_f = 1/2 x'Px, vx_f = Px are formed on the device as the reference's vmap does"
                      if not dense else "define_solve_backward_lqr(...)(xfs on the device after an explicit host->device copy); "
                      "endpoints copied back to pinned host memory, dense nodes stay on the device"}

    # clocks: the K-step region lasts tens of ms, shorter than nvidia-smi's sampling period, so the
    # sampler runs from the first warm-up step to the end of the e2e loop (GPU busy throughout)
    clocks = sampler.stop(t_begin, time.time()) if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps + e2e loop"

    # ---- parity of this run's numbers + cpu baseline (rank 0, N=1 only) ----------------------------------
    cpu = None
    parity = None
    os.sched_setaffinity(0, all_cpus)  # the CPU legs below (oracle, OpenMP) use every host core the process was given
    if rank == 0 and world == 1 and not (args.no_cpu_baseline and args.no_parity):
        from oracle import oracle as O
    if rank == 0 and world == 1 and not args.no_parity and not dense:
        m = min(n, 1 << 12)
        tol = 1e-5
        rows = list(range(6)) + list(range(7, 14))
        got = main_run["out"]["y_end"][:, :m].double().cpu().numpy()
        orc = O.solve_batch(O.flatquad_problem(), oracle_opts(O, args), y_np[:, :m])
        y64 = terminal_batch(1 << args.log2n, 1 + rank, np.float64)[0][:, :m] if args.scaling == "weak" else \
            terminal_batch(1 << args.log2n, 1, np.float64)[0][:, :m]
        truth = O.solve_batch(O.flatquad_problem(), O.make_opts(dtype=O.F64, rtol=1e-12, atol=1e-12, max_steps=100000, dtmin=0.0),
                              y64)["y_end"]
        se = lambda a, b: (np.abs(a[rows] - b[rows]) / (tol + tol * np.abs(b[rows]))).max(0)
        eg, eo, d = se(got, truth), se(orc["y_end"].astype(np.float64), truth), se(got, orc["y_end"].astype(np.float64))
        gpu_steps = main_run["out"]["n_steps"][:m].cpu().numpy()
        parity = {"sample": f"first 2^{int(np.log2(m))} trajectories of the timed batch", "unit": "solver tolerance units: |a-b| / (tol + tol |b|), tol = 1e-5, max over x, lambda, v",
                  "gpu_vs_fp64_truth_p50": float(np.percentile(eg, 50)), "gpu_vs_fp64_truth_p99": float(np.percentile(eg, 99)),
                  "oracle_vs_fp64_truth_p50": float(np.percentile(eo, 50)), "oracle_vs_fp64_truth_p99": float(np.percentile(eo, 99)),
                  "gpu_vs_oracle_p50": float(np.percentile(d, 50)), "gpu_vs_oracle_p99": float(np.percentile(d, 99)),
                  "identical_step_counts": float((gpu_steps == orc["n_steps"]).mean()),
                  "results_identical": bool((main_run["out"]["result"][:m].cpu().numpy() == orc["result"]).all()),
                  "truth": "CPU oracle, fp64, Tsit5 at rtol = atol = 1e-12 (converged solution of the reference's ODE)",
                  "step_control": "unpinned (diffrax 0.4.1 absent: oracle and kernel implement the restatement of DESIGN.md 5)",
                  "note": "fp32 at T=3 amplifies rounding to tens of tolerance units (tests/test_oracle.py::test_fp32_reproducibility_floor); "
                          "the gate is gpu_vs_fp64_truth <= 1.25 x oracle_vs_fp64_truth per percentile (tests/test_gpu_parity.py)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nc = 1 << args.cpu_log2n
        ys = y_np[:, :nc] if nc <= n else y_np
        ncores = host_cores()
        oo = oracle_opts(O, args)
        O.solve_batch(O.flatquad_problem(), oo, ys[:, :4096], n_threads=ncores)

        def cpu_rate(o_, y_, threads, seconds, max_passes=64):
            t0 = time.perf_counter()
            passes = 0
            while True:  # bounded sample: whole passes over the batch until `seconds` of wall time
                O.solve_batch(O.flatquad_problem(), o_, y_, n_threads=threads)
                passes += 1
                dtc = time.perf_counter() - t0
                if dtc >= seconds or passes >= max_passes:
                    return passes * y_.shape[1] / dtc, passes, dtc

        v, passes, dtc = cpu_rate(oo, ys, ncores, args.cpu_seconds)
        v1, p1, d1 = cpu_rate(oo, ys[:, : 1 << 14], 1, min(4.0, args.cpu_seconds))
        o64 = oracle_opts(O, args, dtype="f64")
        v64, p64, d64 = cpu_rate(o64, ys[:, : 1 << 17].astype(np.float64), ncores, min(4.0, args.cpu_seconds))
        cpu = {"value": v, "unit": "trajectories/s", "cores": ncores, "kind": "port",
               "sample": f"{passes} passes over the first 2^{int(np.log2(ys.shape[1]))} trajectories of the same batch, same "
                         f"solver settings, {dtc:.1f} s wall on {ncores} threads; C++ restatement of the reference "
                         "(jax/diffrax not installable offline)",
               "one_core": {"value": v1, "cores": 1, "sample": f"{p1} passes over 2^14 trajectories, {d1:.1f} s"},
               "fp64": {"value": v64, "cores": ncores, "sample": f"{p64} passes over 2^17 trajectories in fp64, {d64:.1f} s"}}

    if rank == 0:
        line = {
            "metric": "PMP trajectories/sec", "value": value, "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": main_run["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, n, world),
            "rk_step_attempts_per_s": attempts_all * args.steps / total_s,
            "mean_attempts_per_trajectory": attempts_all / (n * world),
            "solve_ms_per_step": solve_s * 1e3,
            "roofline": roof, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(gpu_launches), "host_binding": host_binding,
            "gather": gather, "gather_check": gather_check, "strong": strong, "dense": dense_block, "forward_nn": forward_block,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
