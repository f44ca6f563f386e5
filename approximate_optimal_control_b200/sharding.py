"""Multi-GPU: shard terminal samples by index, one process per GPU, no collective during the
solve, ONE all-gather of the endpoint records afterwards.

The reference has no distributed code (single-process jax.vmap); this is the north-star extension:
"Trajectory batches shard across the 8 GPUs by terminal-sample index, with only an NCCL all-gather
of endpoint (x, lambda, v) datasets for the value-function fit."  torch.distributed is the plumbing
(backend nccl on GPUs, gloo in the CPU tests of this logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_INDEP_TIME = 0  # PMP_INDEP_TIME (include/b200pmp.h)


def bind_host_to_gpu(device_index: int) -> dict:
    """Run the calling process on the CPUs next to its GPU (one process per GPU: call it FIRST, before anything allocates
    pinned host memory).  The host <-> device streams of pmp_solve_host move 88 MB per 2^20 trajectories; on a two-socket box a
    rank that the launcher left on the other socket pushes all of it across the inter-socket link, which the eight ranks then
    share.  With the process on the GPU's own NUMA node, first-touch places its pinned buffers (cudaHostAlloc follows the calling
    thread's memory policy) on that node too.  Source of the mapping: sysfs (/sys/bus/pci/devices/<bdf>/local_cpulist,
    numa_node).  Never fails: returns what it did ({'bound': False, 'why': ...} when the platform does not say, e.g. a VM
    without NUMA information, or when the cpuset of the container does not intersect the GPU's CPUs)."""
    import os
    info = {"bound": False, "device": int(device_index)}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        info["pci"] = bdf
        node = int(open(base + "/numa_node").read().strip())
        info["numa_node"] = node
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        info["allowed_cpus"] = len(allowed)
        if node < 0:
            info["why"] = "the platform reports no NUMA node for this GPU"
            return info
        use = cpus & allowed
        if not use:
            info["why"] = "the GPU's CPUs are outside this process's cpuset"
            return info
        if use == allowed:
            info["why"] = "already there"
            info["bound"] = True
            info["cpus"] = len(use)
            return info
        os.sched_setaffinity(0, use)
        info["bound"] = True
        info["cpus"] = len(use)
    except Exception as e:  # noqa: BLE001  (a missing sysfs entry must not take the solve down)
        info["why"] = f"{type(e).__name__}: {e}"
    return info


def shard_size(n: int, world: int) -> int:
    """Every rank owns the same number of slots (the last ones may be padding)."""
    return (n + world - 1) // world


def shard_indices(n: int, rank: int, world: int, interleaved: bool = False) -> torch.Tensor:
    """Global indices of rank's shard, padded with -1 up to shard_size(n, world).
    contiguous: rank r owns [r*m, (r+1)*m).  interleaved: rank r owns r, r+W, r+2W, ... -- use it
    when the step count correlates with the sample index (e.g. the orbits theta sweep)."""
    m = shard_size(n, world)
    j = torch.arange(m, dtype=torch.int64)
    idx = j * world + rank if interleaved else rank * m + j
    return torch.where(idx < n, idx, torch.full_like(idx, -1))


def take_shard(y: torch.Tensor, n: int, rank: int, world: int, interleaved: bool = False) -> torch.Tensor:
    """y [rows, n] (struct-of-arrays) -> this rank's [rows, m] shard; padding columns are NaN, which
    the integrator retires immediately (result 4)."""
    idx = shard_indices(n, rank, world, interleaved).to(y.device)
    out = y[:, idx.clamp(min=0)].clone()
    if (idx < 0).any():
        if out.is_floating_point():
            out[:, idx < 0] = float("nan")
        else:
            out[:, idx < 0] = -1
    return out.contiguous()


def allgather_endpoints(local: torch.Tensor, n: int, interleaved: bool = False, group=None) -> torch.Tensor:
    """All-gather of per-rank records local [rows, m] into the global [rows, n] array, in the
    original sample order, on every rank.  One collective (all_gather_into_tensor)."""
    world = dist.get_world_size(group)
    rows, m = local.shape
    assert m == shard_size(n, world), "local shard has the wrong size"
    flat = torch.empty((world * rows, m), dtype=local.dtype, device=local.device)  # rank-major concat
    dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
    gathered = flat.view(world, rows, m)
    if interleaved:
        full = gathered.permute(1, 2, 0).reshape(rows, m * world)  # column j*W + r
    else:
        full = gathered.permute(1, 0, 2).reshape(rows, world * m)  # column r*m + j
    return full[:, :n].contiguous()


class PipelinedGather:
    """Integrate this rank's shard and gather the endpoint records of all ranks on every rank, with the
    transfer hidden behind compute:

      * within a step the shard can be cut into `chunks` pieces; the gather of piece k runs while piece
        k+1 is integrated (costs one extra kernel tail, ~50 us, per piece);
      * across steps (`run(..., wait=False)`): the gather of step s runs while step s+1 is integrated;
        two sets of gather buffers alternate, so the result of step s stays valid until run() of step s+2 is called
        (readers enqueued on the current stream before that call are safe, see run()).

    transport="p2p": every rank pushes its piece
        straight into the peers' gather buffers with peer-to-peer cudaMemcpyAsync over NVLink
        (copy engines: no SM is taken from the integrator, a persistent kernel that owns all of them),
        followed by one symmetric-memory barrier per step.
    transport="nccl": all_gather_into_tensor on a side stream (NCCL's kernels need SMs, so they only
        partly overlap with the integrator).
    transport="fused" (what "auto" picks when torch symmetric memory is available and the shard is one piece): the gather is
        part of the integrator (pmp_solve_batch_push_cuda): the warp that finishes a work-queue
        chunk of 32 trajectories stores its endpoint rows straight into every peer's gather buffer (peer memory over NVLink)
        while the rest of the shard is integrated.  No copy engine, no transfer left when the kernel ends; THREE buffer sets,
        so that the "every rank has consumed this set" barrier of a set completes a whole step before the kernel that
        overwrites it is launched.

        pg = PipelinedGather(problem, opts, n_local, dtype, device)
        pieces = pg.run(y_local)        # list of [world, NS, n_piece] tensors, rank-major, on every rank
    """

    def __init__(self, problem, opts, n_local: int, dtype, device, chunks: int = 1, group=None, transport: str = "auto"):
        from . import pontryagin_utils as pu
        self._pu = pu
        self.problem, self.opts, self.group = problem, opts, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ns = 2 * problem.nx + 2
        self.n_local = n_local
        chunks = max(1, min(chunks, n_local))
        edges = [n_local * c // chunks for c in range(chunks + 1)]
        self.slices = [(edges[c], edges[c + 1]) for c in range(chunks) if edges[c + 1] > edges[c]]
        # staging of the pieces' inputs (one piece: the caller's shard is integrated in place, nothing is staged)
        self.inputs = [torch.empty((self.ns, b - a) if len(self.slices) > 1 else (0,), dtype=dtype, device=device)
                       for a, b in self.slices]
        # "auto": the fused gather where it applies (one piece per step, time-parameterised endpoint solve with the default u*
        # selection), else peer-to-peer copies; NCCL where symmetric memory is not available
        fusable = (self.world > 1 and len(self.slices) == 1 and not opts.save_dense and opts.flags == 0
                   and opts.indep == _INDEP_TIME)
        if transport == "fused" and not fusable:
            transport = "auto" if self.world == 1 else "p2p"
        elif transport == "auto" and fusable:
            transport = "fused?"  # fused if symmetric memory comes up, NCCL otherwise
        self.nsets = (3 if transport.startswith("fused") else 2) if self.world > 1 else 1
        self.outs = [[{} for _ in self.slices] for _ in range(self.nsets)]
        self.side = torch.cuda.Stream(device=device) if self.world > 1 else None
        # a second copy stream: one stream = one copy engine pushed 7 x 59 MB in ~1.2 ms at N = 8, a little longer than the
        # integration of the next step, and step s+2 waited 0.06 ms per step for step s's pushes to release the local slot
        self.side2 = torch.cuda.Stream(device=device) if self.world > 2 else None
        self.transport = "none" if self.world == 1 else transport
        self.peer = None
        self.step = 0
        self._copied = [None] * self.nsets   # event: this set's local slot has been read by the pushes
        self._pending = False
        if self.world > 1 and transport in ("auto", "p2p", "fused", "fused?"):
            try:
                self._init_p2p(dtype, device)
                self.transport = "fused" if transport.startswith("fused") else "p2p"
            except Exception as e:  # symmetric memory unavailable: NCCL does the same job
                if transport in ("p2p", "fused"):
                    raise
                self.transport = "nccl"
                self.nsets = 2
                self.outs = [[{} for _ in self.slices] for _ in range(self.nsets)]
                self._copied = [None] * self.nsets
                self.p2p_error = repr(e)
        if self.transport == "fused":
            self._free = [None] * self.nsets  # event: every rank has consumed this set (it may be overwritten)
        if self.transport == "nccl":
            # per-piece gather buffers, rank-major: [world * NS, n_piece]
            self.gathered = [[torch.empty((self.world * self.ns, b - a), dtype=dtype, device=device) for a, b in self.slices]
                             for _ in range(self.nsets)]
        self._works = []

    def _init_p2p(self, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        per_set = sum((b - a) for a, b in self.slices) * self.world * self.ns
        flat = symm_mem.empty(per_set * self.nsets, dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(flat, grp)
        # carve the symmetric allocation into per-set, per-piece [world*NS, n_piece] buffers; the same
        # offsets are valid in every peer's copy
        self.gathered, self.peer = [], []
        # NVSwitch multicast mapping of the same allocation (NVLS), if the platform has one: the fused gather then needs one
        # store per row instead of one per row and peer.  B200PMP_GATHER_MULTICAST=0 keeps the per-peer stores.
        import os
        mc = 0
        if os.environ.get("B200PMP_GATHER_MULTICAST", "1") != "0":
            try:
                mc = int(self.hdl.multicast_ptr) if getattr(self.hdl, "has_multicast_support", False) else 0
                if mc:  # (the multicast mapping starts where the symmetric buffer does; `flat` may sit at an offset in it)
                    mc += flat.data_ptr() - int(self.hdl.buffer_ptrs[self.rank])
            except Exception:  # noqa: BLE001
                mc = 0
        self.multicast = bool(mc)
        self.mc_slot = []
        off = 0
        for s in range(self.nsets):
            gs, ps = [], []
            for a, b in self.slices:
                m = self.world * self.ns * (b - a)
                gs.append(flat[off:off + m].view(self.world * self.ns, b - a))
                ps.append([self.hdl.get_buffer(r, (self.world * self.ns, b - a), dtype, off) for r in range(self.world)])
                off += m
            self.gathered.append(gs)
            self.peer.append(ps)
            # this rank's block inside piece 0 of the set, through the multicast mapping
            self.mc_slot.append(mc + (off - sum(self.world * self.ns * (b - a) for a, b in self.slices)
                                      + self.rank * self.ns * (self.slices[0][1] - self.slices[0][0])) * flat.element_size() if mc else 0)
            # the integrator writes this rank's endpoints directly into its own slot of the gather buffer
            for o, g in zip(self.outs[s], gs):
                o["y_end"] = g[self.rank * self.ns:(self.rank + 1) * self.ns]

    def run(self, y_local: torch.Tensor, wait: bool = True):
        """y_local [NS, n_local] -> list of per-piece gathered endpoint arrays [world, NS, n_piece].
        wait=True: the current stream waits for the transfers of this step.  wait=False: the transfers
        keep running on the side stream (call wait() before reading the result / at the end).

        Buffer protocol (two sets, step s uses set s % 2): the arrays returned by step s stay valid until run() of
        step s+2 is CALLED; whatever reads them must have been enqueued on the current stream by then.  run() records
        that point of the current stream, and the side stream passes a symmetric-memory barrier behind it BEFORE any
        peer may push step s+2 into the set ("every rank has consumed set s % 2"); that barrier runs while the
        shard is being integrated, so it costs nothing.  A second barrier after the pushes says "every rank's
        step-s records have landed"."""
        if self.transport == "fused":
            return self._run_fused(y_local, wait)
        main = torch.cuda.current_stream()
        s = self.step % self.nsets
        self.step += 1
        outs = self.outs[s]
        rows = slice(self.rank * self.ns, (self.rank + 1) * self.ns)
        if self.world > 1 and self.transport == "p2p":
            consumed = torch.cuda.Event()
            consumed.record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(consumed)
                self.hdl.barrier(channel=1)  # all ranks are past their readers of this set: peers may overwrite it
        if self._copied[s] is not None:  # the pushes of step-2 must have finished reading this set's local slot
            main.wait_event(self._copied[s])
        for c, ((a, b), y_in, out) in enumerate(zip(self.slices, self.inputs, outs)):
            if len(self.slices) == 1 and y_local.is_contiguous():
                y_in = y_local  # one piece: integrate straight from the caller's shard (no staging copy)
            else:
                y_in.copy_(y_local[:, a:b])
            self._pu.solve_batch_soa(self.problem, self.opts, y_in, out)
            if self.world == 1:
                continue
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                if self.transport == "p2p":
                    if self.side2 is not None:  # the pushes to every other peer leave on the second stream
                        fork = torch.cuda.Event()
                        fork.record(self.side)  # (behind the "consumed" barrier and the integration of this piece)
                        with torch.cuda.stream(self.side2):
                            self.side2.wait_event(fork)
                            for d in range(2, self.world, 2):
                                r = (self.rank + d) % self.world
                                self.peer[s][c][r][rows].copy_(out["y_end"], non_blocking=True)
                            join = torch.cuda.Event()
                            join.record(self.side2)
                    for d in range(1, self.world, 2 if self.side2 is not None else 1):  # staggered: not everybody hits the same peer
                        r = (self.rank + d) % self.world
                        self.peer[s][c][r][rows].copy_(out["y_end"], non_blocking=True)
                    if self.side2 is not None:
                        self.side.wait_event(join)
                else:
                    self._works.append(dist.all_gather_into_tensor(self.gathered[s][c], out["y_end"], group=self.group,
                                                                   async_op=True))
        if self.world == 1:
            return [o["y_end"].view(1, self.ns, -1) for o in outs]
        with torch.cuda.stream(self.side):
            done = torch.cuda.Event()
            done.record(self.side)
            self._copied[s] = done
            if self.transport == "p2p":
                self.hdl.barrier(channel=0)  # every rank's pushes of this step have completed
        self._pending = True
        if wait:
            self.wait()
        return [g.view(self.world, self.ns, -1) for g in self.gathered[s]]

    def _run_fused(self, y_local: torch.Tensor, wait: bool):
        """Step t uses set t % 3.  The result of step t stays valid until run() of step t+2 is CALLED (as with the other
        transports).  run(t) records that point of the current stream; behind it the side stream passes the "consumed" barrier
        and records free[(t+1) % 3]: set (t-2) % 3 == (t+1) % 3 may be overwritten -- by this rank's kernel of step t+1 and, since
        all ranks pass the same barrier, by the peers' kernels of step t+1.  The kernel of step t waits for free[t % 3], which
        was recorded during step t-1: never on the critical path.  A second barrier behind the kernel says "every rank's
        records of step t have landed everywhere"."""
        main = torch.cuda.current_stream()
        t = self.step
        s = t % self.nsets
        self.step += 1
        out = self.outs[s][0]
        rows = slice(self.rank * self.ns, (self.rank + 1) * self.ns)
        consumed = torch.cuda.Event()
        consumed.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(consumed)
            self.hdl.barrier(channel=1)
            free = torch.cuda.Event()
            free.record(self.side)
            self._free[(t + 1) % self.nsets] = free
        if self._free[s] is not None and t > 0:
            main.wait_event(self._free[s])
        y_in = y_local if y_local.is_contiguous() else y_local.contiguous()
        slots = [self.peer[s][0][(self.rank + d) % self.world][rows] for d in range(1, self.world)]
        self._pu.solve_batch_soa(self.problem, self.opts, y_in, out, push=slots, push_mc=self.mc_slot[s] or None)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.hdl.barrier(channel=0)
        self._pending = True
        if wait:
            self.wait()
        return [g.view(self.world, self.ns, -1) for g in self.gathered[s]]

    def wait(self):
        """Make the current stream wait for every transfer issued so far."""
        if not self._pending:
            return
        for w in self._works:
            w.wait()
        self._works = []
        torch.cuda.current_stream().wait_stream(self.side)
        self._pending = False

    def attempts(self) -> torch.Tensor:
        return sum(o["n_steps"].sum() for o in self.outs[(self.step - 1) % self.nsets])
