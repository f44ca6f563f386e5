"""N>1 path on real GPUs (skipped with fewer than 2 devices): shard by terminal-sample index, integrate
locally, gather the endpoints on every rank -- peer-to-peer (symmetric memory + copy engines), NCCL and the gather fused
into the integrator (pmp_solve_batch_push_cuda: stores to peer memory) must give the same, correct, global endpoint array."""
import os

import numpy as np
import pytest

from common import _capi, c_problem, flatquad_terminal

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_local, q):
    import torch
    import torch.distributed as dist
    from approximate_optimal_control_b200 import pontryagin_utils as pu, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    prob, _ = c_problem("flatquad")
    opts = _capi.make_opts(dtype=_capi.F64)
    n = n_local * world
    y_all = torch.as_tensor(flatquad_terminal(n), device=dev)
    y_loc = sharding.take_shard(y_all, n, rank, world)
    ref = pu.solve_batch_soa(prob, opts, y_all.contiguous())["y_end"]  # every rank solves everything once
    res = {}
    for transport, chunks in (("nccl", 1), ("p2p", 1), ("p2p", 3), ("fused", 1)):
        pg = sharding.PipelinedGather(prob, opts, n_local, torch.float64, dev, chunks=chunks, transport=transport)
        assert pg.transport == transport
        for _ in range(4):  # several times: buffers are reused across steps (two sets; three with the fused gather)
            pieces = pg.run(y_loc)
        torch.cuda.synchronize()
        full = torch.cat([p for p in pieces], dim=2)              # [world, NS, n_local]
        full = full.permute(1, 0, 2).reshape(14, n)               # contiguous sharding: column r*n_local + j
        res[(transport, chunks)] = bool(torch.equal(full, ref))
    # wait=False across steps with skewed ranks (write-after-read on the p2p gather buffers): the result of step s
    # is read AFTER step s+1 has been issued, by a reader that is slow on rank 0, while rank 1 races ahead into step s+2,
    # whose pushes land in the buffers of step s.  The 'consumed' barrier of PipelinedGather.run must hold them back.
    ys = [torch.as_tensor(flatquad_terminal(n, seed=11 + k), device=dev) for k in range(3)]
    refs = [pu.solve_batch_soa(prob, opts, y.contiguous())["y_end"].clone() for y in ys]
    locs = [sharding.take_shard(y, n, rank, world) for y in ys]
    for transport in ("p2p", "nccl", "fused"):
        pg = sharding.PipelinedGather(prob, opts, n_local, torch.float64, dev, chunks=1, transport=transport)
        pa = pg.run(locs[0], wait=False)
        pb = pg.run(locs[1], wait=False)
        pg.wait()
        if rank == 0:
            torch.cuda._sleep(int(1.5e9))  # ~0.8 s: rank 0's reader of step 0 is late
        got_a = torch.cat(pa, dim=2).clone()   # enqueued before run() of step 2 is called: must see step 0's records
        pc = pg.run(locs[2], wait=False)       # step 2 reuses the buffer set of step 0
        pg.wait()
        got_b, got_c = torch.cat(pb, dim=2).clone(), torch.cat(pc, dim=2).clone()
        if rank == world - 1:
            torch.cuda._sleep(int(1.0e9))      # now the last rank lags: its peers' step 3 must not overwrite what it still reads
        got_c2 = torch.cat(pc, dim=2).clone()
        pd = pg.run(locs[0], wait=False)       # step 3: the fused gather's third set wraps around to the set of step 0
        pe = pg.run(locs[1], wait=False)       # step 4
        pg.wait()
        got_d, got_e = torch.cat(pd, dim=2).clone(), torch.cat(pe, dim=2).clone()
        torch.cuda.synchronize()
        for tag, g, r in (("a", got_a, refs[0]), ("b", got_b, refs[1]), ("c", got_c, refs[2]), ("c2", got_c2, refs[2]),
                          ("d", got_d, refs[0]), ("e", got_e, refs[1])):
            res[(transport, "nowait", tag)] = bool(torch.equal(g.permute(1, 0, 2).reshape(14, n), r))
    # the C-ABI gather (pmp_comm_* / pmp_allgather_endpoints: NCCL without torch in the data path); the 128-byte id travels
    # through torch.distributed here only because the test already has it
    import ctypes as C
    L = _capi.lib()
    idbuf = (C.c_char * 128)()
    if rank == 0:
        _capi.check(L.pmp_comm_unique_id(idbuf))
    idt = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(idt, 0)
    idbuf = (C.c_char * 128).from_buffer_copy(bytes(idt.cpu().numpy().tobytes()))
    comm = C.c_void_p()
    _capi.check(L.pmp_comm_init(C.byref(comm), world, rank, idbuf))
    local = pu.solve_batch_soa(prob, opts, locs[0])["y_end"].contiguous()
    glob = torch.empty((world, 14, n_local), dtype=torch.float64, device=dev)
    _capi.check(L.pmp_allgather_endpoints(comm, local.data_ptr(), glob.data_ptr(), n_local, 14, _capi.F64,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    res[("c_abi_nccl",)] = bool(torch.equal(glob.permute(1, 0, 2).reshape(14, n), refs[0]))
    _capi.check(L.pmp_comm_destroy(comm))
    q.put((rank, res))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_solve_and_gather_two_gpus(world):
    """world 4 also covers the second copy stream of the peer pushes (sharding.PipelinedGather.side2, world > 2)"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, 3001, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, res in out:
        assert all(res.values()), (rank, res)
