"""Multi-rank path on CPU: world_size-2 gloo process group, each rank synthesises its column range and the
pieces are assembled.  The kernels run on the test-only SIMT emulator here (no GPU in this container); the same
sharding code drives one CUDA handle per GPU under torchrun in bench.py."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT

from slc_b200 import sharding


def test_balanced_ranges_cover_and_balance():
    cost = sharding.chain_cost_model(1000, 9, 29, 1.5)
    for world in (1, 2, 3, 8):
        b = sharding.balanced_ranges(cost, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and b.size == world + 1
        per = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert per.max() / per.mean() < 1.02
    # a grid-like cost profile (cheap ends): equal-count chunks are unbalanced, cost-balanced ones are not
    cost = np.r_[np.full(200, 0.05), np.full(600, 1.0), np.full(200, 0.05)]
    e = sharding.equal_ranges(1000, 4)
    b = sharding.balanced_ranges(cost, 4)
    imb = lambda r: max(cost[r[i]:r[i + 1]].sum() for i in range(4)) / (cost.sum() / 4)
    assert imb(e) > 1.5 and imb(b) < 1.05
    assert list(sharding.equal_ranges(10, 4)) == [0, 3, 6, 9, 10]          # ceil(10/4) = 3, like Iterators.partition


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import torch.distributed as dist
    import build_emu
    import slc_b200
    from slc_b200 import sharding as sh, workloads as W
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = slc_b200._capi.CApi(build_emu.build())
    Nx = 40
    A, B1, B2 = W.chain_plant(Nx)
    ranges = sh.balanced_ranges(sh.chain_cost_model(Nx, 9, 29, 1.5), world)
    c0, c1 = int(ranges[rank]), int(ranges[rank + 1])
    Sx, Su = W.chain_supports(Nx, 9, 29, 1.5, max(0, c0 - 1), min(Nx, c1 + 1))     # slab + halo only
    Phix, Phiu = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), [Sx, Su], I=[[j] for j in range(c0, c1)], api=api)
    # the same shard through the one-shot pipelined entry point (sls_synthesize with a column range), what bench.py's
    # e2e leg runs per rank: must equal the three-call result bit for bit
    os.environ["SLS_PIPE_CHUNKS"] = "2"
    syn = slc_b200.Synthesizer(api=api)
    syn.set_plant(A, B2)
    out = syn.synthesize(Sx, Su, col_range=(c0, c1))
    Px, Pu = syn.outputs_to_scipy(out)
    syn.close()
    for t in range(len(Phix)):
        for F, G in ((Px[t], Phix[t]), (Pu[t], Phiu[t])):
            assert np.array_equal(F.indptr, G.indptr) and np.array_equal(F.indices, G.indices) and np.array_equal(F.data, G.data)
    gathered = [None] * world
    dist.all_gather_object(gathered, (Phix, Phiu))
    if rank == 0:
        T = len(Phix)
        Fx = [sh.assemble_columns([g[0][t] for g in gathered], ranges, (Nx, Nx)) for t in range(T)]
        Fu = [sh.assemble_columns([g[1][t] for g in gathered], ranges, (B2.shape[1], Nx)) for t in range(T)]
        q.put((ranges, Fx, Fu))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_equals_single(emu_api):
    import slc_b200
    from slc_b200 import workloads as W
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ranges, Fx, Fu = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    Nx = 40
    A, B1, B2 = W.chain_plant(Nx)
    Sx, Su = W.chain_supports(Nx, 9, 29, 1.5)
    Gx, Gu = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), [Sx, Su], api=emu_api)
    assert 0 < ranges[1] < Nx
    for t in range(len(Gx)):
        for F, G in ((Fx[t], Gx[t]), (Fu[t], Gu[t])):
            assert np.array_equal(F.indptr, G.indptr) and np.array_equal(F.indices, G.indices)
            assert np.array_equal(F.data, G.data)          # bit-for-bit: same per-column code path
