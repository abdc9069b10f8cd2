"""N>1 host logic on CPU: world_size-2 gloo processes shard the FASTA windows, pass the replay
generator state rank to rank (the draw chain is sequential across shards, include/qtipsim.h
qtipsim_run_params.cg1/cg2), and the rank-ordered concatenation of their plans equals the
single-process plan -- the same protocol bench.py runs over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = cases.CASES[name]()
        kw = cases.sim_kwargs(case)
        hs = util.HostSide(case)
        nwin = len(hs.windows(kw["chunksz"], int(hs.L.qthost_olap(hs.h))))
        lo, hi = rank * nwin // world, (rank + 1) * nwin // world
        state = torch.zeros(2, dtype=torch.int64)
        cg = kw["cg"]
        if rank > 0:
            dist.recv(state, src=rank - 1)
            cg = (int(state[0]), int(state[1]))
        rc, cg_out, draws, n = hs.plan(kw, lo, hi, cg=cg)
        assert rc == 0
        if rank < world - 1:
            state[0], state[1] = cg_out
            dist.send(state, dst=rank + 1)
        mine = [hs.plan_recs(c, n[c]) for c in range(4)]
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, draws, cg_out))
        if rank == 0:
            _, full_cg, full_draws, full_n = hs.plan(kw)
            full = [hs.plan_recs(c, full_n[c]) for c in range(4)]
            cat = [sum((g[0][c] for g in gathered), []) for c in range(4)]
            q.put((cat == full, sum(g[1] for g in gathered) == full_draws, gathered[-1][2] == full_cg))
        hs.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c4_mix_small", "c3_shape_small"])
def test_two_rank_shards_match_single_process(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    same_plan, same_draws, same_state = q.get(timeout=10)
    assert same_plan and same_draws and same_state
