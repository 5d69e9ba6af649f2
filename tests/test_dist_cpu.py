"""world_size-2 gloo test of the multi-GPU host logic (no GPU needed): the per-round exchange is an
integer all-reduce over 8 x u64 lanes of 32-bit limbs followed by the carry + mod-r fold (jb_lanes_reduce_host)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bn254 as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jolt_b200 import field as F
    from jolt_b200.dist import elems_to_lanes, lanes_to_ints, splitmix_challenge
    from jolt_b200.api import UnivariatePoly
    # each rank's "partial sums": 3 evaluations, including the extremes p-1 on every rank
    vals = O.random_fr(100 + rank, 3)
    vals[1] = O.R_MOD - 1
    lanes = torch.from_numpy(elems_to_lanes(F.ints_to_limbs(vals)).astype(np.int64))
    dist.all_reduce(lanes)
    got = lanes_to_ints(lanes.numpy().astype(np.uint64))
    ch = splitmix_challenge(7, UnivariatePoly.from_evals(got))
    q.put((rank, got, ch))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_lane_allreduce_is_exact(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    per_rank = [O.random_fr(100 + r, 3) for r in range(world)]
    for v in per_rank:
        v[1] = O.R_MOD - 1
    want = [sum(v[k] for v in per_rank) % O.R_MOD for k in range(3)]
    for rank, got, ch in res:
        assert got == want
    assert len({ch for _, _, ch in res}) == 1      # every rank derives the same challenge


def test_lanes_fold_extremes():
    from jolt_b200 import field as F
    from jolt_b200.dist import elems_to_lanes, lanes_to_ints
    lanes = elems_to_lanes(F.ints_to_limbs([O.R_MOD - 1, 0, 1]))
    for ranks in (1, 8, 1 << 20, (1 << 32) - 1):
        got = lanes_to_ints(lanes * np.uint64(ranks))
        assert got == [(O.R_MOD - 1) * ranks % O.R_MOD, 0, ranks % O.R_MOD]


def _shard_worker(rank, world, port, order, log_n, m, q):
    """The sharded member's partition + exchange + gather logic (capi.cu sharded_prove_round /
    gather_into_tail) replayed with the oracle's kernels standing in for the device kernels."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jolt_b200 import field as F
    from jolt_b200.api import UnivariatePoly
    from jolt_b200.dist import elems_to_lanes, lanes_to_ints, shard_of, splitmix_challenge
    glob = [O.random_fr(500 + j, 1 << log_n) for j in range(m)]
    tabs = [list(shard_of(g, rank, world, order)) for g in glob]
    polys, challenges = [], []
    bind = None
    while True:
        if bind is not None:
            tabs = [O.bind(t, bind, order) for t in tabs]
        if len(tabs[0]) == 1:
            break
        part = O.product_round_evals(tabs, m, order)                    # this rank's d+1 partial sums
        lanes = torch.from_numpy(elems_to_lanes(F.ints_to_limbs(part)).astype(np.int64))
        dist.all_reduce(lanes)                                          # the ONE exchange of a sharded round
        evals = lanes_to_ints(lanes.numpy().astype(np.uint64))
        poly = UnivariatePoly.from_evals(evals)
        polys.append(poly.coefficients)
        bind = splitmix_challenge(3, poly)
        challenges.append(bind)
    # gather: rank order is global order for contiguous blocks, strided shards interleave
    mine = torch.from_numpy(F.ints_to_limbs([t[0] for t in tabs]).astype(np.int64))
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    per_rank = [F.limbs_to_ints(p.numpy().astype(np.uint64)) for p in parts]      # [g][table]
    gathered = [[per_rank[g][j] for g in range(world)] for j in range(m)]          # global index j*G+g with len 1 == g
    q.put((rank, polys, challenges, gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("order", [O.HIGH_TO_LOW, O.LOW_TO_HIGH])
def test_sharded_partition_reproduces_global_rounds(order):
    world, log_n, m = 2, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, order, log_n, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]
    # the single-process run over the global tables, same stand-in transcript
    from jolt_b200.api import UnivariatePoly
    from jolt_b200.dist import splitmix_challenge
    tabs = [O.random_fr(500 + j, 1 << log_n) for j in range(m)]
    want_polys, want_ch = [], []
    for _ in range(log_n - 1):                      # the rounds the shards serve locally (local len 2^(log_n-1))
        poly = UnivariatePoly.from_evals(O.product_round_evals(tabs, m, order))
        want_polys.append(poly.coefficients)
        want_ch.append(splitmix_challenge(3, poly))
        tabs = [O.bind(t, want_ch[-1], order) for t in tabs]
    assert res[0][1] == want_polys and res[0][2] == want_ch
    assert res[0][3] == tabs                        # gathered tables == the global tables after the same binds
