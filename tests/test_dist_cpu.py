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
