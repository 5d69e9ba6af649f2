"""Index-sharded product sumcheck across the GPUs of one box (one process per GPU, torch.distributed
for the plumbing) - SURVEY.md section 8e.

Partitioning: LowToHigh binding pairs (2i, 2i+1) (crates/jolt-poly/src/dense.rs:236-238), so rank g
owning the CONTIGUOUS global block [g*n, (g+1)*n) keeps every pair local while its shard is longer
than one element. Per round each rank produces degree+1 partial sums (its fused bind+eval pass); they
leave the kernel as 8 x u64 lanes of 32-bit limbs, ONE integer all-reduce (ncclSum, exact) combines
them, and the carry + mod-r fold happens on the O(degree) result. No table data crosses NVLink until
the shards are tiny: at 2^GATHER_LOG entries per rank the bound shards are all-gathered once (rank
order == global order) and every rank finishes the remaining rounds redundantly with no further
communication - results are identical on every rank and identical to the single-GPU run on the same
global polynomial (the stand-in transcript depends on the round polynomials only)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from . import field as F
from .api import (BatchMember, LOW_TO_HIGH, Polynomial, ProductMember, ProvedBatch, Session, SumcheckError,
                  UnivariatePoly, _p, prove_batch_native)

GATHER_LOG = 12  # all-gather the shards once each is <= 2^12 entries (128 KiB per table per rank)


def lanes_to_ints(lanes_host: np.ndarray) -> list[int]:
    """(count, 8) u64 lanes (sums of 32-bit limbs) -> canonical field values, via the C-ABI host fold."""
    lanes_host = np.ascontiguousarray(lanes_host, dtype=np.uint64).reshape(-1, 8)
    out = np.empty((lanes_host.shape[0], 4), dtype=np.uint64)
    st = _lib.load().jb_lanes_reduce_host(_p(lanes_host), lanes_host.shape[0], _p(out))
    if st != _lib.JB_OK:
        raise RuntimeError("jb_lanes_reduce_host failed")
    return F.limbs_to_ints(out)


def elems_to_lanes(limbs: np.ndarray) -> np.ndarray:
    """(count, 4) u64 Montgomery limbs -> (count, 8) lanes of 32-bit limbs (host mirror of the kernel epilogue)."""
    a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    lo = a & np.uint64(0xFFFFFFFF)
    hi = a >> np.uint64(32)
    return np.stack([lo, hi], axis=2).reshape(-1, 8)


def splitmix_challenge(seed: int, poly: UnivariatePoly) -> int:
    """The stand-in transcript of the C++ engine (jb_absorb_round_splitmix125), called from Python so
    sharded and single-GPU runs derive identical challenges."""
    lib = _lib.load()
    coeffs = F.ints_to_limbs(poly.coefficients)
    out = np.zeros(4, dtype=np.uint64)
    s = ctypes.c_uint64(seed)
    lib.jb_absorb_round_splitmix125(ctypes.cast(ctypes.byref(s), ctypes.c_void_p), 0, _p(coeffs), coeffs.shape[0], _p(out))
    return F.from_limbs(out)


class ShardedProductSumcheck:
    """Drives one rank's shard (a ProductMember over its contiguous block, LowToHigh)."""

    def __init__(self, sess: Session, member: ProductMember, m: int, log_n_local: int, dist, seed: int,
                 gather_log: int = GATHER_LOG):
        import torch
        self.torch = torch
        self.s, self.mem, self.m, self.log_n, self.dist, self.seed = sess, member, m, log_n_local, dist, seed
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        assert self.world & (self.world - 1) == 0, "world size must be a power of two"
        self.log_g = self.world.bit_length() - 1
        self.gather_log = min(gather_log, log_n_local)

    def prove(self, claimed_sum: int | None = None):
        torch, lib, mem, m = self.torch, self.s.lib, self.mem, self.m
        lanes = torch.zeros((m + 1) * 8, dtype=torch.int64, device="cuda")
        out = np.empty((m + 1, 4), dtype=np.uint64)
        challenges, polys = [], []
        bind, claim = None, claimed_sum
        sharded_rounds = self.log_n - self.gather_log
        for rnd in range(sharded_rounds):
            b = None if bind is None else F.to_limbs(bind)
            skip = claim is not None  # s(1) = claim - s(0): one point fewer to compute and to reduce
            cnt = m if skip else m + 1
            self.s.check(lib.jb_member_prove_round_partials(mem.h, _p(b) if b is not None else None, rnd,
                                                            1 if skip else 0, ctypes.c_void_p(lanes.data_ptr())))
            self.dist.all_reduce(lanes[: cnt * 8])  # integer sum of 32-bit limbs: exact
            self.s.check(lib.jb_partials_finalize(self.s.h, ctypes.c_void_p(lanes.data_ptr()), cnt, _p(out)))
            ev = F.limbs_to_ints(out[:cnt])
            if skip:
                ev.insert(1, (claim - ev[0]) % F.R_MOD)
            poly = UnivariatePoly.from_evals(ev)
            c = splitmix_challenge(self.seed, poly)
            claim = poly.evaluate(c)
            challenges.append(c)
            polys.append(poly)
            bind = c
        # apply the pending bind, then gather the (now small) shards: rank order == global order
        if bind is not None:
            self.s.check(lib.jb_member_finish_rounds(mem.h, _p(F.to_limbs(bind))))
        shard = 1 << self.gather_log
        gathered = []
        for j in range(m):
            local = torch.empty((shard, 4), dtype=torch.int64, device="cuda")
            n_out = ctypes.c_size_t()
            self.s.check(lib.jb_member_export_table(mem.h, j, ctypes.c_void_p(local.data_ptr()), shard, ctypes.byref(n_out)))
            assert n_out.value == shard
            full = torch.empty((self.world * shard, 4), dtype=torch.int64, device="cuda")
            self.dist.all_gather_into_tensor(full, local)
            gathered.append(full)
        tail_polys = [Polynomial.wrap_device(self.s, g.data_ptr(), g.shape[0]) for g in gathered]
        tail = ProductMember(self.s, tail_polys, LOW_TO_HIGH)
        tail_rounds = self.gather_log + self.log_g
        if claim is None:  # no sharded round ran: derive the claim from the gathered tables
            ev = tail.prove_round_evals(None, 0)
            claim = (ev[0] + ev[1]) % F.R_MOD
            tail.close()
            tail = ProductMember(self.s, [Polynomial.wrap_device(self.s, g.data_ptr(), g.shape[0]) for g in gathered],
                                 LOW_TO_HIGH)
        res = prove_batch_native([BatchMember(claim, 1, tail_rounds, 0)], [tail], tail_rounds, m, claim, seed=self.seed)
        fe = tail.final_evals()
        tail.close()
        self._keep = gathered
        return ProvedBatch(challenges + res.challenges, res.final_claim, res.member_claims, polys + res.round_polynomials), fe
