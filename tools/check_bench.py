"""Debug: the bench's N = 1 flow (torch stream, wrapped device tables, timing on) - final-claim identity."""
import os, sys, pathlib
import numpy as np
import torch
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import jolt_b200
from jolt_b200 import BatchMember, LOW_TO_HIGH, Polynomial, ProductMember
from jolt_b200 import field as F

lg, m = 22, 2
n = 1 << lg
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sess = jolt_b200.Session(0, cuda_stream=stream.cuda_stream)
order = LOW_TO_HIGH

def synth(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, 2 ** 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= (1 << 60) - 1
    return t

base = [synth(0xB200 + j) for j in range(m)]
pb = [b.clone() for b in base]
probe = ProductMember(sess, [Polynomial.wrap_device(sess, b.data_ptr(), n) for b in pb], order)
ev = probe.prove_round_evals(None, 0)
claim = (ev[0] + ev[1]) % F.R_MOD
probe.close()
desc = [BatchMember(claim, 1, lg, 0)]
for timing in (False, True):
    for sync in (True, False):
        copies = [[b.clone() for b in base] for _ in range(4)]
        if sync:
            torch.cuda.synchronize()
        sess.timing_enable(timing, min_items=1 << (lg - 3))
        oks = []
        for k in range(4):
            polys = [Polynomial.wrap_device(sess, b.data_ptr(), n) for b in copies[k]]
            mem = ProductMember(sess, polys, order)
            res = jolt_b200.prove_batch_native(desc, [mem], lg, m, claim, seed=7, raw=True)
            fe = mem.final_evals(raw=True)
            mem.close()
            prod = 1
            for v in F.limbs_to_ints(fe):
                prod = prod * v % F.R_MOD
            oks.append(prod == F.from_limbs(res[1]))
        sess.timing_collect()
        print("timing", timing, "sync-before", sync, oks, flush=True)
