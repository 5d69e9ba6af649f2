#!/usr/bin/env python
"""Where the time of the compact-column e2e step goes (bench.py e2e_compact_u64): upload+promote, member, prove, close."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import jolt_b200  # noqa: E402
from jolt_b200 import BatchMember, LOW_TO_HIGH, Polynomial, ProductMember, Session  # noqa: E402
from jolt_b200 import field as F  # noqa: E402

n, m, lg = 1 << 22, 2, 22
sess = Session(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
g = torch.Generator().manual_seed(0xC0)
cols = [torch.randint(-(2 ** 63), 2 ** 63 - 1, (n,), dtype=torch.int64, generator=g).pin_memory() for _ in range(m)]
cols_np = [c.numpy().view(np.uint64) for c in cols]
probe = ProductMember(sess, [Polynomial.from_small(sess, c) for c in cols_np], LOW_TO_HIGH)
ev = probe.prove_round_evals(None, 0)
claim = (ev[0] + ev[1]) % F.R_MOD
probe.close()
desc = [BatchMember(claim, 1, lg, 0)]
for rep in range(4):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    polys = [Polynomial.from_small(sess, c) for c in cols_np]
    torch.cuda.synchronize(); t.append(time.perf_counter())
    mem = ProductMember(sess, polys, LOW_TO_HIGH)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    r = jolt_b200.prove_batch_native(desc, [mem], lg, m, claim, seed=7, raw=True)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    f = mem.final_evals(raw=True)
    mem.close()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    print("rep", rep, "from_small x2 / member / prove / final+close (ms):", [round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])])
