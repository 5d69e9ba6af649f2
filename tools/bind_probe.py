#!/usr/bin/env python
"""Per-repetition timing of jb_table_bind at 2^24 (events + host wall + the library's own kernel timer): separates the
kernel from host-side stalls (allocation of the LowToHigh ping-pong buffer) inside the timed region."""
import ctypes
import sys
import time

import torch

sys.path.insert(0, ".")
from jolt_b200 import HIGH_TO_LOW, LOW_TO_HIGH, Polynomial, Session  # noqa: E402
from jolt_b200 import field as F  # noqa: E402
from oracle.coracle import rand_challenge  # noqa: E402  (tools/: input generation only)

n = 1 << 24
sess = Session(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.randint(0, 2 ** 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
src[:, 3] &= (1 << 60) - 1
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for order, oname in ((LOW_TO_HIGH, "l2h"), (HIGH_TO_LOW, "h2l")):
    for ch, cname in ((rand_challenge(5), "challenge125"), (F.to_limbs(F.R_MOD - 12345), "full254")):
        rows = []
        for rep in range(6):
            buf = src.clone()
            p = Polynomial.wrap_device(sess, buf.data_ptr(), n)
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            p.bind_with_order(ch, order)
            t1 = time.perf_counter()
            e1.record()
            e1.synchronize()
            rows.append((round(e0.elapsed_time(e1), 3), round((t1 - t0) * 1e3, 3)))
            p.free()
            del p, buf
        print(oname, cname, "event_ms/host_call_ms per rep:", rows)
