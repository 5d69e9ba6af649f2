"""Per-round anatomy of one resident-kernel sumcheck (jb_ctx_run_log): device pass time, device idle between
rounds (publish -> host -> command -> decode), host turn-around. Usage: python tools/round_probe.py [log_n] [m]"""
import ctypes
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import jolt_b200
from jolt_b200 import BatchMember, LOW_TO_HIGH, Polynomial, ProductMember
from jolt_b200 import field as F
from jolt_b200.api import _p
from oracle.coracle import rand_limbs

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sess = jolt_b200.Session(0)
tabs = [rand_limbs(1 + j, 1 << lg) for j in range(m)]
polys0 = [Polynomial.new(sess, t) for t in tabs]
probe = ProductMember(sess, [p.clone() for p in polys0], LOW_TO_HIGH)
ev = probe.prove_round_evals(None, 0)
claim = (ev[0] + ev[1]) % F.R_MOD
probe.close()
best = None
for rep in range(8):
    mem = ProductMember(sess, [p.clone() for p in polys0], LOW_TO_HIGH)
    sess.synchronize()
    t0 = time.perf_counter()
    jolt_b200.prove_batch_native([BatchMember(claim, 1, lg, 0)], [mem], lg, m, claim, seed=3)
    dt = time.perf_counter() - t0
    mem.close()
    log = np.zeros((64, 8), dtype=np.uint64)
    n = ctypes.c_size_t()
    sess.check(sess.lib.jb_ctx_run_log(sess.h, _p(log), 64, ctypes.byref(n)))
    if best is None or dt < best[0]:
        best = (dt, log[: n.value].astype(np.int64))
dt, log = best
print(f"2^{lg} m={m}: {dt * 1e6:.1f} us wall for {len(log)} mailbox commands")
print(" cmd  pass_us  dev_idle_us  host_turn_us  host_wait_us | b0_pass  b0_arrive  last_seen  fold")
for k in range(len(log)):
    pass_us = (log[k, 1] - log[k, 0]) / 1e3
    idle = (log[k + 1, 0] - log[k, 1]) / 1e3 if k + 1 < len(log) else float("nan")
    turn = (log[k + 1, 2] - log[k, 3]) / 1e3 if k + 1 < len(log) else float("nan")
    wait = (log[k, 3] - log[k, 2]) / 1e3
    b0p, b0a, ls = ((log[k, 4 + i] - log[k, 0]) / 1e3 for i in range(3))
    print(f"{k:4d} {pass_us:8.2f} {idle:12.2f} {turn:13.2f} {wait:13.2f} | {b0p:7.2f} {b0a:9.2f} {ls:10.2f} {pass_us - ls:5.2f}")
tot_pass = float(np.sum(log[:, 1] - log[:, 0])) / 1e3
print(f"sum pass {tot_pass:.1f} us; device span {(log[-1, 1] - log[0, 0]) / 1e3:.1f} us; "
      f"host span {(log[-1, 3] - log[0, 2]) / 1e3:.1f} us")
