"""Round-trip latency probe: time per round of the C++ engine (jb_prove_batch) at small sizes."""
import sys, time, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import jolt_b200
from jolt_b200 import BatchMember, Polynomial, ProductMember, LOW_TO_HIGH, HIGH_TO_LOW
from jolt_b200 import field as F
from oracle.coracle import rand_limbs

sess = jolt_b200.Session(0)
for lg in (2, 6, 10, 14, 18):
    tabs = [rand_limbs(1 + j, 1 << lg) for j in range(2)]
    for order in (LOW_TO_HIGH, HIGH_TO_LOW):
        polys0 = [Polynomial.new(sess, t) for t in tabs]
        probe = ProductMember(sess, [p.clone() for p in polys0], order)
        ev = probe.prove_round_evals(None, 0)
        claim = (ev[0] + ev[1]) % F.R_MOD
        probe.close()
        best = 1e9
        import ctypes
        d = (ctypes.c_double * 4)()
        for rep in range(6):
            mem = ProductMember(sess, [p.clone() for p in polys0], order)
            sess.synchronize()
            sess.lib.jb_ctx_diag(sess.h, d)
            t0 = time.perf_counter()
            jolt_b200.prove_batch_native([BatchMember(claim, 1, lg, 0)], [mem], lg, 2, claim, seed=3)
            dt = time.perf_counter() - t0
            sess.lib.jb_ctx_diag(sess.h, d)
            if dt < best:
                best, wait_us, waits = dt, d[0] / 1e3, d[1]
            mem.close()
        print(f"log_n={lg} order={order} total={best*1e6:.1f}us per_round={best*1e6/lg:.1f}us "
              f"device_wait={wait_us:.1f}us over {waits:.0f} waits ({wait_us/max(waits,1):.1f}us each)", flush=True)
    for p in polys0:
        p.free()
# raw single round at tiny size through ctypes
mem = ProductMember(sess, [Polynomial.new(sess, rand_limbs(9, 1 << 20)), Polynomial.new(sess, rand_limbs(8, 1 << 20))], LOW_TO_HIGH)
out = np.empty((3, 4), dtype=np.uint64)
from jolt_b200.api import _p
t0 = time.perf_counter()
for i in range(200):
    sess.check(sess.lib.jb_member_prove_round(mem.h, None, 0, None, _p(out)))
dt = time.perf_counter() - t0
print(f"eval-only round at 2^20 x2: {dt/200*1e6:.1f} us per call")
