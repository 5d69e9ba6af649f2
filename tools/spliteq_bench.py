"""Split-eq member vs the plain member over a materialised eq table: whole sumcheck at 2^22 (m = 2 witness
tables, degree 3), C++ engine, device-resident inputs."""
import json, sys, time, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import jolt_b200
from jolt_b200 import BatchMember, EqPolynomial, EqProductMember, LOW_TO_HIGH, Polynomial, ProductMember
from jolt_b200 import field as F
from oracle import coracle as C

sess = jolt_b200.Session(0)
for lg in [int(a) for a in sys.argv[1:] if a.isdigit()] or [20, 22]:
    n, m = 1 << lg, 2
    tabs = [C.rand_limbs(1 + j, n) for j in range(m)]
    w = np.stack([C.rand_challenge(100 + i) for i in range(lg)])
    base = [Polynomial.new(sess, t) for t in tabs]
    eqp = EqPolynomial.evals(sess, w)
    probe = ProductMember(sess, [eqp.clone()] + [p.clone() for p in base], LOW_TO_HIGH)
    ev = probe.prove_round_evals(None, 0)
    claim = (ev[0] + ev[1]) % F.R_MOD
    probe.close()
    out = {}
    for kind in ("split_eq", "materialised_eq"):
        best = 1e9
        for rep in range(5):
            if kind == "split_eq":
                mem = EqProductMember(sess, [p.clone() for p in base], w)
            else:
                mem = ProductMember(sess, [eqp.clone()] + [p.clone() for p in base], LOW_TO_HIGH)
            sess.synchronize()
            t0 = time.perf_counter()
            res = jolt_b200.prove_batch_native([BatchMember(claim, 1, lg, 0)], [mem], lg, m + 1, claim, seed=9, raw=True)
            dt = time.perf_counter() - t0
            best = min(best, dt)
            mem.close()
        out[kind] = (best, res[1].copy())
    assert (out["split_eq"][1] == out["materialised_eq"][1]).all(), "final claims differ"
    print(json.dumps(dict(kind="spliteq", log_n=lg, m=m, split_eq_ms=round(out["split_eq"][0] * 1e3, 3),
                          materialised_eq_ms=round(out["materialised_eq"][0] * 1e3, 3))), flush=True)
