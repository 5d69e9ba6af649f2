import sys, numpy as np
sys.path.insert(0, "/root/repo")
import jolt_b200
from jolt_b200 import Polynomial, ProductMember, LOW_TO_HIGH, HIGH_TO_LOW, UnivariatePoly
from jolt_b200 import field as F
from oracle import bn254 as O
sess = jolt_b200.Session(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tabs = [O.random_fr(500 + j, 1 << n) for j in range(2)]
ref = O.ProductMember(tabs, order)
gpu = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in tabs], order)
claim = sum(a * b for a, b in zip(*tabs)) % O.R_MOD
ch = O.synthetic_point(n, 401)
bind = None
for rnd in range(n):
    want = ref.prove_round(bind, rnd, claim)
    got = gpu.prove_round(bind, rnd, claim)
    print(rnd, got.coefficients == want, flush=True)
    bind = ch[rnd]
    claim = got.evaluate(bind)
ref.finish_rounds(bind); gpu.finish_rounds(bind)
print("final", gpu.final_evals() == ref.final_evals())
