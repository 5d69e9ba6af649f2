"""A/B of the eq table's store flavour (JB_EQ_STORE=0 default caching / 1 streaming): run under ncu with
--metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum -k regex:eq_stream_kernel."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import jolt_b200
from jolt_b200 import EqPolynomial
from oracle.coracle import rand_challenge
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
sess = jolt_b200.Session(0)
r = np.stack([rand_challenge(9 + i) for i in range(lg)])
for _ in range(3):
    EqPolynomial.evals(sess, r).free()
sess.synchronize()
print("done", lg)
