"""Device Fr/Fq add/sub/mul vs the oracle, bit-exact (mirrors bn254_differential.rs:75-99 at scale)."""
import numpy as np
import pytest

import jolt_b200
from oracle import bn254 as O
from oracle import coracle as C
from gpu_util import EDGE_INTS, rand_limbs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    s = jolt_b200.Session(0)
    yield s
    s.close()


@pytest.mark.parametrize("fld,p", [(0, O.R_MOD), (1, O.Q_MOD)])
def test_vec_ops_bit_exact(sess, fld, p):
    n = 1 << 20
    a = rand_limbs(11 + fld, n)   # raw < 2^253 < both moduli: canonical for Fr and Fq
    b = rand_limbs(22 + fld, n)
    edge = C.ints_to_mont([e % p for e in EDGE_INTS], p)
    k = len(EDGE_INTS)
    # all edge x edge pairs up front
    a[: k * k] = np.repeat(edge, k, axis=0)
    b[: k * k] = np.tile(edge, (k, 1))
    for op in (0, 1, 2):
        got = sess.vec_op(fld, op, a, b)
        want = C.f_vec(fld, op, a, b)
        assert (got == want).all(), f"field {fld} op {op}: {(got != want).any(axis=1).sum()} mismatches"
    # neg and square (bn254_differential.rs:75-99 also pins these): -a == 0 - a, a^2 == a * a
    assert (sess.vec_op(fld, 4, a, b) == C.f_vec(fld, 1, np.zeros_like(a), a)).all()
    assert (sess.vec_op(fld, 5, a, b) == C.f_vec(fld, 2, a, a)).all()


def test_mul_by_challenge_limbs(sess):
    # montmul(a, [0,0,lo,hi]) via the 4-row path == generic Montgomery product
    n = 1 << 16
    a = rand_limbs(5, n)
    b = rand_limbs(6, n)
    b[:, 0] = 0
    b[:, 1] = 0
    b[:, 3] &= np.uint64((1 << 61) - 1)
    a[:9] = C.ints_to_mont(EDGE_INTS)
    got = sess.vec_op(0, 3, a, b)
    assert (got == C.f_vec(0, 2, a, b)).all()
    assert (sess.vec_op(0, 2, a, b) == got).all()
