"""Parity at BASELINE.json's full sizes through closed forms (size-independent properties).

Over the synthetic bases (i + 1) * G an MSM collapses to ONE scalar multiplication,
    msm(s) = (sum_i s_i (i + 1) mod r) * G,
and the weighted sum is pure field arithmetic - the side of the oracle that IS pinned by the reference's golden
vectors. So MSM 2^24 (config 3) and every group element of a HyperKZG open at ell = 22 (kzg_commit, the ell - 1
intermediate commitments, the three witness commitments; crates/jolt-hyperkzg/src/scheme.rs:122-158, kzg.rs:15-126)
are checked exactly at full size; the table-sized field work of the open is the C oracle's."""
import numpy as np
import pytest

import jolt_b200
from jolt_b200 import G1Bases, HyperKZG, Polynomial, g1_jacobian_to_affine
from jolt_b200 import field as F
from oracle import bn254 as O
from oracle import coracle as C
from gpu_util import rand_challenge, rand_limbs

pytestmark = pytest.mark.gpu

G_LIMBS = np.array(O.to_mont_limbs(1, O.Q_MOD) + O.to_mont_limbs(2, O.Q_MOD), dtype=np.uint64)
G_AFF = (1, 2)
RINV = pow(1 << 256, -1, O.R_MOD)


@pytest.fixture(scope="module")
def sess():
    s = jolt_b200.Session(0)
    yield s
    s.close()


def weighted_sum(limbs: np.ndarray) -> int:
    """sum_i value_i * (i + 1) mod r for (n, 4) Montgomery limbs: the raw 256-bit integers are summed limb half by
    limb half in u64 (64-term chunks cannot overflow: 2^32 * 2^25 * 2^6 = 2^63), then one Montgomery decode."""
    a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    n = a.shape[0]
    if n == 0:
        return 0
    pad = (-n) % 64
    w = np.arange(1, n + 1, dtype=np.uint64)
    total = 0
    for k in range(4):
        for half in (0, 1):
            part = (a[:, k] >> np.uint64(32 * half)) & np.uint64(0xFFFFFFFF)
            prod = part * w
            if pad:
                prod = np.concatenate([prod, np.zeros(pad, dtype=np.uint64)])
            chunks = prod.reshape(-1, 64).sum(axis=1, dtype=np.uint64)
            total += sum(int(x) for x in chunks) << (64 * k + 32 * half)
    return total * RINV % O.R_MOD


def expect_point(scalar: int):
    return O.g1_scalar_mul(G_AFF, scalar % O.R_MOD) if scalar % O.R_MOD else None


def test_weighted_sum_helper_against_python_ints():
    limbs = rand_limbs(5, 1000)
    vals = C.mont_to_ints(limbs)
    assert weighted_sum(limbs) == sum(v * (i + 1) for i, v in enumerate(vals)) % O.R_MOD


@pytest.mark.parametrize("log_n,precompute", [(20, False), (24, False), (24, True)])
def test_msm_closed_form_at_baseline_sizes(sess, log_n, precompute):
    """BASELINE config 3: G1 MSM over 2^20 / 2^24 uniform scalars, exact."""
    n = 1 << log_n
    rng = np.random.Generator(np.random.PCG64(0x5CA1A2 + log_n))
    sc = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64(((1 << 64) - 1) >> 3)      # raw < 2^253 < r: canonical Montgomery limbs
    bases = G1Bases.generate_multiples(sess, G_LIMBS, n)
    if precompute:
        bases.precompute()
    got = g1_jacobian_to_affine(bases.msm(sc))
    assert got == expect_point(weighted_sum(sc))
    # the legacy small-scalar facade at the same size (msm_u64): the same closed form on the integers
    if log_n <= 20:
        col = sc[:, 0].copy()
        weights = np.arange(1, n + 1, dtype=object)
        s_small = int((col.astype(object) * weights).sum()) % O.R_MOD
        assert g1_jacobian_to_affine(bases.msm_small(col)) == expect_point(s_small)
    bases.free()


def test_hyperkzg_commit_and_open_closed_form_ell22(sess):
    """BASELINE-size HyperKZG: ell = 22. Every commitment of the open is (weighted sum) * G; the evaluations v are
    compared with the C oracle's Horner evaluations exactly."""
    ell = 22
    n = 1 << ell
    thr = C.max_threads()
    bases = G1Bases.generate_multiples(sess, G_LIMBS, n)
    bases.precompute()
    evals = rand_limbs(2200, n)
    point = np.stack([rand_challenge(7 + i) if i % 2 else rand_limbs(9 + i, 1)[0] for i in range(ell)])
    r_int, q_int = O.random_fr(1022, 2)
    poly = Polynomial.new(sess, evals)
    assert g1_jacobian_to_affine(HyperKZG.commit(bases, poly)) == expect_point(weighted_sum(evals))
    proof = HyperKZG.open(bases, poly, point, lambda com: r_int, lambda v: q_int)
    # oracle: folds (scheme.rs:88-114), evaluations at r, -r, r^2 (kzg.rs:85-98), RLC, witness polynomials (kzg.rs:34-46)
    polys = [evals]
    for i in range(1, ell):
        polys.append(C.bind(polys[-1], point[ell - i], O.LOW_TO_HIGH, thr))
    assert [g1_jacobian_to_affine(c) for c in proof.com] == [expect_point(weighted_sum(p)) for p in polys[1:]]
    u = [r_int % O.R_MOD, (-r_int) % O.R_MOD, r_int * r_int % O.R_MOD]
    v = [[C.mont_to_ints(C.eval_univariate(p, C.ints_to_mont([ui])[0]))[0] for p in polys] for ui in u]
    assert proof.v == v
    b = np.zeros((n, 4), dtype=np.uint64)
    qj = 1
    for p in polys:
        scaled = C.f_vec(0, 2, p, np.tile(C.ints_to_mont([qj]), (p.shape[0], 1)))
        b[: p.shape[0]] = C.f_vec(0, 0, b[: p.shape[0]], scaled)
        qj = qj * q_int % O.R_MOD
    want_w = [expect_point(weighted_sum(C.witness_polynomial(b, C.ints_to_mont([ui])[0]))) for ui in u]
    assert [g1_jacobian_to_affine(x) for x in proof.w] == want_w
    assert (poly.evals() == evals).all()
    bases.free()
