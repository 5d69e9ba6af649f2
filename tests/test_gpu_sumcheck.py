"""Fused bind+eval ProveRounds member and the batched engine on the device vs the oracle
(mirrors jolt-sumcheck/src/tests.rs:1123-1290, tests/roundtrip.rs:26-183, optimized/parity.rs:79-118)."""
import numpy as np
import pytest

import jolt_b200
from jolt_b200 import HIGH_TO_LOW, LOW_TO_HIGH, BatchMember, Polynomial, ProductMember, UnivariatePoly
from jolt_b200 import field as F
from oracle import bn254 as O
from oracle import coracle as C
from gpu_util import rand_challenge, rand_full, rand_limbs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    s = jolt_b200.Session(0)
    yield s
    s.close()


def run_lockstep(sess, tables_int, order, challenges):
    """parity.rs:79-118 shape: drive the oracle member and the GPU member with identical
    (bind, round, claim); assert coefficient equality every round."""
    m = len(tables_int)
    n = len(tables_int[0]).bit_length() - 1
    ref = O.ProductMember(tables_int, order)
    gpu = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in tables_int], order)
    assert gpu.num_rounds() == n and gpu.degree() == m
    claim = sum(np.prod([t[i] for t in tables_int], dtype=object) for i in range(1 << n)) % O.R_MOD
    assert claim != 0, "reject zero initial claim (parity.rs)"
    bind = None
    for rnd in range(n):
        want = ref.prove_round(bind, rnd, claim)
        got = gpu.prove_round(bind, rnd, claim)
        assert got.coefficients == want, f"round {rnd}"
        assert (got.evaluate(0) + got.evaluate(1)) % O.R_MOD == claim
        bind = challenges[rnd]
        claim = got.evaluate(bind)
    ref.finish_rounds(bind)
    gpu.finish_rounds(bind)
    assert gpu.final_evals() == ref.final_evals()
    assert np.prod(gpu.final_evals(), dtype=object) % O.R_MOD == claim
    return gpu


@pytest.mark.parametrize("m", [1, 2, 3, 4])
@pytest.mark.parametrize("order", [HIGH_TO_LOW, LOW_TO_HIGH])
def test_product_member_lockstep(sess, m, order):
    n = 9
    tabs = [O.random_fr(500 + 10 * m + j, 1 << n) for j in range(m)]
    run_lockstep(sess, tabs, order, O.synthetic_point(n, 401))


def test_lockstep_125bit_challenges(sess):
    n = 8
    tabs = [O.random_fr(900 + j, 1 << n) for j in range(2)]
    ch = [O.from_mont_limbs(rand_challenge(40 + i)) for i in range(n)]
    for order in (HIGH_TO_LOW, LOW_TO_HIGH):
        run_lockstep(sess, tabs, order, ch)


def test_dense_member_fixture(sess):
    # tests.rs:1129-1135: evals = from_u64(seed + 31 i + 11), evals[0] += sum - current
    nr, total = 4, 90210
    evals = O.dense_member_with_sum(nr, total, 41)
    run_lockstep(sess, [evals], HIGH_TO_LOW, O.synthetic_point(nr, 401))


def test_roundtrip_degree2_and_3_fixtures(sess):
    # tests/roundtrip.rs:99-160 inputs
    f = [i + 1 for i in range(16)]
    g = [i * 3 + 7 for i in range(16)]
    run_lockstep(sess, [f, g], HIGH_TO_LOW, O.random_fr(1, 4))
    f = [i + 1 for i in range(8)]
    g = [i * 2 + 3 for i in range(8)]
    h = [i + 10 for i in range(8)]
    run_lockstep(sess, [f, g, h], HIGH_TO_LOW, O.random_fr(2, 3))


def test_round_check_failure_is_an_error(sess):
    # reference tier behaviour (naive.rs:301-308): with round verification on, a wrong claim is an error
    tabs = [O.random_fr(1, 16), O.random_fr(2, 16)]
    gpu = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in tabs])
    sess.set_verify_rounds(True)
    try:
        with pytest.raises(jolt_b200.SumcheckError, match="RoundCheckFailed"):
            gpu.prove_round(None, 0, 12345)
    finally:
        sess.set_verify_rounds(False)


@pytest.mark.parametrize("m", [1, 2, 3])
def test_verify_and_hint_modes_agree(sess, m):
    """s(1) derived from the claim (optimized tier) == s(1) computed (reference tier) == no claim at all."""
    n = 8
    tabs = [O.random_fr(700 + j, 1 << n) for j in range(m)]
    ch = O.random_fr(5, n)
    outs = []
    for mode in ("hint", "verify", "noclaim"):
        sess.set_verify_rounds(mode == "verify")
        gpu = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in tabs], LOW_TO_HIGH)
        claim = sum(np.prod([t[i] for t in tabs], dtype=object) for i in range(1 << n)) % O.R_MOD
        bind, seq = None, []
        for rnd in range(n):
            ev = gpu.prove_round_evals(bind, rnd, None if mode == "noclaim" else claim)
            seq.append(ev)
            bind = ch[rnd]
            claim = UnivariatePoly.from_evals(ev).evaluate(bind)
        outs.append(seq)
    sess.set_verify_rounds(False)
    assert outs[0] == outs[1] == outs[2]


def test_member_misuse_errors(sess):
    tabs = [O.random_fr(1, 4)]
    gpu = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in tabs])
    with pytest.raises(jolt_b200.JoltB200Error, match="NotFullyBound"):
        gpu.final_evals()
    with pytest.raises(jolt_b200.JoltB200Error):   # round index out of sequence
        gpu.prove_round_evals(None, 1)
    a = Polynomial.from_ints(sess, O.random_fr(1, 4))
    b = Polynomial.from_ints(sess, O.random_fr(1, 8))
    with pytest.raises(jolt_b200.JoltB200Error):   # length mismatch
        ProductMember(sess, [a, b])


def test_engine_on_gpu_matches_oracle_engine(sess):
    # batch of two members with different lengths + front-loaded padding (prover.rs:246-343)
    t_a = [O.random_fr(1, 64), O.random_fr(2, 64)]
    t_b = [O.dense_member_with_sum(3, 555, 41)]
    claim_a = sum(x * y for x, y in zip(*t_a)) % O.R_MOD
    desc = [dict(input_claim=claim_a, coefficient=O.random_fr(8, 1)[0], rounds=6, offset=0),
            dict(input_claim=555, coefficient=O.random_fr(9, 1)[0], rounds=3, offset=3)]
    total = (desc[0]["coefficient"] * claim_a + desc[1]["coefficient"] * 555 * 8) % O.R_MOD
    pts = O.synthetic_point(6, 401)
    want = O.prove_batch(desc, [O.ProductMember(t_a), O.ProductMember(t_b)], 6, 2, total, lambda r, c: pts[r])
    ma = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in t_a])
    mb = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in t_b])
    got = jolt_b200.prove_batch([BatchMember(**d) for d in desc], [ma, mb], 6, 2, total, lambda r, poly: pts[r])
    assert got.challenges == want["challenges"] and got.final_claim == want["final_claim"]
    assert got.member_claims == want["member_claims"]
    assert [p.coefficients for p in got.round_polynomials] == want["round_polys"]
    assert ma.final_evals() == [O.evaluate(t, pts) for t in t_a]
    # the C++ engine (one ABI call, no Python in the round loop) must agree as well
    ma2 = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in t_a])
    mb2 = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in t_b])
    nat = jolt_b200.prove_batch_native([BatchMember(**d) for d in desc], [ma2, mb2], 6, 2, total, lambda r, poly: pts[r])
    assert nat.challenges == want["challenges"] and nat.final_claim == want["final_claim"]
    assert nat.member_claims == want["member_claims"]
    assert [p.coefficients for p in nat.round_polynomials] == want["round_polys"]
    assert ma2.final_evals() == ma.final_evals()
    with pytest.raises(jolt_b200.SumcheckError):
        mc = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in t_a])
        jolt_b200.prove_batch_native([BatchMember(**desc[0])], [mc], 6, 2, total + 1)


def test_native_engine_splitmix_transcript_is_deterministic(sess):
    tabs = [O.random_fr(11, 256), O.random_fr(12, 256)]
    claim = sum(x * y for x, y in zip(*tabs)) % O.R_MOD
    runs = []
    for _ in range(2):
        mem = ProductMember(sess, [Polynomial.from_ints(sess, t) for t in tabs], LOW_TO_HIGH)
        res = jolt_b200.prove_batch_native([BatchMember(claim, 1, 8, 0)], [mem], 8, 2, claim, seed=7)
        runs.append((res.challenges, res.final_claim, mem.final_evals()))
    assert runs[0] == runs[1]
    ch, fin, fe = runs[0]
    # challenges are 125-bit [0,0,lo,hi] limbs; LowToHigh binds the last variable first
    assert all(list(F.to_limbs(c))[:2] == [0, 0] for c in ch)
    assert fe == [O.evaluate(t, list(reversed(ch))) for t in tabs]
    assert fe[0] * fe[1] % O.R_MOD == fin


@pytest.mark.parametrize("order", [HIGH_TO_LOW, LOW_TO_HIGH])
def test_sumcheck_2pow16_vs_c_oracle(sess, order):
    """BASELINE config 1: 2^16 sumcheck, every round polynomial and the final eval compared
    limb-for-limb with the 1-thread C oracle; m = 2."""
    n, m = 16, 2
    tabs = [rand_limbs(0xB200 + j, 1 << n) for j in range(m)]
    gpu = ProductMember(sess, [Polynomial.new(sess, t) for t in tabs], order)
    cur = [t.copy() for t in tabs]
    bind = None
    for rnd in range(n):
        if bind is not None:
            cur = [C.bind(t, bind, order) for t in cur]
        want = C.mont_to_ints(C.product_round_evals(cur, m, order))
        got = gpu.prove_round_evals(bind, rnd, (want[0] + want[1]) % O.R_MOD)
        assert got == want, f"round {rnd}"
        bind = rand_challenge(1000 + rnd) if rnd % 2 else rand_full(1000 + rnd)
    cur = [C.bind(t, bind, order) for t in cur]
    gpu.finish_rounds(bind)
    assert gpu.final_evals() == [C.mont_to_ints(t)[0] for t in cur]


@pytest.mark.parametrize("order", [LOW_TO_HIGH, HIGH_TO_LOW])
def test_sumcheck_2pow22_every_round_vs_c_oracle(sess, order):
    """BASELINE config 2 size: ALL 22 rounds at 2^22 (m = 2) against the threaded C oracle - every round polynomial
    limb for limb and the final evaluations (the oracle's rounds halve, so the whole run costs two first rounds) - plus
    the size-independent property: s(0) + s(1) == claim in every round and the run ends at prod f_j(point)."""
    n, m = 22, 2
    thr = C.max_threads()
    tabs = [rand_limbs(0xB200 + j, 1 << n) for j in range(m)]
    gpu = ProductMember(sess, [Polynomial.new(sess, t) for t in tabs], order)
    cur = tabs
    bind, claim = None, None
    for rnd in range(n):
        if bind is not None:
            cur = [C.bind(t, bind, order, thr) for t in cur]
        want = C.mont_to_ints(C.product_round_evals(cur, m, order, thr))
        claim = (want[0] + want[1]) % O.R_MOD if claim is None else claim
        assert (want[0] + want[1]) % O.R_MOD == claim, f"round {rnd}: the oracle's own round check"
        got = gpu.prove_round_evals(bind, rnd, claim)
        assert got == want, f"round {rnd}"
        poly = UnivariatePoly.from_evals(got)
        bind = rand_challenge(2000 + rnd) if rnd % 3 else rand_full(2000 + rnd)   # 125-bit challenges and full scalars mixed
        claim = poly.evaluate(F.from_limbs(bind))
    cur = [C.bind(t, bind, order, thr) for t in cur]
    gpu.finish_rounds(bind)
    fe = gpu.final_evals()
    assert fe == [C.mont_to_ints(t)[0] for t in cur]
    assert fe[0] * fe[1] % O.R_MOD == claim


def test_constant_polynomial_known_answers(sess):
    # crates/jolt-sumcheck/tests/soundness.rs:440-462: f = 7 on {0,1}^3 -> sum 56, final evaluation 7 at ANY point;
    # every round polynomial of a constant table is itself constant: s_k(X) = 7 * 2^(2-k)
    for order in (HIGH_TO_LOW, LOW_TO_HIGH):
        gpu = run_lockstep(sess, [[7] * 8], order, O.random_fr(9, 3))
        assert gpu.final_evals() == [7]
    gpu = ProductMember(sess, [Polynomial.from_ints(sess, [7] * 8)], HIGH_TO_LOW)
    claim, bind = 56, None
    for rnd, want in enumerate((28, 14, 7)):
        poly = gpu.prove_round(bind, rnd, claim)
        assert poly.coefficients[0] == want and all(c == 0 for c in poly.coefficients[1:])
        bind = O.synthetic_point(3, 401)[rnd]
        claim = poly.evaluate(bind)
    gpu.finish_rounds(bind)
    assert gpu.final_evals() == [7] and claim == 7


def test_small_scalar_accumulator_known_answer_on_device(sess):
    # crates/jolt-field/src/bn254/mont.rs:663-676 (3*16 + 5*(-7) + 11*1 + 9*(-13) + 2*7 == -79), with the device doing
    # the signed promotions (jb_table_upload_small) and the products (jb_vec_op)
    a = Polynomial.from_small(sess, np.array([3, 5, 11, 9, 2, 0, 0, 0], dtype=np.uint64)).evals()
    s = Polynomial.from_small(sess, np.array([16, -7, 1, -13, 7, 0, 0, 0], dtype=np.int64)).evals()
    prods = F.limbs_to_ints(sess.vec_op(0, 2, a, s))
    assert sum(prods) % O.R_MOD == O.R_MOD - 79
