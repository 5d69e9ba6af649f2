"""C oracle (oracle/oracle.c) == Python big-int oracle (oracle/bn254.py), plus the reference's
own property tests restated on the oracle."""
import numpy as np
import pytest

from oracle import bn254 as O
from oracle import coracle as C

EDGE = [0, 1, 2, O.R_MOD - 1, O.R_MOD - 2, (1 << 256) % O.R_MOD, (1 << 64) - 1, (1 << 128) - 1, (1 << 253) + 5]


def rnd(seed, n, p=O.R_MOD):
    return O.random_fr(seed, n, p)


@pytest.mark.parametrize("sel,p", [(0, O.R_MOD), (1, O.Q_MOD)])
def test_field_arith_matches_bigint(sel, p):
    # mirrors bn254_differential.rs:75-99 (add/sub/mul vs big-int)
    a = [e % p for e in EDGE] + rnd(1, 300, p)
    b = list(reversed([e % p for e in EDGE])) + rnd(2, 300, p)
    A, B = C.ints_to_mont(a, p), C.ints_to_mont(b, p)
    for op, fn in ((0, lambda x, y: (x + y) % p), (1, lambda x, y: (x - y) % p), (2, lambda x, y: x * y % p)):
        got = C.mont_to_ints(C.f_vec(sel, op, A, B), p)
        assert got == [fn(x, y) for x, y in zip(a, b)]
    # outputs are canonical (< p) Montgomery limbs
    out = C.f_vec(sel, 2, A, B)
    assert all(O.mont_raw(row) < p for row in out)


def test_to_mont_matches():
    canon = np.array([[(v >> (64 * k)) & O.MASK64 for k in range(4)] for v in EDGE], dtype=np.uint64)
    assert (C.to_mont(0, canon) == C.ints_to_mont(EDGE)).all()


@pytest.mark.parametrize("n", [1, 2, 5, 10, 11])
@pytest.mark.parametrize("order", [O.HIGH_TO_LOW, O.LOW_TO_HIGH])
def test_bind_matches(n, order):
    t = rnd(10 + n, 1 << n)
    s = rnd(99, 1)[0]
    got = C.mont_to_ints(C.bind(C.ints_to_mont(t), C.ints_to_mont([s])[0], order, threads=1 + (n % 2) * 3))
    assert got == O.bind(t, s, order)


def test_bind_sequences_equal_evaluate():
    # dense.rs:1095-1118: HighToLow over point[0..] and LowToHigh over point[n-1..0] both give evaluate(point)
    n = 6
    t = rnd(5, 1 << n)
    pt = rnd(6, n)
    a, b = list(t), list(t)
    for i in range(n):
        a = O.bind(a, pt[i], O.HIGH_TO_LOW)
        b = O.bind(b, pt[n - 1 - i], O.LOW_TO_HIGH)
    ev = O.evaluate(t, pt)
    assert a == [ev] and b == [ev]


@pytest.mark.parametrize("n", [0, 1, 5, 12])
def test_eq_matches(n):
    r = rnd(20 + n, n)
    sc = rnd(77, 1)[0]
    R = C.ints_to_mont(r).reshape(n, 4)
    assert C.mont_to_ints(C.eq_evals(R)) == O.eq_evals(r)
    assert C.mont_to_ints(C.eq_evals(R, C.ints_to_mont([sc])[0])) == O.eq_evals(r, sc)
    assert O.eq_evals(r) == O.eq_evaluations(r)           # eq.rs:642-662
    assert sum(O.eq_evals(r)) % O.R_MOD == 1               # eq.rs: sum == 1
    if n >= 11:
        assert C.mont_to_ints(C.eq_evals(R, None, threads=4)) == O.eq_evals(r)


def test_eq_pointwise_and_aligned_block():
    n = 7
    r = rnd(31, n)
    tab = O.eq_evals(r)
    for x in (0, 1, 77, 127):
        v = 1
        for i in range(n):
            bit = (x >> (n - 1 - i)) & 1                  # r[0] <-> MSB (eq.rs:218-219)
            v = v * (r[i] if bit else 1 - r[i]) % O.R_MOD
        assert tab[x] == v
    assert O.eq_evals_for_aligned_block(r, 32, 16) == tab[32:48]   # eq.rs:238-263


@pytest.mark.parametrize("m", [1, 2, 3])
@pytest.mark.parametrize("order", [O.HIGH_TO_LOW, O.LOW_TO_HIGH])
def test_sweep_matches(m, order):
    n = 7
    tabs = [rnd(40 + j, 1 << n) for j in range(m)]
    got = C.mont_to_ints(C.product_round_evals([C.ints_to_mont(t) for t in tabs], m, order, threads=3))
    assert got == O.product_round_evals(tabs, m, order)


def test_interpolation():
    coeffs = rnd(3, 4)
    ev = [O.uni_evaluate(coeffs, x) for x in range(4)]
    assert O.uni_from_evals(ev) == coeffs
    hint = (ev[0] + ev[1]) % O.R_MOD
    assert O.uni_from_evals_and_hint(hint, [ev[0]] + ev[2:]) == coeffs
    assert O.uni_compress(coeffs) == [coeffs[0]] + coeffs[2:]


def test_engine_dense_member_roundtrip():
    # tests.rs:1123-1290 shape: DenseMember::with_sum fixture, s(0)+s(1)==claim every round
    nr, total = 4, 90210
    evals = O.dense_member_with_sum(nr, total, 41)
    assert sum(evals) % O.R_MOD == total
    mem = O.ProductMember([evals])
    pt = O.synthetic_point(nr, 401)
    res = O.prove_batch([{"input_claim": total, "coefficient": 7, "rounds": nr, "offset": 0}], [mem],
                        nr, 1, 7 * total % O.R_MOD, lambda rnd_, poly: pt[rnd_])
    assert mem.final_evals()[0] == O.evaluate(evals, pt)
    assert res["member_claims"][0] == mem.final_evals()[0]


def _aff(P):
    if P is None:
        return np.zeros(8, dtype=np.uint64)
    return np.array(O.to_mont_limbs(P[0], O.Q_MOD) + O.to_mont_limbs(P[1], O.Q_MOD), dtype=np.uint64)


def _from_aff(xy, inf):
    if inf:
        return None
    return (O.from_mont_limbs(xy[:4], O.Q_MOD), O.from_mont_limbs(xy[4:], O.Q_MOD))


def test_g1_ops_match():
    G = O.G1_GEN
    assert O.g1_is_on_curve(G) and C.g1_on_curve(_aff(G))
    assert O.g1_scalar_mul(G, O.R_MOD) is None          # group order r
    for k in (1, 2, 3, 0xDEADBEEF, O.R_MOD - 1, rnd(8, 1)[0]):
        xy, inf = C.g1_scalar_mul(_aff(G), C.ints_to_mont([k])[0])
        assert _from_aff(xy, inf) == O.g1_scalar_mul(G, k)
    P, Q = O.g1_scalar_mul(G, 5), O.g1_scalar_mul(G, 11)
    for a, b in ((P, Q), (P, P), (P, O.g1_neg(P)), (None, Q), (P, None)):
        xy, inf = C.g1_add(_aff(a), _aff(b))
        assert _from_aff(xy, inf) == O.g1_add(a, b)


@pytest.mark.parametrize("n", [0, 1, 7, 64])
def test_msm_matches_naive(n):
    # group_laws.rs:69-79, :135-146
    ks = rnd(50, n)
    bases = [O.g1_scalar_mul(O.G1_GEN, k) for k in rnd(51, n)]
    sc = rnd(52, n)
    if n >= 7:
        sc[0] = 0
        bases[2] = bases[1]
        bases[4] = O.g1_neg(bases[3]); sc[4] = sc[3]
    B = np.array([_aff(b) for b in bases], dtype=np.uint64).reshape(n, 8)
    S = C.ints_to_mont(sc).reshape(n, 4)
    want = O.g1_msm_naive(bases, sc)
    if n:
        assert _from_aff(*C.g1_msm_naive(B, S)) == want
    assert _from_aff(*C.g1_msm_pippenger(B, S, 0, 2)) == want
    assert _from_aff(*C.g1_msm_pippenger(B, S, 7, 1)) == want
    if n <= 7:
        assert O.g1_msm_pippenger(bases, sc) == want


def test_hyperkzg_pieces():
    # kzg.rs witness_polynomial_division / eval tests; scheme.rs fold relation
    f = rnd(60, 9)
    u = rnd(61, 1)[0]
    h = O.compute_witness_polynomial(f, u)
    fu = O.eval_univariate(f, u)
    # f(x) - f(u) == h(x) * (x - u) at a random point
    x = rnd(62, 1)[0]
    assert (O.eval_univariate(f, x) - fu) % O.R_MOD == O.eval_univariate(h, x) * (x - u) % O.R_MOD
    assert C.mont_to_ints(C.witness_polynomial(C.ints_to_mont(f), C.ints_to_mont([u])[0])) == h
    assert C.mont_to_ints(C.eval_univariate(C.ints_to_mont(f), C.ints_to_mont([u])[0])) == [fu]
    ell = 4
    ev = rnd(63, 1 << ell)
    pt = rnd(64, ell)
    polys = O.fold_polynomials(ev, pt)
    assert [len(p) for p in polys] == [16, 8, 4, 2]
    # final fold with point[0] gives the multilinear evaluation (HyperKZG correctness)
    last = polys[-1]
    assert (last[0] + pt[0] * (last[1] - last[0])) % O.R_MOD == O.evaluate(ev, pt)
    srs = O.hyperkzg_setup_from_secret(rnd(65, 1)[0], 16)
    beta = rnd(65, 1)[0]
    P = np.array([_aff(O.G1_GEN)], dtype=np.uint64)
    got = C.g1_powers(5, P[0], C.ints_to_mont([beta])[0])
    assert [_from_aff(got[i], False) for i in range(5)] == srs[:5]


def test_inc_claim_reduction_restatement_matches_reference_tier_definition():
    """The oracle's restatement of the optimized IncClaimReduction kernel (paired-eq fusion, evaluations at {0, 2},
    s(1) from the claim; crates/jolt-kernels/src/optimized/inc_claim_reduction.rs:47-203) against the reference
    tier's definition: the six-table summand (eq(r_ram_rw) + g eq(r_ram_val)) RamInc + (g^2 eq(s_reg_rw) +
    g^3 eq(s_reg_val)) RdInc evaluated at t = 0, 1, 2 directly - the same lockstep the reference's own test runs."""
    n, gamma = 6, 29
    pts = [O.synthetic_point(n, s) for s in (3, 5, 7, 11)]
    ram, rd = O.random_fr(1, 1 << n), O.random_fr(2, 1 << n)
    k = O.IncClaimReductionKernel(pts, gamma, ram, rd)
    eqs = [O.eq_evals(pt) for pt in pts]
    scal = [1, gamma, gamma ** 2, gamma ** 3]
    tabs = [list(ram), list(rd)] + [list(e) for e in eqs]
    claim = sum((eqs[0][j] + gamma * eqs[1][j]) * ram[j] + (gamma ** 2 * eqs[2][j] + gamma ** 3 * eqs[3][j]) * rd[j]
                for j in range(1 << n)) % O.R_MOD
    ch = O.synthetic_point(n, 401)
    bind = None
    for rnd in range(n):
        if bind is not None:
            tabs = [O.bind(t, bind, O.LOW_TO_HIGH) for t in tabs]
        ev = []
        for t in range(3):
            acc = 0
            for y in range(len(tabs[0]) // 2):
                v = [tb[2 * y] + t * (tb[2 * y + 1] - tb[2 * y]) for tb in tabs]
                acc += (scal[0] * v[2] + scal[1] * v[3]) * v[0] + (scal[2] * v[4] + scal[3] * v[5]) * v[1]
            ev.append(acc % O.R_MOD)
        assert (ev[0] + ev[1]) % O.R_MOD == claim
        got = k.prove_round(bind, rnd, claim)
        assert got == O.uni_from_evals(ev), f"round {rnd}"
        bind = ch[rnd]
        claim = O.uni_evaluate(got, bind)
    k.finish_rounds(bind)
    tabs = [O.bind(t, bind, O.LOW_TO_HIGH) for t in tabs]
    assert k.output_claims() == {"ram_inc": tabs[0][0], "rd_inc": tabs[1][0]}
