"""Compact (primitive-integer) columns on the device vs the oracle.
  * promotion F::from(T)            crates/jolt-field/src/bn254/mod.rs:265-298 (from_u64/from_i64/from_u128/from_i128)
  * Polynomial<T>::bind_to_field    crates/jolt-poly/src/dense.rs:129-142 (test bind_matches_bind_to_field, :613-625)
  * small-scalar MSMs               crates/jolt-prover-legacy/src/msm/mod.rs:27-158 (msm_u8 .. msm_i128, msm_s64 / msm_s128,
                                    msm_binary), batch_msm / batch_msm_univariate :160-181
Every result is compared bit-for-bit with the oracle applied to the promoted values (v mod r)."""
import numpy as np
import pytest

import jolt_b200
from jolt_b200 import G1Bases, Polynomial, g1_jacobian_to_affine, HIGH_TO_LOW, LOW_TO_HIGH
from oracle import bn254 as O
from oracle import coracle as C

pytestmark = pytest.mark.gpu

G = np.array(O.to_mont_limbs(1, O.Q_MOD) + O.to_mont_limbs(2, O.Q_MOD), dtype=np.uint64)
KINDS = ["u8", "u16", "u32", "u64", "i64", "u128", "i128", "s64", "s128"]
RANGE = {"u8": (0, 1 << 8), "u16": (0, 1 << 16), "u32": (0, 1 << 32), "u64": (0, 1 << 64),
         "i64": (-(1 << 63), 1 << 63), "u128": (0, 1 << 128), "i128": (-(1 << 127), 1 << 127),
         # sign-magnitude (crates/jolt-field/src/signed.rs:25-32): the full unsigned magnitude with either sign
         "s64": (-(1 << 64) + 1, 1 << 64), "s128": (-(1 << 128) + 1, 1 << 128)}
DTYPE = {"u8": np.uint8, "u16": np.uint16, "u32": np.uint32, "u64": np.uint64, "i64": np.int64}


@pytest.fixture(scope="module")
def sess():
    s = jolt_b200.Session(0)
    yield s
    s.close()


@pytest.fixture(scope="module")
def srs_bases():
    beta = C.ints_to_mont([O.random_fr(0x4D534D, 1)[0]])[0]
    return C.g1_powers(1 << 13, G, beta)


def column(kind: str, n: int, seed: int) -> list[int]:
    """n values of the kind's range: uniform, with the extremes and small values mixed in."""
    lo, hi = RANGE[kind]
    rng = np.random.default_rng(seed)
    raw = [int.from_bytes(rng.bytes(16), "little") for _ in range(n)]
    vals = [lo + (x % (hi - lo)) for x in raw]
    edge = [lo, hi - 1, 0, 1, -1 if lo < 0 else 2, lo + 1, hi - 2]
    for i, e in enumerate(edge):
        if i < n:
            vals[(7 * i + 3) % n] = e
    return vals


def as_input(kind: str, vals: list[int]):
    return (np.array(vals, dtype=DTYPE[kind]), None) if kind in DTYPE else (vals, kind)


def oracle_affine(xy, inf):
    return None if inf else (O.from_mont_limbs(xy[:4], O.Q_MOD), O.from_mont_limbs(xy[4:], O.Q_MOD))


@pytest.mark.parametrize("kind", KINDS)
def test_promotion_matches_from_primitive(sess, kind):
    vals = column(kind, 1 << 10, 11)
    a, k = as_input(kind, vals)
    p = Polynomial.from_small(sess, a, k)
    assert p.to_ints() == [v % O.R_MOD for v in vals]
    p.free()


def test_bool_column(sess):
    bits = np.random.default_rng(5).integers(0, 2, size=256).astype(np.bool_)
    p = Polynomial.from_small(sess, bits)
    assert p.to_ints() == [int(b) for b in bits]


@pytest.mark.parametrize("order", [HIGH_TO_LOW, LOW_TO_HIGH])
@pytest.mark.parametrize("kind", KINDS)
def test_bind_to_field_matches_promote_then_bind(sess, kind, order):
    # dense.rs:613-625 bind_matches_bind_to_field
    n = 1 << 11
    vals = column(kind, n, 23)
    a, k = as_input(kind, vals)
    # a full 254-bit scalar and a 125-bit challenge (Montgomery limbs [0, 0, lo, hi])
    chal = ((0x0FEDCBA987654321 << 128) | (0x0234567898765432 << 192)) * O.inv_mod(O.MONT_R, O.R_MOD) % O.R_MOD
    for r in (O.random_fr(77, 1)[0], chal):
        got = Polynomial.bind_to_field(sess, a, r, order, k)
        want = O.bind([v % O.R_MOD for v in vals], r, order)
        assert got.to_ints() == want
        # and == the device's own promote-then-bind
        q = Polynomial.from_small(sess, a, k)
        q.bind_with_order(r, order)
        assert np.array_equal(q.evals(), got.evals())
        got.free()
        q.free()


def test_compact_errors(sess):
    with pytest.raises(ValueError, match="power of 2"):
        Polynomial.from_small(sess, np.arange(3, dtype=np.uint8))
    with pytest.raises(jolt_b200.JoltB200Error, match="power of two"):
        Polynomial.bind_to_field(sess, np.arange(3, dtype=np.uint8), 5)
    with pytest.raises(ValueError, match="unsupported compact dtype"):
        Polynomial.from_small(sess, np.zeros(4, dtype=np.float32))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [1, 5, 300, 1 << 12])
def test_msm_small_matches_field_msm(sess, srs_bases, kind, n):
    vals = column(kind, n, 31 + n)
    a, k = as_input(kind, vals)
    bases = G1Bases.from_affine(sess, srs_bases[:n])
    sc = C.ints_to_mont([v % O.R_MOD for v in vals])
    want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:n], sc, 0, C.max_threads()))
    got = g1_jacobian_to_affine(bases.msm_small(a, kind=k))
    assert got == want
    assert got == g1_jacobian_to_affine(bases.msm(sc))   # the device's own field-scalar path
    bases.free()


def test_msm_small_skewed_columns(sess, srs_bases):
    # what the witness really looks like: all zero, binary, one-hot, a constant column (every point in one bucket)
    n = 1 << 12
    bases = G1Bases.from_affine(sess, srs_bases[:n])
    rng = np.random.default_rng(9)
    cols = {
        "zero": np.zeros(n, dtype=np.uint8),
        "binary": rng.integers(0, 2, size=n).astype(np.uint8),
        "bool": rng.integers(0, 2, size=n).astype(np.bool_),
        "one_hot": np.eye(1, n, 1234, dtype=np.uint64).reshape(-1) * np.uint64(0xFFFF_FFFF_FFFF_FFFF),
        "constant": np.full(n, 200, dtype=np.uint16),
        "minus_one": np.full(n, -1, dtype=np.int64),
        "small_mixed": rng.integers(-3, 4, size=n).astype(np.int64),
    }
    for name, col in cols.items():
        vals = [int(v) for v in col]
        sc = C.ints_to_mont([v % O.R_MOD for v in vals])
        want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:n], sc, 0, C.max_threads()))
        assert g1_jacobian_to_affine(bases.msm_small(col)) == want, name
    bases.free()


def test_msm_small_with_precomputed_srs_and_offsets(sess, srs_bases):
    n = 1 << 12
    bases = G1Bases.from_affine(sess, srs_bases[:n]).precompute(8)
    vals = column("i64", n, 3)
    sc = C.ints_to_mont([v % O.R_MOD for v in vals])
    want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:n], sc, 0, C.max_threads()))
    assert g1_jacobian_to_affine(bases.msm_small(np.array(vals, dtype=np.int64))) == want
    half = n // 2
    want2 = oracle_affine(*C.g1_msm_pippenger(srs_bases[half:n], sc[:half], 0, C.max_threads()))
    assert g1_jacobian_to_affine(bases.msm_small(np.array(vals[:half], dtype=np.int64), offset=half)) == want2
    assert g1_jacobian_to_affine(bases.msm_small(np.zeros(0, dtype=np.uint32))) is None
    with pytest.raises(jolt_b200.JoltB200Error, match="length mismatch"):
        bases.msm_small(np.zeros(n + 1, dtype=np.uint8))
    bases.free()
    # n > 4096 on a precomputed handle: the shared-bucket path (one bucket set, 9-bit windows) with few windows
    n = 1 << 13
    big = G1Bases.from_affine(sess, srs_bases[:n]).precompute(9)
    for kind in ("u16", "i64", "i128"):
        vals = column(kind, n, 41)
        a, k = as_input(kind, vals)
        sc = C.ints_to_mont([v % O.R_MOD for v in vals])
        want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:n], sc, 0, C.max_threads()))
        assert g1_jacobian_to_affine(big.msm_small(a, kind=k)) == want, kind
    big.free()


def test_sign_magnitude_negative_zero_and_extremes(sess, srs_bases):
    # "Zero is not canonicalized: a zero magnitude may carry either sign" (crates/jolt-field/src/signed.rs:16-17;
    # limbs_signed_differential.rs:239-240 builds both): -0 promotes to 0 and contributes nothing to an MSM
    n = 8
    m64, m128 = (1 << 64) - 1, (1 << 128) - 1
    for kind, top in (("s64", m64), ("s128", m128)):
        recs = [(0, False), (0, True), (1, False), (1, True), (top, False), (top, True), (5, False), (7, True)]
        vals = [m if pos else -m for m, pos in recs]
        p = Polynomial.from_small(sess, recs, kind)
        assert p.to_ints() == [v % O.R_MOD for v in vals], kind
        p.free()
        bases = G1Bases.from_affine(sess, srs_bases[:n])
        sc = C.ints_to_mont([v % O.R_MOD for v in vals])
        want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:n], sc, 0, 1))
        assert g1_jacobian_to_affine(bases.msm_small(recs, kind=kind)) == want, kind
        bases.free()


def test_batch_msm_matches_single_msms(sess, srs_bases):
    # batch_msm: every column against the prefix bases[..len] (msm/mod.rs:160-168); batch_msm_univariate: field columns (:170-181)
    n = 1 << 11
    bases = G1Bases.from_affine(sess, srs_bases[:n])
    rng = np.random.default_rng(21)
    fr_col = C.ints_to_mont(O.random_fr(5, 700))
    cols = [
        rng.integers(0, 2, size=n).astype(np.uint8),          # binary -> msm_binary arm
        np.zeros(300, dtype=np.uint8),                        # all zero -> identity
        rng.integers(0, 1 << 16, size=1000).astype(np.uint16),
        rng.integers(-(1 << 62), 1 << 62, size=n).astype(np.int64),
        (column("s128", 513, 4), "s128"),
        (column("i128", 64, 6), "i128"),
        fr_col,                                               # LargeScalars / UniPoly coefficients
        np.zeros(0, dtype=np.uint32),                         # empty column -> identity
    ]
    got = bases.batch_msm(cols)
    assert got.shape == (len(cols), 12)
    for j, col in enumerate(cols):
        if isinstance(col, tuple):
            vals = [int(v) for v in col[0]]
        elif col.dtype == np.uint64 and col.ndim == 2:
            vals = None
        else:
            vals = [int(v) for v in col]
        sc = col if vals is None else C.ints_to_mont([v % O.R_MOD for v in vals])
        m = len(sc)
        want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:m], sc, 0, C.max_threads())) if m else None
        assert g1_jacobian_to_affine(got[j]) == want, j
    with pytest.raises(jolt_b200.JoltB200Error, match="longer than the base set"):
        bases.batch_msm([np.zeros(n + 1, dtype=np.uint8)])
    bases.free()


@pytest.mark.parametrize("rows,width", [(2, 64), (5, 256), (37, 1 << 10), (3, 1 << 13)])
def test_msm_rows_matches_row_msms(sess, srs_bases, rows, width):
    # Dory tier-1: one MSM per matrix row against the same bases (crates/jolt-dory/src/streaming.rs:113-201)
    bases = G1Bases.from_affine(sess, srs_bases[:width])
    kinds = ["u64", "i128", "s64", "u8"] if width <= 1 << 10 else ["i64"]
    for kind in kinds:
        vals = column(kind, rows * width, 100 + rows)
        if kind == "u8":   # one binary row, one zero row, one constant row (every point of the row in one bucket)
            vals[:width] = [v & 1 for v in vals[:width]]
            vals[width:2 * width] = [0] * width
            if rows > 2:
                vals[2 * width:3 * width] = [200] * width
        a, k = as_input(kind, vals)
        got = bases.msm_rows(a, rows, kind=k)
        for r in range(rows):
            sc = C.ints_to_mont([v % O.R_MOD for v in vals[r * width:(r + 1) * width]])
            want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:width], sc, 0, C.max_threads()))
            assert g1_jacobian_to_affine(got[r]) == want, (kind, r)
    if width <= 256:   # field rows (DoryScheme::feed)
        sc = C.ints_to_mont(O.random_fr(rows, rows * width))
        got = bases.msm_rows(sc, rows, kind="fr")
        for r in range(rows):
            want = oracle_affine(*C.g1_msm_pippenger(srs_bases[:width], sc[r * width:(r + 1) * width], 0, C.max_threads()))
            assert g1_jacobian_to_affine(got[r]) == want, ("fr", r)
    with pytest.raises(jolt_b200.JoltB200Error, match="length mismatch"):
        bases.msm_rows(np.zeros(2 * (width + 1), dtype=np.uint8), 2)
    bases.free()


def test_msm_binary_fast_path(sess, srs_bases):
    # from 2^14 terms a u8 / bool column is scanned once (the reference's par_iter().all(), msm/mod.rs:35-47, 96-106):
    # all zero -> identity, all <= 1 -> msm_binary (the select-sum kernel), otherwise msm_u8
    n = (1 << 14) + 37
    reps = -(-n // len(srs_bases))
    xy = np.concatenate([srs_bases] * reps)[:n].copy()     # repeated bases: P + P inside one thread's running sum
    xy[5] = 0                                              # an identity base that is selected
    xy[n - 1] = 0
    bases = G1Bases.from_affine(sess, xy)
    rng = np.random.default_rng(77)
    for name, col in (("random", rng.integers(0, 2, size=n).astype(np.uint8)), ("bool", rng.integers(0, 2, size=n).astype(np.bool_)),
                      ("all_ones", np.ones(n, dtype=np.uint8)), ("one_hot", np.eye(1, n, n - 3, dtype=np.uint8).reshape(-1)),
                      ("zero", np.zeros(n, dtype=np.uint8)), ("not_binary", (rng.integers(0, 2, size=n) * 2).astype(np.uint8))):
        vals = [int(v) for v in col]
        vals[5] = int(col[5])
        sc = C.ints_to_mont(vals)
        want = oracle_affine(*C.g1_msm_pippenger(xy, sc, 0, C.max_threads()))
        assert g1_jacobian_to_affine(bases.msm_small(col)) == want, name
    bases.free()


@pytest.mark.parametrize("kind", ["u64", "i64", "u128", "i128", "s64", "s128"])
def test_scalar_mul_fast_paths_match(sess, kind):
    # crates/jolt-field/tests/bn254_differential.rs:126-152 (scalar_mul_fast_paths_match): t.mul_u64(s), mul_i64, mul_u128,
    # mul_i128 == t * s mod r. The reference's Barrett fast paths exist to skip a Montgomery pass on the CPU; on the device
    # the product of a field element by a primitive integer is the promotion (jb_table_upload_small) followed by the
    # ordinary product - same VALUE, which is what the differential test pins (edges 0, 1, 2 and the type's extremes included)
    n = 1 << 9
    t = O.random_fr(0x7157 + len(kind), n)
    s = column(kind, n, 17)
    a, k = as_input(kind, s)
    t_limbs = Polynomial.from_ints(sess, t).evals()
    s_limbs = Polynomial.from_small(sess, a, k).evals()
    from jolt_b200 import field as F
    got = F.limbs_to_ints(sess.vec_op(0, 2, t_limbs, s_limbs))
    assert got == [(x * y) % O.R_MOD for x, y in zip(t, s)]
