"""Pins the oracle against the reference's own golden vectors
(crates/jolt-field/tests/golden_bytes.rs:68-329, committed as tests/golden/field_golden.json)."""
import json
import pathlib

import pytest

from oracle import bn254 as O

GOLD = json.loads((pathlib.Path(__file__).parent / "golden" / "field_golden.json").read_text())


@pytest.mark.parametrize("name,p", [("FIX_BN254_FR", O.R_MOD), ("FIX_BN254_FQ", O.Q_MOD)])
def test_prime_rows(name, p):
    for inp, exp in GOLD[name]:
        v = O.from_bytes_le_reduced(bytes.fromhex(inp), p)
        assert O.to_bytes_le(v).hex() == exp
        # checked decode round-trip (golden_bytes.rs:56-60)
        assert int.from_bytes(bytes.fromhex(exp), "little") == v < p


def test_fr_challenge_rows():
    for inp, exp in GOLD["FIX_BN254_FR_CHALLENGE"]:
        v = O.fr_from_challenge_bytes(bytes.fromhex(inp))
        assert O.to_bytes_le(v).hex() == exp
        # the Montgomery limbs of that value ARE [0, 0, low, high] (mod.rs:254)
        low, high = O.challenge_limbs(bytes.fromhex(inp))
        assert O.to_mont_limbs(v) == [0, 0, low, high]


def test_fq_challenge_rows():
    for inp, exp in GOLD["FIX_BN254_FQ_CHALLENGE"]:
        assert O.to_bytes_le(O.fq_from_challenge_bytes(bytes.fromhex(inp))).hex() == exp


@pytest.mark.parametrize("name,p", [("FIX_BN254_FR_SCALAR_CHALLENGE", O.R_MOD),
                                    ("FIX_BN254_FQ_SCALAR_CHALLENGE", O.Q_MOD)])
def test_scalar_challenge_rows(name, p):
    for inp, exp in GOLD[name]:
        assert O.to_bytes_le(O.from_scalar_challenge_bytes(bytes.fromhex(inp), p)).hex() == exp


def test_mont_constants():
    # derived constants
    assert (-pow(O.R_MOD, -1, 1 << 64)) % (1 << 64) == 0xC2E1F593EFFFFFFF
    assert (-pow(O.Q_MOD, -1, 1 << 64)) % (1 << 64) == 0x87D20782E4866389
    assert O.mont_raw(O.to_mont_limbs(1)) == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert O.from_mont_limbs(O.to_mont_limbs(12345)) == 12345


def test_small_scalar_msm_facade_restatement():
    """oracle.msm_small / batch_msm / msm_rows (crates/jolt-prover-legacy/src/msm/mod.rs:27-181): the three u8 arms agree,
    negatives and sign-magnitude scalars promote as +-|v| mod r (the reference's -0 included), columns use base prefixes."""
    from oracle import bn254 as O
    pts = [O.g1_scalar_mul(O.G1_GEN, k) for k in (1, 2, 3, 5, 7, 11)]
    mult = [1, 2, 3, 5, 7, 11]
    def closed(vals):
        t = sum(m * v for m, v in zip(mult, vals)) % O.R_MOD
        return O.g1_scalar_mul(O.G1_GEN, t) if t else None
    assert O.msm_small(pts, [0] * 6, "u8") is None
    bits = [1, 0, 1, 1, 0, 1]
    assert O.msm_small(pts, bits, "u8") == closed(bits) == O.g1_msm_naive(pts, bits)
    u8 = [200, 0, 1, 255, 3, 9]
    assert O.msm_small(pts, u8, "u8") == closed(u8)
    i64 = [-1, 5, -(1 << 63), (1 << 63) - 1, 0, -7]
    assert O.msm_small(pts, i64, "i64") == closed(i64)
    s64 = [(0, False), (0, True), ((1 << 64) - 1, False), (5, True), (1, False), (9, True)]
    assert O.signed_to_fr(0, False) == 0 and O.signed_to_fr(5, False) == O.R_MOD - 5
    assert O.msm_small(pts, s64, "s64") == closed([m if pos else -m for m, pos in s64])
    cols = [(bits[:4], "u8"), (i64, "i64"), ([], "u16")]
    got = O.batch_msm(pts, cols)
    assert got[0] == O.g1_msm_naive(pts[:4], bits[:4]) and got[1] == closed(i64) and got[2] is None
    rows = O.msm_rows(pts, [1, 2, 3, 4, 5, 6, 7, 8, 9], 3, "u16")
    assert rows == [O.g1_msm_naive(pts[:3], [1, 2, 3]), O.g1_msm_naive(pts[:3], [4, 5, 6]), O.g1_msm_naive(pts[:3], [7, 8, 9])]
