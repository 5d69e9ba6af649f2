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
