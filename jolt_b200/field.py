"""Host-side BN254 Fr helpers for the O(rounds * degree) glue the reference keeps on the host
(crates/jolt-sumcheck/src/prover.rs:246-343, crates/jolt-poly/src/univariate.rs). Values are
Python ints mod r; the device ABI speaks 4 x u64 Montgomery limbs
(crates/jolt-field/src/bn254/mod.rs:33-42)."""
from __future__ import annotations

import numpy as np

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_MASK = (1 << 64) - 1
_MONT = 1 << 256
_RINV = {R_MOD: pow(_MONT, -1, R_MOD), Q_MOD: pow(_MONT, -1, Q_MOD)}


def to_limbs(v: int, p: int = R_MOD) -> np.ndarray:
    m = (v % p) * _MONT % p
    return np.array([(m >> (64 * i)) & _MASK for i in range(4)], dtype=np.uint64)


def from_limbs(limbs, p: int = R_MOD) -> int:
    raw = 0
    for i in range(4):
        raw |= int(limbs[i]) << (64 * i)
    return raw * _RINV[p] % p


def ints_to_limbs(vals, p: int = R_MOD) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = to_limbs(v, p)
    return out


def limbs_to_ints(arr, p: int = R_MOD) -> list[int]:
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [from_limbs(a[i], p) for i in range(a.shape[0])]


def challenge_limbs(low: int, high: int) -> np.ndarray:
    """The 125-bit sumcheck challenge as raw Montgomery limbs [0, 0, low, high]
    (from_challenge_bytes, crates/jolt-field/src/bn254/mod.rs:172-184, :254)."""
    return np.array([0, 0, low & _MASK, high & (_MASK >> 3)], dtype=np.uint64)


def challenge_from_bytes(b: bytes) -> np.ndarray:
    buf = bytes(b[:16]).ljust(16, b"\0")
    v = int.from_bytes(buf, "little")
    return challenge_limbs(v & _MASK, v >> 64)
