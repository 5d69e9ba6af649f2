"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np

from oracle import bn254 as O
from oracle import coracle as C

_MASK = (1 << 64) - 1


def rand_limbs(seed: int, n: int, p: int = O.R_MOD) -> np.ndarray:
    """n canonical Montgomery elements from numpy's PCG64 (fast path for big tables): draw 256
    bits, clear the top 3 so the raw value < 2^253 < p - every limb pattern below 2^253 is a
    valid canonical Montgomery representative."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(_MASK >> 3)
    return a


def rand_challenge(seed: int) -> np.ndarray:
    """125-bit challenge as raw Montgomery limbs [0,0,lo,hi] (mod.rs:172-184)."""
    st, lo = O.splitmix64(seed)
    st, hi = O.splitmix64(st)
    return np.array(O.challenge_to_mont_limbs(lo, hi), dtype=np.uint64)


def rand_full(seed: int) -> np.ndarray:
    return np.array(O.to_mont_limbs(O.random_fr(seed, 1)[0]), dtype=np.uint64)


EDGE_INTS = [0, 1, 2, O.R_MOD - 1, O.R_MOD - 2, (1 << 256) % O.R_MOD, (1 << 64) - 1, (1 << 128) - 1, (1 << 253) + 5]
