"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np

from oracle import bn254 as O
from oracle import coracle as C



from oracle.coracle import rand_challenge, rand_limbs  # noqa: E402,F401


def rand_full(seed: int) -> np.ndarray:
    return np.array(O.to_mont_limbs(O.random_fr(seed, 1)[0]), dtype=np.uint64)


EDGE_INTS = [0, 1, 2, O.R_MOD - 1, O.R_MOD - 2, (1 << 256) % O.R_MOD, (1 << 64) - 1, (1 << 128) - 1, (1 << 253) + 5]
