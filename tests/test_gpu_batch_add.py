"""Batch affine G1 addition on the device (jb_g1_batch_add) vs the oracle restatement of
crates/jolt-crypto/src/ec/bn254/batch_addition.rs:53-150 and vs plain group addition; mirrors the reference's own
tests (:160-240: empty / singleton sets, random sets of unique indices against projective sums)."""
import numpy as np
import pytest

import jolt_b200
from jolt_b200 import G1Bases, g1_affine_limbs, g1_jacobian_to_affine
from jolt_b200 import field as F
from oracle import bn254 as O
from gpu_util import rand_limbs

pytestmark = pytest.mark.gpu

G = np.array(O.to_mont_limbs(1, O.Q_MOD) + O.to_mont_limbs(2, O.Q_MOD), dtype=np.uint64)


@pytest.fixture(scope="module")
def sess():
    s = jolt_b200.Session(0)
    yield s
    s.close()


def affine_points(xy_limbs):
    out = []
    for row in np.asarray(xy_limbs).reshape(-1, 8):
        x, y = F.from_limbs(row[:4], F.Q_MOD), F.from_limbs(row[4:], F.Q_MOD)
        out.append(None if x == 0 and y == 0 else (x, y))
    return out


def test_empty_and_singleton_sets(sess):
    # batch_addition.rs:167-178
    bases = G1Bases.generate_multiples(sess, G, 4)
    pts = affine_points(bases.affine())
    assert bases.batch_add([]).shape[0] == 0
    got = affine_points(bases.batch_add([[], [2], [0, 3]]))
    assert got[0] is None and got[1] == pts[2] and got[2] == O.g1_add(pts[0], pts[3])


def test_random_sets_match_restatement_and_group_sums(sess):
    # batch_addition.rs:181-: 1000 bases, 10 sets of 1..50 unique indices
    n = 1000
    # random-looking bases beta^i G (structured multiples (i + 1) G would let partial sums collide: 2G + 3G meets 5G)
    from oracle import coracle as C
    bases = G1Bases.from_affine(sess, C.g1_powers(n, G, C.ints_to_mont([O.random_fr(0x4D534D, 1)[0]])[0]))
    pts = affine_points(bases.affine())
    rng = np.random.Generator(np.random.PCG64(7))
    sets = [list(rng.choice(n, size=int(rng.integers(1, 51)), replace=False)) for _ in range(10)]
    sets += [[], [5], list(range(0, 64)), list(range(100, 133))]
    got = affine_points(bases.batch_add(sets))
    assert got == O.batch_g1_additions_multi_affine(pts, sets)
    for g, st in zip(got, sets):
        acc = None
        for i in st:
            acc = O.g1_add(acc, pts[i])
        assert g == acc


def test_large_one_hot_like_sets(sess):
    """A binary / one-hot column: a few very large sets (many levels, several blocks per level). The sum of the
    bases (i + 1) G over an index set is (sum (i + 1)) G."""
    n = 1 << 16
    bases = G1Bases.generate_multiples(sess, G, n)
    rng = np.random.Generator(np.random.PCG64(9))
    perm = rng.permutation(n)
    sets = [perm[:40000], perm[40000:40001], perm[40001:65000], np.arange(0, n, 2)]
    got = affine_points(bases.batch_add(sets))
    for g, st in zip(got, sets):
        k = int(sum(int(i) + 1 for i in st)) % O.R_MOD
        assert g == O.g1_scalar_mul((1, 2), k)
    # the same sums through the small-scalar MSM (msm_binary) agree
    col = np.zeros(n, dtype=np.uint8)
    col[sets[0]] = 1
    assert g1_jacobian_to_affine(bases.msm_small(col)) == got[0]


def test_equal_x_pair_is_the_references_unchecked_garbage_for_that_pair_only(sess):
    """The distinct-x precondition is not checked by the reference (batch_addition.rs:44-49): ark's batch inversion
    leaves the zero denominator zero, so the pair's lambda is 0. The device must reproduce exactly that value and
    keep every other set of the same batch correct."""
    n = 64
    bases = G1Bases.generate_multiples(sess, G, n)
    pts = affine_points(bases.affine())
    sets = [[3, 3], [1, 2, 5], [7, 7, 9, 10], [5]]
    got = affine_points(bases.batch_add(sets))
    want = O.batch_g1_additions_multi_affine(pts, sets)
    assert got == want
    assert got[1] == O.g1_add(O.g1_add(pts[1], pts[2]), pts[5]) and got[3] == pts[5]
    assert not O.g1_is_on_curve(got[0])


def test_index_out_of_bounds_is_an_error(sess):
    bases = G1Bases.generate_multiples(sess, G, 8)
    with pytest.raises(jolt_b200.JoltB200Error):
        bases.batch_add([[1, 9]])
