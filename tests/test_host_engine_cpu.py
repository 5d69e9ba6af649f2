"""The C++ host engine (jolt_b200/csrc/sumcheck_host.cu: prove_batch, SequentialRounds, UnivariatePoly::from_evals,
the SplitMix stand-in transcript) checked against the oracle's prove_batch WITHOUT a GPU: the same source file is
compiled with g++ next to a harness whose members keep their tables on the host (tests/native/host_engine_harness.cpp).
Mirrors crates/jolt-sumcheck/src/tests.rs (batched members with different round counts, offsets and coefficients)."""
import pathlib
import shutil
import subprocess

import numpy as np
import pytest

from jolt_b200 import UnivariatePoly
from jolt_b200 import field as F
from jolt_b200.dist import splitmix_challenge
from oracle import bn254 as O

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("he") / "host_engine_harness"
    cmd = [gxx, "-O1", "-std=c++17", "-x", "c++", str(ROOT / "jolt_b200/csrc/sumcheck_host.cu"),
           "-x", "c++", str(ROOT / "tests/native/host_engine_harness.cpp"), "-o", str(out)]
    subprocess.run(cmd, check=True, capture_output=True, timeout=300)
    return out


def hexlimbs(v: int) -> str:
    return " ".join(f"{int(x):016x}" for x in F.to_limbs(v))


def parse_fr(tokens) -> int:
    return F.from_limbs(np.array([int(t, 16) for t in tokens], dtype=np.uint64))


def run_case(harness, members, max_num_vars, max_degree, seed):
    """members: list of dict(tables=[[int]], order, rounds, offset, coefficient)"""
    lines = [f"{len(members)} {max_num_vars} {max_degree} {seed}"]
    for mb in members:
        m, n = len(mb["tables"]), len(mb["tables"][0]).bit_length() - 1
        lines.append(f"{m} {n} {mb['order']} {mb['rounds']} {mb['offset']} {hexlimbs(mb['coefficient'])}")
        for t in mb["tables"]:
            lines.extend(hexlimbs(v) for v in t)
    res = subprocess.run([str(harness)], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=120, check=True)
    got = {"challenges": [], "round_polys": [], "member_claims": [], "bound": []}
    for line in res.stdout.splitlines():
        tag, *rest = line.split()
        if tag == "error":
            got["error"] = " ".join(rest)
        elif tag == "challenge":
            got["challenges"].append(parse_fr(rest))
        elif tag == "poly":
            got["round_polys"].append([])
        elif tag == "coeff":
            got["round_polys"][-1].append(parse_fr(rest))
        elif tag == "final":
            got["final_claim"] = parse_fr(rest)
        elif tag == "claim":
            got["member_claims"].append(parse_fr(rest))
        elif tag == "bound":
            got["bound"].append(parse_fr(rest))
        elif tag == "sum":
            got["sum"] = parse_fr(rest)
    return got


def oracle_case(members, max_num_vars, max_degree, seed):
    desc, objs = [], []
    total = 0
    for mb in members:
        n = len(mb["tables"][0])
        claim = sum(int(np.prod([t[x] for t in mb["tables"]], dtype=object)) for x in range(n)) % O.R_MOD
        desc.append(dict(input_claim=claim, coefficient=mb["coefficient"], rounds=mb["rounds"], offset=mb["offset"]))
        objs.append(O.ProductMember(mb["tables"], mb["order"]))
        total = (total + mb["coefficient"] * claim * pow(2, max_num_vars - mb["rounds"], O.R_MOD)) % O.R_MOD
    want = O.prove_batch(desc, objs, max_num_vars, max_degree, total,
                         lambda r, coeffs: splitmix_challenge(seed, UnivariatePoly(list(coeffs))))
    want["sum"] = total
    want["bound"] = [v for o in objs for v in o.final_evals()]
    return want


# members are tail-aligned (offset + rounds == max_num_vars): "a tail-aligned member reaches its true input claim exactly
# when it activates" (prover.rs:241-245); before that it contributes the constant claim/2 per round
CASES = {
    "one_member_deg1": ([dict(tables=[O.random_fr(1, 16)], order=O.HIGH_TO_LOW, rounds=4, offset=0, coefficient=1)], 4, 1),
    "one_member_deg2_l2h": ([dict(tables=[O.random_fr(2, 32), O.random_fr(3, 32)], order=O.LOW_TO_HIGH, rounds=5, offset=0,
                                  coefficient=O.random_fr(4, 1)[0])], 5, 2),
    "three_members_windows": ([
        dict(tables=[O.random_fr(5, 64), O.random_fr(6, 64), O.random_fr(7, 64)], order=O.HIGH_TO_LOW, rounds=6, offset=0,
             coefficient=O.random_fr(8, 1)[0]),
        dict(tables=[O.random_fr(9, 8), O.random_fr(10, 8)], order=O.LOW_TO_HIGH, rounds=3, offset=3, coefficient=O.random_fr(11, 1)[0]),
        dict(tables=[O.dense_member_with_sum(4, 90210, 41)], order=O.HIGH_TO_LOW, rounds=4, offset=2, coefficient=O.R_MOD - 1),
    ], 6, 3),
    "late_member_degree_trim": ([
        dict(tables=[[7] * 8], order=O.HIGH_TO_LOW, rounds=3, offset=0, coefficient=1),          # constant rounds: trimmed to length 2
        dict(tables=[O.random_fr(12, 2), O.random_fr(13, 2)], order=O.HIGH_TO_LOW, rounds=1, offset=2, coefficient=5),
    ], 3, 2),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_cpp_engine_matches_oracle_engine(harness, name):
    members, max_num_vars, max_degree = CASES[name]
    got = run_case(harness, members, max_num_vars, max_degree, seed=11)
    want = oracle_case(members, max_num_vars, max_degree, seed=11)
    assert "error" not in got, got.get("error")
    for key in ("sum", "challenges", "round_polys", "final_claim", "member_claims", "bound"):
        assert got[key] == want[key], key


def test_cpp_engine_rejects_bad_batches(harness):
    tabs = [O.random_fr(1, 8)]
    # member window outside the batch (prover.rs: BatchMemberWindowOutOfRange)
    got = run_case(harness, [dict(tables=tabs, order=O.HIGH_TO_LOW, rounds=3, offset=2, coefficient=1)], 4, 1, 1)
    assert "BatchMemberWindowOutOfRange" in got.get("error", "")
    # a degree-2 member in a degree-1 batch (DegreeBoundExceeded)
    got = run_case(harness, [dict(tables=[O.random_fr(1, 8), O.random_fr(2, 8)], order=O.HIGH_TO_LOW, rounds=3, offset=0, coefficient=1)],
                   3, 1, 1)
    assert "DegreeBoundExceeded" in got.get("error", "")


def test_host_round_cost_is_small(harness):
    """The engine's own work per round sits on the Fiat-Shamir round trip; keep it in the microsecond range."""
    import json
    res = subprocess.run([str(harness), "bench", "22"], capture_output=True, text=True, timeout=120, check=True)
    ns = json.loads(res.stdout)["host_ns_per_round"]
    assert ns < 50_000, ns          # generous: shared CI cores; typically ~1 us


def test_sign_magnitude_records_layout():
    """jb_s64 / jb_s128 (include/jolt_b200.h): N u64 magnitude limbs, then the sign byte (is_positive), padded to 8 -
    the #[repr(C)] mirror of jolt_field::signed::SignedBigInt<N> (crates/jolt-field/src/signed.rs:25-32)."""
    import numpy as np
    from jolt_b200.api import SCALAR_KINDS, small_scalars
    a, k, n = small_scalars([5, -7, (0, False), ((1 << 64) - 1, False)], "s64")
    assert (k, n) == (SCALAR_KINDS["s64"], 4) and a.dtype == np.uint64 and a.shape == (4, 2) and a.nbytes == 4 * 16
    raw = a.tobytes()
    assert raw[0:8] == (5).to_bytes(8, "little") and raw[8] == 1 and raw[9:16] == bytes(7)
    assert raw[16:24] == (7).to_bytes(8, "little") and raw[24] == 0          # -7: magnitude 7, is_positive = false
    assert raw[32:40] == bytes(8) and raw[40] == 0                           # the reference's -0
    assert raw[48:56] == b"\xff" * 8 and raw[56] == 0
    b, k, n = small_scalars([-(1 << 100) - 3, ((1 << 128) - 1, True)], "s128")
    assert (k, n) == (SCALAR_KINDS["s128"], 2) and b.shape == (2, 3) and b.nbytes == 2 * 24
    raw = b.tobytes()
    assert int.from_bytes(raw[0:16], "little") == (1 << 100) + 3 and raw[16] == 0 and raw[17:24] == bytes(7)
    assert raw[24:40] == b"\xff" * 16 and raw[40] == 1
    import pytest
    with pytest.raises(ValueError):
        small_scalars([1 << 64], "s64")
