"""ORACLE (test infrastructure, NOT product code).

Plain Python big-int restatement of the a16z/jolt hot path: BN254 Fr/Fq
encodings, multilinear bind, eq tables, the univariate-evaluation sweep, the
batched sumcheck engine, BN254 G1 arithmetic, MSM and the HyperKZG prover side.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module. Every function cites the reference file:line it restates
(paths relative to /root/reference).

Parity status
-------------
* Fr/Fq encode / reduce / challenge decode: PINNED by the reference's golden
  vectors (crates/jolt-field/tests/golden_bytes.rs:68-329), reproduced verbatim
  in tests/golden/field_golden.json and checked by tests/test_oracle_golden.py.
* Fr arithmetic == arithmetic mod r: the reference pins this differentially
  against num-bigint (crates/jolt-field/tests/bn254_differential.rs:75-99), so a
  Python int mod r is authoritative; Montgomery limbs are (v * 2^256 mod r) LE.
* bind / eq / sweep / engine: property-pinned only in the reference (no stored
  outputs); exactness follows from field exactness + the cited definitions.
* G1 / MSM / HyperKZG: "parity unpinned" by golden vectors - the reference holds
  no serialized G1 point anywhere. Anchored on the published curve constants
  (y^2 = x^3 + 3 over Fq, generator (1, 2), order r) and on the reference's own
  property tests (crates/jolt-crypto/tests/group_laws.rs:12-146).

The arithmetic itself (Montgomery mul, G1 formulas, Pippenger) lives in the
third-party arkworks fork a16z/arkworks-algebra @ 76bb3a4518928f1ff7f15875f940d614bb9845e6
(Cargo.lock:883-885), absent from /root/reference; results are defined by value
(field element / group element), independent of schedule.
"""
from __future__ import annotations

# --------------------------------------------------------------------------- #
# Constants (crates/jolt-field/tests/bn254_differential.rs:21-36)
# --------------------------------------------------------------------------- #
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr modulus r
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq modulus q
MONT_R = 1 << 256
MASK64 = (1 << 64) - 1
CURVE_B = 3
G1_GEN = (1, 2)


def inv_mod(a: int, p: int) -> int:
    return pow(a, -1, p)


# --------------------------------------------------------------------------- #
# Representation: Montgomery limbs  (crates/jolt-field/src/bn254/mod.rs:33-42)
# --------------------------------------------------------------------------- #
def to_mont_limbs(v: int, p: int = R_MOD) -> list[int]:
    """Canonical value -> 4 x u64 little-endian Montgomery limbs (v * 2^256 mod p)."""
    m = (v % p) * MONT_R % p
    return [(m >> (64 * i)) & MASK64 for i in range(4)]


def from_mont_limbs(limbs, p: int = R_MOD) -> int:
    m = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    assert m < p, "non-canonical Montgomery limbs"
    return m * inv_mod(MONT_R, p) % p


def mont_raw(limbs) -> int:
    """The raw integer held in the limbs (no Montgomery decode)."""
    return sum(int(l) << (64 * i) for i, l in enumerate(limbs))


# --------------------------------------------------------------------------- #
# Byte encodings (crates/jolt-field/src/bn254/mod.rs:113-193)
# --------------------------------------------------------------------------- #
def from_bytes_le_reduced(b: bytes, p: int = R_MOD) -> int:
    return int.from_bytes(b, "little") % p


def to_bytes_le(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def challenge_limbs(b: bytes) -> tuple[int, int]:
    """16 bytes -> (low, high) with the top 3 bits of `high` cleared (mod.rs:172-181)."""
    buf = bytes(b[:16]).ljust(16, b"\0")
    value = int.from_bytes(buf, "little")
    low = value & MASK64
    high = (value >> 64) & (MASK64 >> 3)
    return low, high


def fr_from_challenge_bytes(b: bytes) -> int:
    """Fr: limbs [0,0,low,high] are taken AS Montgomery form (from_bigint_unchecked,
    mod.rs:254) -> value = (low*2^128 + high*2^192) * 2^-256 mod r."""
    low, high = challenge_limbs(b)
    raw = (low << 128) | (high << 192)
    return raw * inv_mod(MONT_R, R_MOD) % R_MOD


def fq_from_challenge_bytes(b: bytes) -> int:
    """Fq: checked from_bigint (mod.rs:262) -> the canonical value is the integer itself."""
    low, high = challenge_limbs(b)
    return ((low << 128) | (high << 192)) % Q_MOD


def from_scalar_challenge_bytes(b: bytes, p: int = R_MOD) -> int:
    """Digest bytes read big-endian then reduced (mod.rs:189-193)."""
    return int.from_bytes(b, "big") % p


def challenge_to_mont_limbs(low: int, high: int) -> list[int]:
    """Raw Montgomery limbs of a 125-bit sumcheck challenge: [0, 0, low, high]."""
    return [0, 0, low & MASK64, high & (MASK64 >> 3)]


# --------------------------------------------------------------------------- #
# Small-int conversions (crates/jolt-field/src/bn254/mod.rs:264-300, mont.rs:287-325)
# --------------------------------------------------------------------------- #
def fr_from_u64(v: int) -> int:
    return v % R_MOD


def fr_from_i64(v: int) -> int:
    return v % R_MOD


# --------------------------------------------------------------------------- #
# Multilinear bind (crates/jolt-poly/src/dense.rs:180-263;
# legacy crates/jolt-prover-legacy/src/poly/dense_mlpoly.rs:85-221)
# --------------------------------------------------------------------------- #
HIGH_TO_LOW = 0
LOW_TO_HIGH = 1


def bind(evals: list[int], s: int, order: int = HIGH_TO_LOW, p: int = R_MOD) -> list[int]:
    half = len(evals) // 2
    assert half >= 1, "cannot bind a zero-variable polynomial"
    if order == HIGH_TO_LOW:  # dense.rs:188-220
        return [(evals[i] + s * (evals[i + half] - evals[i])) % p for i in range(half)]
    # dense.rs:223-263
    return [(evals[2 * i] + s * (evals[2 * i + 1] - evals[2 * i])) % p for i in range(half)]


def evaluate(evals: list[int], point: list[int], p: int = R_MOD) -> int:
    """Polynomial::evaluate (dense.rs:339-360): sum_x f(x) * eq(x, point)."""
    table = eq_evals(point, None, p)
    return sum(a * b for a, b in zip(evals, table)) % p


# --------------------------------------------------------------------------- #
# Eq tables (crates/jolt-poly/src/eq.rs:50-98, 221-231, 299-315)
# --------------------------------------------------------------------------- #
def eq_evals(r: list[int], scale: int | None = None, p: int = R_MOD) -> list[int]:
    """EqPolynomial::evals - big-endian: r[0] is the MSB of the table index (eq.rs:299-315)."""
    evals = [1 % p if scale is None else scale % p] * (1 << len(r))
    size = 1
    for r_j in r:
        size *= 2
        for i in range(size - 1, 0, -2):
            scalar = evals[i // 2]
            evals[i] = scalar * r_j % p
            evals[i - 1] = (scalar - evals[i]) % p
    return evals


def eq_evaluations(point: list[int], p: int = R_MOD) -> list[int]:
    """EqPolynomial::evaluations (eq.rs:50-98) - same table via 2 muls per pair."""
    table = [1]
    for r_i in point:
        nxt = [0] * (2 * len(table))
        for j, base in enumerate(table):
            nxt[2 * j] = base * (1 - r_i) % p
            nxt[2 * j + 1] = base * r_i % p
        table = nxt
    return table


def eq_evals_for_aligned_block(r: list[int], start_index: int, block_size: int, p: int = R_MOD):
    """eq.rs:238-263."""
    assert block_size & (block_size - 1) == 0 and block_size and start_index % block_size == 0
    block_vars = block_size.bit_length() - 1
    prefix_len = len(r) - block_vars
    prefix_value = start_index >> block_vars
    scale = 1
    for i in range(prefix_len):
        bit = (prefix_value >> (prefix_len - 1 - i)) & 1
        scale = scale * (r[i] if bit else (1 - r[i])) % p
    return eq_evals(r[prefix_len:], scale, p)


# --------------------------------------------------------------------------- #
# Univariate polynomials (crates/jolt-poly/src/univariate.rs:58-69, 185-216, 470-487)
# --------------------------------------------------------------------------- #
def uni_evaluate(coeffs: list[int], x: int, p: int = R_MOD) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % p
    return acc


def uni_from_evals(evals: list[int], p: int = R_MOD) -> list[int]:
    """Interpolate from evaluations at 0..n-1 (Vandermonde solve, univariate.rs:198-202).
    Value-equal to any interpolation method; done here by Gauss-Jordan over the field."""
    n = len(evals)
    mat = [[pow(x, k, p) for k in range(n)] + [evals[x] % p] for x in range(n)]
    for col in range(n):
        piv = next(rw for rw in range(col, n) if mat[rw][col] % p)
        mat[col], mat[piv] = mat[piv], mat[col]
        inv = inv_mod(mat[col][col], p)
        mat[col] = [v * inv % p for v in mat[col]]
        for rw in range(n):
            if rw != col and mat[rw][col]:
                f = mat[rw][col]
                mat[rw] = [(a - f * b) % p for a, b in zip(mat[rw], mat[col])]
    return [mat[i][n] for i in range(n)]


def uni_from_evals_and_hint(hint: int, evals: list[int], p: int = R_MOD) -> list[int]:
    """Evaluations at [0,2,3,..] + hint = p(0)+p(1) (univariate.rs:210-216)."""
    full = list(evals)
    full.insert(1, (hint - full[0]) % p)
    return uni_from_evals(full, p)


def uni_compress(coeffs: list[int]) -> list[int]:
    """Drop the linear coefficient (univariate.rs:185-193)."""
    assert len(coeffs) >= 2
    return coeffs[:1] + coeffs[2:]


# --------------------------------------------------------------------------- #
# Univariate-evaluation sweep (crates/jolt-kernels/src/reference/naive.rs:241-310,
# crates/jolt-sumcheck/tests/roundtrip.rs:26-97)
# --------------------------------------------------------------------------- #
def pair(evals: list[int], y: int, order: int):
    """Polynomial::sumcheck_eval_pair (dense.rs:311-320)."""
    if order == HIGH_TO_LOW:
        return evals[y], evals[y + len(evals) // 2]
    return evals[2 * y], evals[2 * y + 1]


def product_round_evals(tables: list[list[int]], degree: int, order: int = HIGH_TO_LOW, p: int = R_MOD):
    """s(t) = sum_y prod_j (lo_j + t*(hi_j - lo_j)) for t = 0..degree."""
    half = len(tables[0]) // 2
    out = []
    for t in range(degree + 1):
        acc = 0
        for y in range(half):
            prod = 1
            for tab in tables:
                lo, hi = pair(tab, y, order)
                prod = prod * (lo + t * (hi - lo)) % p
            acc += prod
        out.append(acc % p)
    return out


# --------------------------------------------------------------------------- #
# Engine fixtures (crates/jolt-sumcheck/src/tests.rs:1123-1235,
# crates/jolt-kernels/src/optimized/parity.rs:25-34)
# --------------------------------------------------------------------------- #
def dense_member_with_sum(num_rounds: int, total: int, seed: int, p: int = R_MOD) -> list[int]:
    size = 1 << num_rounds
    evals = [(seed + 31 * i + 11) % p for i in range(size)]
    evals[0] = (evals[0] + total - sum(evals)) % p
    return evals


def synthetic_point(length: int, seed: int, p: int = R_MOD) -> list[int]:
    return [((seed * 6364136223846793005) + (i * 2 + 3)) % (1 << 64) % p for i in range(length)]


class ProductMember:
    """Reference-tier member: product of m dense tables, degree m, fused contract
    (naive.rs:241-316 with Expr = product of the opening tables)."""

    def __init__(self, tables, order=HIGH_TO_LOW, p=R_MOD):
        self.tables = [list(t) for t in tables]
        self.order = order
        self.p = p
        self.rounds = len(tables[0]).bit_length() - 1
        self.degree = len(tables)

    def num_rounds(self):
        return self.rounds

    def prove_round(self, bind_c, rnd, previous_claim):
        if bind_c is not None:
            self.tables = [bind(t, bind_c, self.order, self.p) for t in self.tables]
        ev = product_round_evals(self.tables, self.degree, self.order, self.p)
        if (ev[0] + ev[1]) % self.p != previous_claim % self.p:
            raise ValueError(f"RoundCheckFailed round={rnd}")
        return uni_from_evals(ev, self.p)

    def finish_rounds(self, bind_c):
        self.tables = [bind(t, bind_c, self.order, self.p) for t in self.tables]

    def final_evals(self):
        return [t[0] for t in self.tables]


class IncClaimReductionKernel:
    """Restatement of the optimized stage-6b increment claim reduction
    (crates/jolt-kernels/src/optimized/inc_claim_reduction.rs:47-203). prepare (:50-120): the four upstream eq
    leaves collapse into A = eq(r_ram_rw) + gamma eq(r_ram_val), B = gamma^2 eq(s_reg_rw) + gamma^3 eq(s_reg_val)
    (scaled_eq_table = EqPolynomial::evals(point, Some(scale)), big-endian); the summand is A*RamInc + B*RdInc.
    prove_round (:166-186): bind_all low-to-high, group_evals at t in {0, 2} (:146-156),
    round_poly_from_skipped_evals (support.rs:450-460). output_claims (:196-203): ram_inc[0], rd_inc[0]."""

    def __init__(self, cycle_points, gamma, ram_inc, rd_inc, p=R_MOD):
        self.p = p
        n = len(ram_inc).bit_length() - 1
        assert all(len(pt) == n for pt in cycle_points) and len(rd_inc) == len(ram_inc)
        g2 = gamma * gamma % p
        def combine(first, fs, second, ss):
            a, b = eq_evals(first, fs, p), eq_evals(second, ss, p)
            return [(x + y) % p for x, y in zip(a, b)]
        self.ram_weights = combine(cycle_points[0], 1, cycle_points[1], gamma % p)
        self.rd_weights = combine(cycle_points[2], g2, cycle_points[3], g2 * gamma % p)
        self.ram_inc, self.rd_inc = [v % p for v in ram_inc], [v % p for v in rd_inc]
        self.rounds = n

    def tables(self):
        """term-major order of the device member: (A, RamInc), (B, RdInc)"""
        return [self.ram_weights, self.ram_inc, self.rd_weights, self.rd_inc]

    def num_rounds(self):
        return self.rounds

    def _bind(self, c):
        self.ram_inc = bind(self.ram_inc, c, LOW_TO_HIGH, self.p)
        self.rd_inc = bind(self.rd_inc, c, LOW_TO_HIGH, self.p)
        self.ram_weights = bind(self.ram_weights, c, LOW_TO_HIGH, self.p)
        self.rd_weights = bind(self.rd_weights, c, LOW_TO_HIGH, self.p)

    def prove_round(self, bind_c, rnd, previous_claim):
        p = self.p
        if bind_c is not None:
            self._bind(bind_c)
        half = len(self.ram_inc) // 2
        e0 = e2 = 0
        for y in range(half):
            ram_lo, ram_hi = self.ram_inc[2 * y], self.ram_inc[2 * y + 1]
            rd_lo, rd_hi = self.rd_inc[2 * y], self.rd_inc[2 * y + 1]
            a_lo, a_hi = self.ram_weights[2 * y], self.ram_weights[2 * y + 1]
            b_lo, b_hi = self.rd_weights[2 * y], self.rd_weights[2 * y + 1]
            e0 += a_lo * ram_lo + b_lo * rd_lo
            e2 += (a_hi + a_hi - a_lo) * (ram_hi + ram_hi - ram_lo) + (b_hi + b_hi - b_lo) * (rd_hi + rd_hi - rd_lo)
        e0, e2 = e0 % p, e2 % p
        return uni_from_evals([e0, (previous_claim - e0) % p, e2], p)   # round_poly_from_skipped_evals

    def finish_rounds(self, bind_c):
        self._bind(bind_c)

    def final_evals(self):
        return [self.ram_weights[0], self.ram_inc[0], self.rd_weights[0], self.rd_inc[0]]

    def output_claims(self):
        return {"ram_inc": self.ram_inc[0], "rd_inc": self.rd_inc[0]}


def prove_batch(members_desc, members, max_num_vars, max_degree, claimed_sum, challenge_fn, p=R_MOD):
    """jolt-sumcheck/src/prover.rs:193-362. members_desc: list of dicts with
    input_claim, coefficient, rounds, offset. challenge_fn(round, coeffs)->challenge
    stands in for recorder.absorb_round (the transcript stays host-side)."""
    two_inv = inv_mod(2, p)
    claims = [d["input_claim"] * pow(2, max_num_vars - d["rounds"], p) % p for d in members_desc]
    running = claimed_sum % p
    challenges = []
    pending = [None] * len(members)
    round_polys = []
    for rnd in range(max_num_vars):
        batched = [0] * (max_degree + 1)
        work = []
        for i, (m, d) in enumerate(zip(members, members_desc)):
            active = d["offset"] <= rnd < d["offset"] + d["rounds"]
            if not active:
                claims[i] = claims[i] * two_inv % p
                batched[0] = (batched[0] + d["coefficient"] * claims[i]) % p
                continue
            b, pending[i] = pending[i], None
            work.append((i, m.prove_round(b, rnd - d["offset"], claims[i])))
        for i, poly in work:
            assert len(poly) - 1 <= max_degree
            for k, c in enumerate(poly):
                batched[k] = (batched[k] + members_desc[i]["coefficient"] * c) % p
        while len(batched) > 2 and batched[-1] == 0:  # trim_round_polynomial :163-168
            batched.pop()
        s = (uni_evaluate(batched, 0, p) + uni_evaluate(batched, 1, p)) % p
        if s != running:
            raise ValueError(f"RoundCheckFailed round={rnd}")
        c = challenge_fn(rnd, batched) % p
        running = uni_evaluate(batched, c, p)
        challenges.append(c)
        round_polys.append(batched)
        for i, poly in work:
            claims[i] = uni_evaluate(poly, c, p)
            pending[i] = c
    for m, b in zip(members, pending):
        if b is not None:
            m.finish_rounds(b)
    return {"challenges": challenges, "final_claim": running, "member_claims": claims,
            "round_polys": round_polys}


# --------------------------------------------------------------------------- #
# BN254 G1 (standard alt_bn128: y^2 = x^3 + 3 over Fq; None = identity).
# Restates what ark_ec supplies to crates/jolt-crypto/src/ec/bn254/mod.rs:85-212.
# --------------------------------------------------------------------------- #
def g1_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q_MOD)


def g1_add(P, Qp):
    if P is None:
        return Qp
    if Qp is None:
        return P
    x1, y1 = P
    x2, y2 = Qp
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * inv_mod(2 * y1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * inv_mod(x2 - x1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    return x3, (lam * (x1 - x3) - y1) % Q_MOD


def g1_scalar_mul(P, k: int):
    k %= R_MOD
    acc = None
    add = P
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


def g1_msm_naive(bases, scalars):
    """group_laws.rs:69-79: msm == sum_i scalar_mul(bases[i], scalars[i])."""
    assert len(bases) == len(scalars), "msm: bases/scalars length mismatch"  # mod.rs:200-204
    acc = None
    for P, s in zip(bases, scalars):
        acc = g1_add(acc, g1_scalar_mul(P, s))
    return acc


def g1_msm_pippenger(bases, scalars, c: int | None = None):
    """Bucket method on canonical scalars (into_bigint, mod.rs:208), arkworks window
    heuristic c ~ ln(n) + 2. Same group value as the naive sum."""
    assert len(bases) == len(scalars), "msm: bases/scalars length mismatch"
    n = len(bases)
    if n == 0:
        return None  # group_laws.rs:143-146
    if c is None:
        import math
        c = 3 if n < 32 else int(math.log(n)) + 2
    total = None
    nwin = (254 + c - 1) // c
    for w in reversed(range(nwin)):
        if total is not None:
            for _ in range(c):
                total = g1_add(total, total)
        buckets = [None] * ((1 << c) - 1)
        for P, s in zip(bases, scalars):
            d = ((s % R_MOD) >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = g1_add(buckets[d - 1], P)
        run = None
        acc = None
        for b in reversed(buckets):
            run = g1_add(run, b)
            acc = g1_add(acc, run)
        total = g1_add(total, acc)
    return total


def signed_to_fr(magnitude: int, is_positive: bool, p: int = R_MOD) -> int:
    """jolt_field::signed::SignedBigInt { magnitude, is_positive } (crates/jolt-field/src/signed.rs:25-32) as a field
    value: +-magnitude mod r; "zero is not canonicalized" (:16-17), a zero magnitude with either sign is 0."""
    return magnitude % p if is_positive else (-magnitude) % p


def msm_small(bases, values, kind: str):
    """The legacy small-scalar facade (crates/jolt-prover-legacy/src/msm/mod.rs:27-158): msm_u8 .. msm_i128 and
    msm_s64 / msm_s128 take primitive (or sign-magnitude) scalars; the VALUE is the field-scalar MSM of their promotions
    (from_u64 / from_i64 / from_u128 / from_i128, crates/jolt-field/src/bn254/mod.rs:265-298). The U8Scalars arm
    (:35-47) and msm_u8 (:96-106) dispatch all-zero -> zero(), all <= 1 -> msm_binary, else msm_u8 - three schedules of
    the same sum, restated here to pin that they agree. `values`: ints, or (magnitude, is_positive) pairs for s64 / s128."""
    assert len(bases) == len(values), "msm: bases/scalars length mismatch"  # KeyLengthError, :48-50
    if kind in ("s64", "s128"):
        sc = [signed_to_fr(*v) if isinstance(v, tuple) else v % R_MOD for v in values]
    else:
        sc = [v % R_MOD for v in values]
    if kind == "u8":
        if all(v == 0 for v in values):
            return None                                  # Self::zero()
        if all(v <= 1 for v in values):                  # msm_binary: the plain sum of the selected bases
            acc = None
            for P, v in zip(bases, values):
                if v:
                    acc = g1_add(acc, P)
            return acc
    return g1_msm_naive(bases, sc)


def batch_msm(bases, columns):
    """VariableBaseMSM::batch_msm (msm/mod.rs:160-168): column k against the PREFIX bases[..len(column k)].
    columns: (values, kind) pairs."""
    return [msm_small(bases[: len(vals)], vals, kind) for vals, kind in columns]


def msm_rows(bases, matrix, rows: int, kind: str):
    """Dory tier-1 row commitments (crates/jolt-dory/src/streaming.rs:113-201): one MSM per matrix row against the same
    bases[..row_width]."""
    w = len(matrix) // rows
    assert w * rows == len(matrix)
    return [msm_small(bases[:w], matrix[r * w:(r + 1) * w], kind) for r in range(rows)]


def batch_g1_additions_multi_affine(bases, indices_sets, q: int = Q_MOD):
    """crates/jolt-crypto/src/ec/bn254/batch_addition.rs:53-150: one affine sum per index set. Every level pairs
    neighbours (2j, 2j + 1) of each working set, all pairs of a level share one batch inversion (ark_ff's
    batch_inversion leaves zero elements zero, :101-105), lambda = (y2 - y1) inv, x3 = lambda^2 - x1 - x2,
    y3 = lambda (x1 - x3) - y1 (:121-125), an odd last element moves up unchanged (:136-140). Points are (x, y)
    integer pairs; the identity (empty set, :69-70) is None. The distinct-x precondition is NOT checked, as in the
    reference: a violating pair gets inv = 0 and an off-curve result."""
    work = [[None] if not idx else [bases[i] for i in idx] for idx in indices_sets]
    while any(len(w) >= 2 for w in work):
        nxt = []
        for w in work:
            row = []
            for j in range(len(w) // 2):
                (x1, y1), (x2, y2) = w[2 * j], w[2 * j + 1]
                d = (x2 - x1) % q
                inv = pow(d, -1, q) if d else 0
                lam = (y2 - y1) * inv % q
                x3 = (lam * lam - x1 - x2) % q
                row.append((x3, (lam * (x1 - x3) - y1) % q))
            if len(w) % 2:
                row.append(w[-1])
            nxt.append(row)
        work = nxt
    return [w[0] for w in work]


# --------------------------------------------------------------------------- #
# HyperKZG prover side (crates/jolt-hyperkzg/src/scheme.rs:54-158, kzg.rs:15-126)
# --------------------------------------------------------------------------- #
def hyperkzg_setup_from_secret(beta: int, max_degree: int, g1=G1_GEN):
    """scheme.rs:54-73: g1_powers[i] = beta^i * g1."""
    out, cur = [], g1
    for _ in range(max_degree + 1):
        out.append(cur)
        cur = g1_scalar_mul(cur, beta)
    return out


def kzg_commit(coeffs, g1_powers):
    """kzg.rs:15-27."""
    assert len(coeffs) <= len(g1_powers), "SrsTooSmall"
    return g1_msm_naive(g1_powers[: len(coeffs)], coeffs)


def compute_witness_polynomial(f, u, p=R_MOD):
    """kzg.rs:34-46: h = f / (x - u) by reverse Horner."""
    d = len(f)
    if d <= 1:
        return []
    h = [0] * (d - 1)
    acc = 0
    for i in range(d - 1, 0, -1):
        acc = (f[i] + acc * u) % p
        h[i - 1] = acc
    return h


def eval_univariate(coeffs, u, p=R_MOD):
    """kzg.rs:51-59."""
    res, power = 0, 1
    for c in coeffs:
        res = (res + c * power) % p
        power = power * u % p
    return res


def fold_polynomials(evals, point, p=R_MOD):
    """scheme.rs:88-114: LowToHigh folds with point[1..] visited back to front."""
    polys = [list(evals)]
    for xi in reversed(point[1:]):
        prev = polys[-1]
        polys.append([(prev[2 * j] + xi * (prev[2 * j + 1] - prev[2 * j])) % p
                      for j in range(len(prev) // 2)])
    return polys


def hyperkzg_open(g1_powers, evals, point, challenge_r, challenge_q, p=R_MOD):
    """scheme.rs:122-158 + kzg.rs:69-126 with the two Fiat-Shamir challenges supplied
    by callbacks: challenge_r(com) -> r, challenge_q(v) -> q."""
    ell = len(point)
    assert ell > 0 and len(evals) == 1 << ell
    polys = fold_polynomials(evals, point, p)
    com = [kzg_commit(pl, g1_powers) for pl in polys[1:]]
    r = challenge_r(com) % p
    u = [r, (-r) % p, r * r % p]
    v = [[eval_univariate(fj, ui, p) for fj in polys] for ui in u]
    q = challenge_q(v) % p
    poly_len = len(polys[0])
    b_poly = [0] * poly_len
    qj = 1
    for fj in polys:
        for i, cf in enumerate(fj):
            b_poly[i] = (b_poly[i] + qj * cf) % p
        qj = qj * q % p
    w = []
    for ui in u:
        h = compute_witness_polynomial(b_poly, ui, p)
        w.append(g1_msm_naive(g1_powers[: len(h)], h))
    return {"com": com, "w": w, "v": v}


# --------------------------------------------------------------------------- #
# Documented PRNG for synthetic inputs: SplitMix64.
# --------------------------------------------------------------------------- #
def splitmix64(state: int) -> tuple[int, int]:
    state = (state + 0x9E3779B97F4A7C15) & MASK64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return state, z ^ (z >> 31)


def random_fr(seed: int, count: int, p: int = R_MOD) -> list[int]:
    """count uniform-ish field elements: 4 SplitMix64 words -> 256-bit int -> mod p."""
    out, st = [], seed & MASK64
    for _ in range(count):
        v = 0
        for k in range(4):
            st, w = splitmix64(st)
            v |= w << (64 * k)
        out.append(v % p)
    return out
