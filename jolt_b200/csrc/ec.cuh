// BN254 G1 (y^2 = x^3 + 3 over Fq) device arithmetic for the Pippenger MSM.
// Accumulators use XYZZ ("extended Jacobian") coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;
// the identity is ZZ == 0. Mixed addition of an affine point costs 8M + 2S, a general addition
// 12M + 2S, a doubling 6M + 4S (Fq Montgomery products). The incomplete-formula cases the
// reference's arkworks backend handles inside its group law - P + P, P + (-P), identity operands
// (what crates/jolt-crypto/tests/group_laws.rs:12-131 pins) - are branched on explicitly.
#pragma once
#include "field.cuh"

namespace jb {

struct Affine {
    Fq x, y;  // identity: x == y == 0 (not on the curve, b = 3)
    __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

struct XYZZ {
    Fq x, y, zz, zzz;
    __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }
    __device__ __forceinline__ static XYZZ inf() {
        XYZZ r;
        r.x = Fq::one();
        r.y = Fq::one();
        r.zz = Fq::zero();
        r.zzz = Fq::zero();
        return r;
    }
    __device__ __forceinline__ static XYZZ from_affine(const Affine& p, bool negate) {
        XYZZ r;
        if (p.is_inf()) return inf();
        r.x = p.x;
        r.y = negate ? fp_neg(p.y) : p.y;
        r.zz = Fq::one();
        r.zzz = Fq::one();
        return r;
    }
};

// 2 * (affine p) -> XYZZ   (mdbl-2008-s-1, a = 0)
__device__ __forceinline__ XYZZ xyzz_double_affine(const Fq& px, const Fq& py) {
    XYZZ r;
    Fq u = fp_dbl(py);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(px, v);
    Fq xx = fp_sqr(px);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, py));
    r.zz = v;
    r.zzz = w;
    return r;
}

// acc = 2 * acc   (dbl-2008-s-1, a = 0)
__device__ __forceinline__ void xyzz_double(XYZZ& a) {
    if (a.is_inf()) return;
    Fq u = fp_dbl(a.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(a.x, v);
    Fq xx = fp_sqr(a.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    Fq x3 = fp_sub(fp_sqr(m), fp_dbl(s));
    Fq y3 = fp_sub(fp_mul(m, fp_sub(s, x3)), fp_mul(w, a.y));
    a.zz = fp_mul(v, a.zz);
    a.zzz = fp_mul(w, a.zzz);
    a.x = x3;
    a.y = y3;
}

// acc += (px, +-py)  (madd-2008-s). The affine operand is never the identity here.
__device__ __forceinline__ void xyzz_add_affine(XYZZ& a, const Fq& px, const Fq& py_in, bool negate) {
    Fq py = negate ? fp_neg(py_in) : py_in;
    if (a.is_inf()) {
        a.x = px;
        a.y = py;
        a.zz = Fq::one();
        a.zzz = Fq::one();
        return;
    }
    Fq u2 = fp_mul(px, a.zz);
    Fq s2 = fp_mul(py, a.zzz);
    Fq p = fp_sub(u2, a.x);
    Fq r = fp_sub(s2, a.y);
    if (p.is_zero()) {
        if (r.is_zero()) a = xyzz_double_affine(px, py);  // acc == P: double
        else a = XYZZ::inf();                              // acc == -P
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(a.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(a.y, ppp));
    a.zz = fp_mul(a.zz, pp);
    a.zzz = fp_mul(a.zzz, ppp);
    a.x = x3;
    a.y = y3;
}

// a += b  (add-2008-s)
__device__ __forceinline__ void xyzz_add(XYZZ& a, const XYZZ& b) {
    if (b.is_inf()) return;
    if (a.is_inf()) {
        a = b;
        return;
    }
    Fq u1 = fp_mul(a.x, b.zz);
    Fq u2 = fp_mul(b.x, a.zz);
    Fq s1 = fp_mul(a.y, b.zzz);
    Fq s2 = fp_mul(b.y, a.zzz);
    Fq p = fp_sub(u2, u1);
    Fq r = fp_sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) xyzz_double(a);
        else a = XYZZ::inf();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(u1, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(s1, ppp));
    a.zz = fp_mul(fp_mul(a.zz, b.zz), pp);
    a.zzz = fp_mul(fp_mul(a.zzz, b.zzz), ppp);
    a.x = x3;
    a.y = y3;
}

// XYZZ <-> memory (4 x 32 B)
__device__ __forceinline__ XYZZ ld_xyzz(const uint64_t* base, size_t idx) {
    XYZZ r;
    r.x = ld_elem_rw<Fq>(base, 4 * idx);
    r.y = ld_elem_rw<Fq>(base, 4 * idx + 1);
    r.zz = ld_elem_rw<Fq>(base, 4 * idx + 2);
    r.zzz = ld_elem_rw<Fq>(base, 4 * idx + 3);
    return r;
}
__device__ __forceinline__ void st_xyzz(uint64_t* base, size_t idx, const XYZZ& p) {
    st_elem(base, 4 * idx, p.x);
    st_elem(base, 4 * idx + 1, p.y);
    st_elem(base, 4 * idx + 2, p.zz);
    st_elem(base, 4 * idx + 3, p.zzz);
}

// a^(p-2) by square-and-multiply (one-off normalisations only)
__device__ __forceinline__ Fq fq_inverse(const Fq& a) {
    // p - 2, little-endian 32-bit words
    const uint32_t e[8] = {FqParams::P(0) - 2u, FqParams::P(1), FqParams::P(2), FqParams::P(3),
                           FqParams::P(4),      FqParams::P(5), FqParams::P(6), FqParams::P(7)};
    Fq acc = Fq::one();
    for (int i = 255; i >= 0; --i) {
        acc = fp_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fp_mul(acc, a);
    }
    return acc;
}

}  // namespace jb
