// Host-side BN254 Fr for the O(rounds * degree) glue that the reference also keeps on the host:
// round-polynomial interpolation (crates/jolt-poly/src/univariate.rs:198-216), batching and
// claim updates in prove_batch (crates/jolt-sumcheck/src/prover.rs:246-343), the s(0)+s(1)
// round check. Nothing table-sized ever goes through this type.
#pragma once
#include <cstdint>
#include <cstring>

namespace jb {

struct HostFr {
    uint64_t l[4];  // Montgomery limbs, canonical

    static constexpr uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                                      0x30644e72e131a029ULL};
    static constexpr uint64_t INV = 0xc2e1f593efffffffULL;
    static constexpr uint64_t R1[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL,
                                       0x0e0a77c19a07df2fULL};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL,
                                       0x0216d0b17f4e44a5ULL};

    static HostFr zero() { return HostFr{{0, 0, 0, 0}}; }
    static HostFr one() { return HostFr{{R1[0], R1[1], R1[2], R1[3]}}; }
    static HostFr from_limbs(const uint64_t* p) {
        HostFr r;
        std::memcpy(r.l, p, 32);
        return r;
    }
    static HostFr from_u64(uint64_t v) {
        HostFr raw{{v, 0, 0, 0}};
        HostFr r2{{R2[0], R2[1], R2[2], R2[3]}};
        return raw * r2;
    }
    void store(uint64_t* p) const { std::memcpy(p, l, 32); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const HostFr& o) const { return std::memcmp(l, o.l, 32) == 0; }
    bool operator!=(const HostFr& o) const { return !(*this == o); }

    static bool geq_p(const uint64_t* a) {
        for (int i = 3; i >= 0; --i)
            if (a[i] != P[i]) return a[i] > P[i];
        return true;
    }
    static void sub_p(uint64_t* a) {
        unsigned __int128 borrow = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 t = (unsigned __int128)a[i] - P[i] - (uint64_t)borrow;
            a[i] = (uint64_t)t;
            borrow = (t >> 64) & 1;
        }
    }
    HostFr operator+(const HostFr& o) const {
        HostFr r;
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (unsigned __int128)l[i] + o.l[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
        if (geq_p(r.l)) sub_p(r.l);
        return r;
    }
    HostFr operator-(const HostFr& o) const {
        HostFr r;
        uint64_t borrow = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 t = (unsigned __int128)l[i] - o.l[i] - borrow;
            r.l[i] = (uint64_t)t;
            borrow = (uint64_t)(t >> 64) & 1;
        }
        if (borrow) {
            unsigned __int128 c = 0;
            for (int i = 0; i < 4; ++i) {
                c += (unsigned __int128)r.l[i] + P[i];
                r.l[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    HostFr operator-() const { return zero() - *this; }
    // word-serial Montgomery product (R = 2^256)
    HostFr operator*(const HostFr& o) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (unsigned __int128)l[j] * o.l[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * INV;
            c = (unsigned __int128)m * P[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (unsigned __int128)m * P[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        HostFr r{{t[0], t[1], t[2], t[3]}};
        if (t[4] || geq_p(r.l)) sub_p(r.l);
        return r;
    }
    HostFr pow(const uint64_t e[4]) const {
        HostFr acc = one(), base = *this;
        for (int i = 0; i < 256; ++i) {
            if ((e[i / 64] >> (i % 64)) & 1) acc = acc * base;
            base = base * base;
        }
        return acc;
    }
    HostFr inverse() const {  // Fermat; callers never invert zero
        uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
        return pow(e);
    }
};

}  // namespace jb
