// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (no golden vectors exist in the reference;
// see oracle/README.md).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, link or call anything under oracle/.
//
// f128 field: CPU restatement of winter-math 0.12.0 `fields::f128::BaseElement`
// (third-party crate, not vendored under /root/reference; pinned by Cargo.lock:1303-1309).
// Selected by the reference at src/training/prover.rs:9, src/aggregation/prover.rs:12,
// src/helper.rs:11.   p = 2^128 - 45*2^40 + 1, canonical u128 representation,
// 16-byte little-endian serialisation, multiplicative generator 3, two-adicity 40.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

typedef unsigned __int128 u128;
typedef uint64_t u64;

static const u128 P = (((u128)0xFFFFFFFFFFFFFFFFULL) << 64) | (u128)0xFFFFD30000000001ULL;
static const u64 PC = 0x2CFFFFFFFFFFULL;  // 2^128 mod p = 45*2^40 - 1

struct Fe {
    u128 v;
    Fe() : v(0) {}
    explicit Fe(u128 x) : v(x >= P ? x - P : x) {}  // BaseElement::new reduces once
    bool operator==(const Fe& o) const { return v == o.v; }
    bool operator!=(const Fe& o) const { return v != o.v; }
};

static inline Fe fe_raw(u128 x) { Fe r; r.v = x; return r; }
static const Fe FE_ZERO = fe_raw(0);
static const Fe FE_ONE = fe_raw(1);

static inline Fe add(Fe a, Fe b) {
    u128 r = a.v + b.v;
    if (r < a.v || r >= P) r -= P;  // wrap-around subtraction is exact mod 2^128
    return fe_raw(r);
}
static inline Fe sub(Fe a, Fe b) {
    u128 r = a.v - b.v;
    if (a.v < b.v) r += P;
    return fe_raw(r);
}
static inline Fe neg(Fe a) { return a.v == 0 ? a : fe_raw(P - a.v); }

// 256-bit product folded with 2^128 = PC (mod p)
static inline Fe mul(Fe a, Fe b) {
    u64 a0 = (u64)a.v, a1 = (u64)(a.v >> 64), b0 = (u64)b.v, b1 = (u64)(b.v >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 mid = p01 + p10;
    u128 midc = mid < p01 ? 1 : 0;
    u128 lo = p00 + (mid << 64);
    u128 loc = lo < p00 ? 1 : 0;
    u128 hi = p11 + (mid >> 64) + (midc << 64) + loc;
    // hi * PC (up to 174 bits)
    u64 h0 = (u64)hi, h1 = (u64)(hi >> 64);
    u128 t0 = (u128)h0 * PC, t1 = (u128)h1 * PC;
    u128 s = t0 + (t1 << 64);
    u64 top = (u64)(t1 >> 64) + (s < t0 ? 1 : 0);
    u128 r = lo + s;
    top += (r < lo ? 1 : 0);
    u128 t2 = (u128)top * PC;  // < 2^94
    u128 r2 = r + t2;
    if (r2 < r) r2 += PC;  // one more wrap; cannot overflow again
    if (r2 >= P) r2 -= P;
    return fe_raw(r2);
}

static inline Fe pow(Fe b, u128 e) {
    Fe r = FE_ONE;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = mul(b, b);
        e >>= 1;
    }
    return r;
}
// winter-math: inv(0) = 0
static inline Fe inv(Fe a) { return a.v == 0 ? a : pow(a, P - 2); }

static const int TWO_ADICITY = 40;
static inline Fe two_adic_root() {
    // 23953097886125630542083529559205016746 = 3^((p-1)/2^40)
    return fe_raw((((u128)0x120532E7B364080AULL) << 64) | (u128)0x86B8723E1920F4AAULL);
}
static inline Fe get_root_of_unity(int log_n) {
    Fe r = two_adic_root();
    for (int i = log_n; i < TWO_ADICITY; i++) r = mul(r, r);
    return r;
}
static const u64 GENERATOR = 3;  // also the STARK domain offset

static inline void fe_to_bytes(Fe a, uint8_t* out) { memcpy(out, &a.v, 16); }  // x86: little-endian
static inline Fe fe_from_bytes(const uint8_t* in) { Fe r; memcpy(&r.v, in, 16); return r; }

// batch inversion (Montgomery trick); zeros stay zero like winter-math batch_inversion
static inline void batch_inv(Fe* x, size_t n) {
    if (n == 0) return;
    Fe* pre = new Fe[n];
    Fe acc = FE_ONE;
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (x[i].v != 0) acc = mul(acc, x[i]);
    }
    acc = inv(acc);
    for (size_t i = n; i-- > 0;) {
        if (x[i].v != 0) {
            Fe t = mul(acc, pre[i]);
            acc = mul(acc, x[i]);
            x[i] = t;
        }
    }
    delete[] pre;
}

}  // namespace orc
