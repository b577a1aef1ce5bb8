// Host-side f128 scalars for the prover's control path (challenges, OOD points, divisor constants,
// root tables).  Bulk arithmetic never runs here — it runs in the CUDA kernels (f128.cuh).
// Field: winter-math f128::BaseElement as used at /root/reference src/training/prover.rs:9.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace zkb {

typedef unsigned __int128 u128;

struct HF {
    u128 v;
    static constexpr u128 modulus() { return (((u128)0xFFFFFFFFFFFFFFFFULL) << 64) | (u128)0xFFFFD30000000001ULL; }
    static constexpr uint64_t fold() { return 0x2CFFFFFFFFFFULL; }  // 2^128 mod p
    HF() : v(0) {}
    static HF raw(u128 x) { HF r; r.v = x; return r; }
    static HF from_u64(uint64_t x) { return raw(x); }
    static HF reduce(u128 x) { return raw(x >= modulus() ? x - modulus() : x); }
    bool operator==(const HF& o) const { return v == o.v; }
    bool operator!=(const HF& o) const { return v != o.v; }
    HF operator+(const HF& o) const { u128 r = v + o.v; if (r < v || r >= modulus()) r -= modulus(); return raw(r); }
    HF operator-(const HF& o) const { u128 r = v - o.v; if (v < o.v) r += modulus(); return raw(r); }
    HF operator*(const HF& o) const {
        uint64_t a0 = (uint64_t)v, a1 = (uint64_t)(v >> 64), b0 = (uint64_t)o.v, b1 = (uint64_t)(o.v >> 64);
        u128 ll = (u128)a0 * b0, lh = (u128)a0 * b1, hl = (u128)a1 * b0, hh = (u128)a1 * b1;
        u128 mid = lh + hl;
        u128 mid_carry = mid < lh ? ((u128)1 << 64) : 0;
        u128 lo = ll + (mid << 64);
        u128 hi = hh + (mid >> 64) + mid_carry + (lo < ll ? 1 : 0);
        // fold the high half twice with 2^128 = fold() (mod p)
        u128 f0 = (u128)(uint64_t)hi * fold(), f1 = (u128)(uint64_t)(hi >> 64) * fold();
        u128 s = f0 + (f1 << 64);
        uint64_t top = (uint64_t)(f1 >> 64) + (s < f0 ? 1 : 0);
        u128 r = lo + s;
        if (r < lo) top++;
        u128 r2 = r + (u128)top * fold();
        if (r2 < r) r2 += fold();
        if (r2 >= modulus()) r2 -= modulus();
        return raw(r2);
    }
    HF pow(u128 e) const {
        HF r = raw(1), b = *this;
        while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
        return r;
    }
    HF inv() const { return v == 0 ? *this : pow(modulus() - 2); }
    void to_bytes(uint8_t* out) const { memcpy(out, &v, 16); }
    static HF from_bytes(const uint8_t* in) { HF r; memcpy(&r.v, in, 16); return r; }
    static HF root_of_unity(int log_n) {
        HF r = raw((((u128)0x120532E7B364080AULL) << 64) | (u128)0x86B8723E1920F4AAULL);  // order 2^40
        for (int i = log_n; i < 40; i++) r = r * r;
        return r;
    }
};

// interpolate `evals` given over offset*<w_n> (n small: FRI remainder, periodic columns); O(n^2) is fine
static inline std::vector<HF> host_interpolate(const std::vector<HF>& evals, HF offset) {
    size_t n = evals.size();
    int ln = 0; while (((size_t)1 << ln) < n) ln++;
    HF winv = HF::root_of_unity(ln).inv(), ninv = HF::from_u64(n).inv(), oinv = offset.inv();
    std::vector<HF> c(n);
    HF wk = HF::raw(1), ok = HF::raw(1);
    for (size_t k = 0; k < n; k++) {  // c_k = offset^-k / n * sum_i e_i w^{-ik}
        HF acc, x = HF::raw(1);
        for (size_t i = 0; i < n; i++) { acc = acc + evals[i] * x; x = x * wk; }
        c[k] = acc * ninv * ok;
        wk = wk * winv; ok = ok * oinv;
    }
    return c;
}

}  // namespace zkb
