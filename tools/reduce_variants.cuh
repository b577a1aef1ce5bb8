// Alternative reductions of a 256-bit product modulo p = 2^128 - 45*2^40 + 1 (experiments).  Shipped since round 2: variant E =
// fe_reduce256 (K = 0x2D00 fold) and, for twiddles, variant F = fe_mul_pre4 (four pre-shifted copies), both in csrc/f128.cuh.
//   A: round-1 reduction (mad chains with C = {0xFFFFFFFF, 0x2CFF})
//   B: hi*C = ((hi*45) << 40) - hi with funnel shifts                       (ALU-heavy)
//   C: hi*C = (hi << 32) - hi + ((hi*0x2CFF) << 32): limb shift/sub + 4 IMAD.WIDE by the small constant
#pragma once
#include "../zk_stark_project_b200/csrc/f128.cuh"
namespace zkb {
__device__ __forceinline__ fe reduce_B(const uint32_t v[8]) {
    const uint32_t h0 = v[4], h1 = v[5], h2 = v[6], h3 = v[7];
    uint32_t t0, t1, t2, t3, t4;
    {
        uint32_t e0, e1, e2, e3, o0, o1, o2, o3;
        asm("mul.lo.u32 %0, %8, 45;\n\t mul.hi.u32 %1, %8, 45;\n\t mul.lo.u32 %2, %10, 45;\n\t mul.hi.u32 %3, %10, 45;\n\t"
            "mul.lo.u32 %4, %9, 45;\n\t mul.hi.u32 %5, %9, 45;\n\t mul.lo.u32 %6, %11, 45;\n\t mul.hi.u32 %7, %11, 45;"
            : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3), "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(h0), "r"(h1), "r"(h2), "r"(h3));
        t0 = e0;
        asm("add.cc.u32 %0, %4, %8;\n\t addc.cc.u32 %1, %5, %9;\n\t addc.cc.u32 %2, %6, %10;\n\t addc.u32 %3, %7, 0;"
            : "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4) : "r"(e1), "r"(e2), "r"(e3), "r"(o3), "r"(o0), "r"(o1), "r"(o2));
    }
    const uint32_t s1 = t0 << 8, s2 = __funnelshift_l(t0, t1, 8), s3 = __funnelshift_l(t1, t2, 8), s4 = __funnelshift_l(t2, t3, 8),
                   s5 = __funnelshift_l(t3, t4, 8);
    uint32_t u0, u1, u2, u3, u4, u5;
    asm("add.cc.u32 %0, %6, 0;\n\t addc.cc.u32 %1, %7, %10;\n\t addc.cc.u32 %2, %8, %11;\n\t addc.cc.u32 %3, %9, %12;\n\t addc.cc.u32 %4, %13, 0;\n\t addc.u32 %5, %14, 0;"
        : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3), "=r"(u4), "=r"(u5)
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5));
    asm("sub.cc.u32 %0, %0, %6;\n\t subc.cc.u32 %1, %1, %7;\n\t subc.cc.u32 %2, %2, %8;\n\t subc.cc.u32 %3, %3, %9;\n\t subc.cc.u32 %4, %4, 0;\n\t subc.u32 %5, %5, 0;"
        : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5) : "r"(h0), "r"(h1), "r"(h2), "r"(h3));
    uint32_t q0, q1;
    asm("mul.lo.u32 %0, %2, 45;\n\t mul.hi.u32 %1, %2, 45;\n\t mad.lo.u32 %1, %3, 45, %1;" : "=&r"(q0), "=&r"(q1) : "r"(u4), "r"(u5));
    const uint32_t w1 = q0 << 8, w2 = __funnelshift_l(q0, q1, 8);
    uint32_t cy, bw;
    asm("add.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, %6;\n\t addc.cc.u32 %3, %3, 0;\n\t addc.u32 %4, 0, 0;"
        : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "=r"(cy) : "r"(w1), "r"(w2));
    asm("sub.cc.u32 %0, %0, %5;\n\t subc.cc.u32 %1, %1, %6;\n\t subc.cc.u32 %2, %2, 0;\n\t subc.cc.u32 %3, %3, 0;\n\t subc.u32 %4, 0, 0;"
        : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "=r"(bw) : "r"(u4), "r"(u5));
    fe out; out.x[0] = u0; out.x[1] = u1; out.x[2] = u2; out.x[3] = u3;
    return fe_canon(out, cy + bw);
}

// C: value = lo + (hi << 32) - hi + ((hi * c1) << 32),  c1 = 0x2CFF
__device__ __forceinline__ fe reduce_C(const uint32_t v[8]) {
    const uint32_t h0 = v[4], h1 = v[5], h2 = v[6], h3 = v[7];
    const uint32_t c1 = ZKB_C1;
    // T = hi * c1, 5 limbs (even/odd products, each < 2^46)
    uint32_t t0, t1, t2, t3, t4;
    {
        uint32_t e0, e1, e2, e3, o0, o1, o2, o3;
        asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t mul.lo.u32 %2, %10, %12;\n\t mul.hi.u32 %3, %10, %12;\n\t"
            "mul.lo.u32 %4, %9, %12;\n\t mul.hi.u32 %5, %9, %12;\n\t mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;"
            : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3), "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(h0), "r"(h1), "r"(h2), "r"(h3), "r"(c1));
        t0 = e0;
        asm("add.cc.u32 %0, %4, %8;\n\t addc.cc.u32 %1, %5, %9;\n\t addc.cc.u32 %2, %6, %10;\n\t addc.u32 %3, %7, 0;"
            : "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4) : "r"(e1), "r"(e2), "r"(e3), "r"(o3), "r"(o0), "r"(o1), "r"(o2));
    }
    // U (6 limbs) = lo + ((hi + T) << 32) - hi ; first W = hi + T (5 limbs + carry into a 6th)
    uint32_t w0, w1, w2, w3, w4;
    asm("add.cc.u32 %0, %5, %10;\n\t addc.cc.u32 %1, %6, %11;\n\t addc.cc.u32 %2, %7, %12;\n\t addc.cc.u32 %3, %8, %13;\n\t addc.u32 %4, %9, 0;"
        : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4) : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(t4), "r"(h0), "r"(h1), "r"(h2), "r"(h3));
    uint32_t u0 = v[0], u1, u2, u3, u4, u5;
    asm("add.cc.u32 %0, %5, %8;\n\t addc.cc.u32 %1, %6, %9;\n\t addc.cc.u32 %2, %7, %10;\n\t addc.cc.u32 %3, %11, 0;\n\t addc.u32 %4, %12, 0;"
        : "=r"(u1), "=r"(u2), "=r"(u3), "=r"(u4), "=r"(u5) : "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(w4));
    asm("sub.cc.u32 %0, %0, %6;\n\t subc.cc.u32 %1, %1, %7;\n\t subc.cc.u32 %2, %2, %8;\n\t subc.cc.u32 %3, %3, %9;\n\t subc.cc.u32 %4, %4, 0;\n\t subc.u32 %5, %5, 0;"
        : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5) : "r"(h0), "r"(h1), "r"(h2), "r"(h3));
    // second fold: top = u5:u4 < 2^46:  r = u + (top << 32) - top + ((top * c1) << 32)
    uint32_t q0, q1, q2;   // top * c1 < 2^60 -> add top itself: W2 = top + top*c1 (3 limbs)
    asm("mul.lo.u32 %0, %3, %5;\n\t mul.hi.u32 %1, %3, %5;\n\t mad.lo.u32 %1, %4, %5, %1;\n\t"
        "add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, %4;\n\t addc.u32 %2, 0, 0;"
        : "=&r"(q0), "=&r"(q1), "=&r"(q2) : "r"(u4), "r"(u5), "r"(c1));
    uint32_t cy, bw;
    asm("add.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, %6;\n\t addc.cc.u32 %3, %3, %7;\n\t addc.u32 %4, 0, 0;"
        : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "=r"(cy) : "r"(q0), "r"(q1), "r"(q2));
    asm("sub.cc.u32 %0, %0, %5;\n\t subc.cc.u32 %1, %1, %6;\n\t subc.cc.u32 %2, %2, 0;\n\t subc.cc.u32 %3, %3, 0;\n\t subc.u32 %4, 0, 0;"
        : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "=r"(bw) : "r"(u4), "r"(u5));
    fe out; out.x[0] = u0; out.x[1] = u1; out.x[2] = u2; out.x[3] = u3;
    return fe_canon(out, cy + bw);
}


// A (round 1, shipped then): folds with C = {0xFFFFFFFF, 0x2CFF} as multiply-accumulate chains; ptxas splits the products by
// 0xFFFFFFFF into IMAD + quarter-rate IMAD.HI
__device__ __forceinline__ fe reduce_A_round1(const uint32_t v[8]) {
    uint32_t r0 = v[0], r1 = v[1], r2 = v[2], r3 = v[3], r4, r5;
    uint32_t o0, o1, o2, o3, o4;
    const uint32_t h0 = v[4], h1 = v[5], h2 = v[6], h3 = v[7];
    const uint32_t c0 = ZKB_C0, c1 = ZKB_C1;
    // first fold: (r0..r5) = lo + hi * C
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %11, %15, %0;\n\t madc.hi.cc.u32 %1, %11, %15, %1;\n\t"
        "madc.lo.cc.u32 %2, %13, %15, %2;\n\t madc.hi.cc.u32 %3, %13, %15, %3;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "mad.lo.cc.u32 %2, %12, %16, %2;\n\t madc.hi.cc.u32 %3, %12, %16, %3;\n\t"
        "madc.lo.cc.u32 %4, %14, %16, %4;\n\t madc.hi.u32 %5, %14, %16, 0;\n\t"
        "mul.lo.u32 %6, %12, %15;\n\t mul.hi.u32 %7, %12, %15;\n\t"
        "mul.lo.u32 %8, %14, %15;\n\t mul.hi.u32 %9, %14, %15;\n\t"
        "mad.lo.cc.u32 %6, %11, %16, %6;\n\t madc.hi.cc.u32 %7, %11, %16, %7;\n\t"
        "madc.lo.cc.u32 %8, %13, %16, %8;\n\t madc.hi.cc.u32 %9, %13, %16, %9;\n\t"
        "addc.u32 %10, 0, 0;\n\t"
        "add.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t addc.cc.u32 %3, %3, %8;\n\t"
        "addc.cc.u32 %4, %4, %9;\n\t addc.u32 %5, %5, %10;\n\t"
        "}"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "=&r"(r4), "=&r"(r5),
          "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3), "=&r"(o4)
        : "r"(h0), "r"(h1), "r"(h2), "r"(h3), "r"(c0), "r"(c1));
    // second fold: top = r5:r4 < 2^46;  r += top * C, counting wraps of 2^128 (at most one)
    uint32_t cy;
    asm("{\n\t"
        ".reg .u32 t;\n\t"
        "mul.lo.u32 t, %6, %8;\n\t"  // t1*c1 < 2^28, weight 2^64
        "mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "addc.cc.u32 %2, %2, t;\n\t addc.cc.u32 %3, %3, 0;\n\t addc.u32 %4, 0, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %8, %1;\n\t madc.hi.cc.u32 %2, %5, %8, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %6, %7, %1;\n\t madc.hi.cc.u32 %2, %6, %7, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t addc.u32 %4, %4, 0;\n\t"
        "}"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "=&r"(cy)
        : "r"(r4), "r"(r5), "r"(c0), "r"(c1));
    // a wrap (cy) and "value >= p" are both fixed by adding C modulo 2^128
    fe out; out.x[0] = r0; out.x[1] = r1; out.x[2] = r2; out.x[3] = r3;
    return fe_canon(out, cy);
}


template <int V> __device__ __forceinline__ fe mulv(const fe& a, const fe& b) {
    uint32_t w[8];
    mul_wide(a, b, w);
    if (V == 0) return reduce_A_round1(w);
    if (V == 1) return reduce_B(w);
    if (V == 4) return fe_reduce256(w);   // E, shipped
    return reduce_C(w);
}

}  // namespace zkb
