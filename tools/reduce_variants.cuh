// Alternative reductions of a 256-bit product modulo p = 2^128 - 45*2^40 + 1 (experiments; the shipped one is fe_reduce256):
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


// E: 2^128 = k*2^32 - 1 (mod p) with k = 45*2^8 = 0x2D00, so hi*2^128 = ((hi*k) << 32) - hi: every product is a 32 x 14-bit
//    IMAD.WIDE by a small immediate (no multiplications by 0xFFFFFFFF, which ptxas splits into IMAD + quarter-rate IMAD.HI) and
//    the shift by 32 bits is a limb move.  U = lo - hi + ((hi*k) << 32) in six limbs (two's complement while negative), then
//    the 47-bit top is folded the same way.
__device__ __forceinline__ fe reduce_E(const uint32_t v[8]) {
    const uint32_t k = 0x2D00u;
    uint32_t d0, d1, d2, d3, d4, d5, o0, o1, o2, o3;
    asm("{\n\t"
        "sub.cc.u32 %0, %10, %14;\n\t subc.cc.u32 %1, %11, %15;\n\t subc.cc.u32 %2, %12, %16;\n\t subc.cc.u32 %3, %13, %17;\n\t"
        "subc.u32 %4, 0, 0;\n\t mov.u32 %5, %4;\n\t"                                   // sign extension of lo - hi
        "mad.lo.cc.u32 %1, %14, %18, %1;\n\t madc.hi.cc.u32 %2, %14, %18, %2;\n\t"     // (d1,d2) += h0 k
        "madc.lo.cc.u32 %3, %16, %18, %3;\n\t madc.hi.cc.u32 %4, %16, %18, %4;\n\t"    // (d3,d4) += h2 k
        "addc.u32 %5, %5, 0;\n\t"
        "mul.lo.u32 %6, %15, %18;\n\t mul.hi.u32 %7, %15, %18;\n\t"                    // (o0,o1) = h1 k
        "mul.lo.u32 %8, %17, %18;\n\t mul.hi.u32 %9, %17, %18;\n\t"                    // (o2,o3) = h3 k
        "add.cc.u32 %2, %2, %6;\n\t addc.cc.u32 %3, %3, %7;\n\t addc.cc.u32 %4, %4, %8;\n\t addc.u32 %5, %5, %9;\n\t"
        "}"
        : "=&r"(d0), "=&r"(d1), "=&r"(d2), "=&r"(d3), "=&r"(d4), "=&r"(d5), "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3)
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(k));
    // second fold: top = d5:d4 < 2^47;  W = ((top*k) << 32) - top  (three limbs, >= 0)
    uint32_t q0, q1, w0, w1, w2, cy;
    fe out;
    asm("{\n\t"
        "mul.lo.u32 %0, %10, %12;\n\t mul.hi.u32 %1, %10, %12;\n\t mad.lo.u32 %1, %11, %12, %1;\n\t"
        "sub.cc.u32 %2, 0, %10;\n\t subc.cc.u32 %3, %0, %11;\n\t subc.u32 %4, %1, 0;\n\t"
        "add.cc.u32 %5, %13, %2;\n\t addc.cc.u32 %6, %14, %3;\n\t addc.cc.u32 %7, %15, %4;\n\t addc.cc.u32 %8, %16, 0;\n\t addc.u32 %9, 0, 0;\n\t"
        "}"
        : "=&r"(q0), "=&r"(q1), "=&r"(w0), "=&r"(w1), "=&r"(w2), "=&r"(out.x[0]), "=&r"(out.x[1]), "=&r"(out.x[2]), "=&r"(out.x[3]), "=&r"(cy)
        : "r"(d4), "r"(d5), "r"(k), "r"(d0), "r"(d1), "r"(d2), "r"(d3));
    return fe_canon(out, cy);
}

// F: the twiddle is held as four pre-shifted copies W_i = w * 2^(32 i) mod p, so x * w = sum_i x_i * W_i is four aligned
//    32 x 128-bit rows (16 IMAD.WIDE) that sum to < 2^162: the 128-bit first fold disappears, only the 34-bit top is folded.
struct fe4 { fe w[4]; };
__device__ __forceinline__ fe mul_pre4(const fe& x, const fe4& t) {
    const uint32_t k = 0x2D00u;
    uint32_t e0, e1, e2, e3, e4, o0, o1, o2, o3, o4;
    asm("{\n\t"
        "mul.lo.u32 %0, %10, %14;\n\t mul.hi.u32 %1, %10, %14;\n\t mul.lo.u32 %2, %10, %16;\n\t mul.hi.u32 %3, %10, %16;\n\t"
        "mul.lo.u32 %5, %10, %15;\n\t mul.hi.u32 %6, %10, %15;\n\t mul.lo.u32 %7, %10, %17;\n\t mul.hi.u32 %8, %10, %17;\n\t"
        "mad.lo.cc.u32 %0, %11, %18, %0;\n\t madc.hi.cc.u32 %1, %11, %18, %1;\n\t madc.lo.cc.u32 %2, %11, %20, %2;\n\t madc.hi.cc.u32 %3, %11, %20, %3;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "mad.lo.cc.u32 %5, %11, %19, %5;\n\t madc.hi.cc.u32 %6, %11, %19, %6;\n\t madc.lo.cc.u32 %7, %11, %21, %7;\n\t madc.hi.cc.u32 %8, %11, %21, %8;\n\t"
        "addc.u32 %9, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, %12, %22, %0;\n\t madc.hi.cc.u32 %1, %12, %22, %1;\n\t madc.lo.cc.u32 %2, %12, %24, %2;\n\t madc.hi.cc.u32 %3, %12, %24, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %5, %12, %23, %5;\n\t madc.hi.cc.u32 %6, %12, %23, %6;\n\t madc.lo.cc.u32 %7, %12, %25, %7;\n\t madc.hi.cc.u32 %8, %12, %25, %8;\n\t"
        "addc.u32 %9, %9, 0;\n\t"
        "mad.lo.cc.u32 %0, %13, %26, %0;\n\t madc.hi.cc.u32 %1, %13, %26, %1;\n\t madc.lo.cc.u32 %2, %13, %28, %2;\n\t madc.hi.cc.u32 %3, %13, %28, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %5, %13, %27, %5;\n\t madc.hi.cc.u32 %6, %13, %27, %6;\n\t madc.lo.cc.u32 %7, %13, %29, %7;\n\t madc.hi.cc.u32 %8, %13, %29, %8;\n\t"
        "addc.u32 %9, %9, 0;\n\t"
        "}"
        : "=&r"(e0), "=&r"(e1), "=&r"(e2), "=&r"(e3), "=&r"(e4), "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3), "=&r"(o4)
        : "r"(x.x[0]), "r"(x.x[1]), "r"(x.x[2]), "r"(x.x[3]),
          "r"(t.w[0].x[0]), "r"(t.w[0].x[1]), "r"(t.w[0].x[2]), "r"(t.w[0].x[3]), "r"(t.w[1].x[0]), "r"(t.w[1].x[1]), "r"(t.w[1].x[2]), "r"(t.w[1].x[3]),
          "r"(t.w[2].x[0]), "r"(t.w[2].x[1]), "r"(t.w[2].x[2]), "r"(t.w[2].x[3]), "r"(t.w[3].x[0]), "r"(t.w[3].x[1]), "r"(t.w[3].x[2]), "r"(t.w[3].x[3]));
    // r = even + (odd << 32): six limbs, r5:r4 < 2^34
    uint32_t r1, r2, r3, r4, r5;
    asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t addc.cc.u32 %3, %8, %12;\n\t addc.u32 %4, %13, 0;"
        : "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5)
        : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4));
    uint32_t q0, q1, w0, w1, w2, cy;
    fe out;
    asm("{\n\t"
        "mul.lo.u32 %0, %10, %12;\n\t mul.hi.u32 %1, %10, %12;\n\t mad.lo.u32 %1, %11, %12, %1;\n\t"
        "sub.cc.u32 %2, 0, %10;\n\t subc.cc.u32 %3, %0, %11;\n\t subc.u32 %4, %1, 0;\n\t"
        "add.cc.u32 %5, %13, %2;\n\t addc.cc.u32 %6, %14, %3;\n\t addc.cc.u32 %7, %15, %4;\n\t addc.cc.u32 %8, %16, 0;\n\t addc.u32 %9, 0, 0;\n\t"
        "}"
        : "=&r"(q0), "=&r"(q1), "=&r"(w0), "=&r"(w1), "=&r"(w2), "=&r"(out.x[0]), "=&r"(out.x[1]), "=&r"(out.x[2]), "=&r"(out.x[3]), "=&r"(cy)
        : "r"(r4), "r"(r5), "r"(k), "r"(e0), "r"(r1), "r"(r2), "r"(r3));
    return fe_canon(out, cy);
}

template <int V> __device__ __forceinline__ fe mulv(const fe& a, const fe& b) {
    uint32_t w[8];
    mul_wide(a, b, w);
    if (V == 0) return fe_reduce256(w);
    if (V == 1) return reduce_B(w);
    if (V == 4) return reduce_E(w);
    return reduce_C(w);
}

}  // namespace zkb
