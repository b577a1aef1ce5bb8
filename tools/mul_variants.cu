// Butterfly throughput of f128 multiplier variants on sm_100a (register-resident, no memory traffic):
//   A: mad-chain folds with C = {0xFFFFFFFF, 0x2CFF}                       (round 1)
//   E: folds with K = 0x2D00, 2^128 = K*2^32 - 1                           (csrc/f128.cuh fe_reduce256 as shipped)
//   F: twiddle held as four pre-shifted copies, no first fold              (csrc/f128.cuh fe_mul_pre4, the NTT's multiplier)
//   B: hi*C = ((hi*45) << 40) - hi with funnel shifts                       (ALU-heavy)
//   C: hi*C = (hi << 32) - hi + ((hi*0x2CFF) << 32): limb shift/sub + 4 IMAD.WIDE by the small constant
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mul_variants tools/mul_variants.cu && ./tools/mul_variants
#include <cstdio>
#include "reduce_variants.cuh"
using namespace zkb;

// D: sums kept in [0, 2^128) instead of [0, p): a + v only wraps on a carry out of 2^128 (v canonical, so a + v - p fits);
//    differences and products accept such operands unchanged.  Outputs are canonicalised once at the end.
__device__ __forceinline__ fe add_lazy(const fe& a, const fe& b) {
    fe s; uint32_t cy;
    asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t addc.cc.u32 %3, %8, %12;\n\t addc.u32 %4, 0, 0;"
        : "=r"(s.x[0]), "=r"(s.x[1]), "=r"(s.x[2]), "=r"(s.x[3]), "=r"(cy)
        : "r"(a.x[0]), "r"(a.x[1]), "r"(a.x[2]), "r"(a.x[3]), "r"(b.x[0]), "r"(b.x[1]), "r"(b.x[2]), "r"(b.x[3]));
    const uint32_t m = 0u - cy;
    asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, %3, 0;"
        : "+r"(s.x[0]), "+r"(s.x[1]), "+r"(s.x[2]), "+r"(s.x[3]) : "r"(m & ZKB_C0), "r"(m & ZKB_C1));
    return s;
}
#ifndef UNITS
#define UNITS 8
#endif
#ifndef BPS
#define BPS 3
#endif
template <int V>
__global__ void __launch_bounds__(256) k(fe* io, const fe* tw, int iters) {
    fe x[UNITS], y[UNITS];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int u = 0; u < UNITS; u++) { x[u] = fe_load(io + (size_t)t * 2 * UNITS + 2 * u); y[u] = fe_load(io + (size_t)t * 2 * UNITS + 2 * u + 1); }
    const fe w = fe_load(tw + (threadIdx.x & 31));
    fe4 w4;
    if (V == 5) w4 = fe4_from(fe_canon(w, 0));
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < UNITS; u++) {
            const fe v = V == 5 ? fe_mul_pre4(y[u], w4) : mulv<(V == 3 ? 0 : V)>(y[u], w);
            const fe a = x[u];
            x[u] = V == 3 ? add_lazy(a, v) : fe_add(a, v);
            y[u] = fe_sub(a, v);
        }
    }
#pragma unroll
    for (int u = 0; u < UNITS; u++) { if (V == 3) { x[u] = fe_canon(x[u], 0); y[u] = fe_canon(y[u], 0); } fe_store(io + (size_t)t * 2 * UNITS + 2 * u, x[u]); fe_store(io + (size_t)t * 2 * UNITS + 2 * u + 1, y[u]); }
}

template <int V> void run(const char* name, fe* io, fe* tw, unsigned long long* ref, bool check) {
    const int blocks = 148 * BPS, threads = 256, iters = 2000;
    cudaMemset(io, 0x5a, (size_t)blocks * threads * 2 * UNITS * 16);
    // make inputs canonical-ish: clear the top bit of every element
    k<V><<<blocks, threads>>>(io, tw, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(io, 0x3c, (size_t)blocks * threads * 2 * UNITS * 16);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(io, tw, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[4];
    cudaMemcpy(h, io, 32, cudaMemcpyDeviceToHost);
    const double bf = (double)blocks * threads * UNITS * iters;
    printf("%-28s %7.3f ms  %7.1f G butterflies/s   checksum %016llx%016llx %s\n", name, ms, bf / (ms * 1e-3) / 1e9, h[1], h[0],
           check ? ((h[0] == ref[0] && h[1] == ref[1]) ? "(matches A)" : "(DIFFERS from A!)") : "");
    if (!check) { ref[0] = h[0]; ref[1] = h[1]; }
}

int main() {
    fe *io, *tw;
    cudaMalloc(&io, (size_t)148 * BPS * 256 * 2 * UNITS * 16);
    cudaMalloc(&tw, 32 * 16);
    unsigned long long host_tw[64];
    for (int i = 0; i < 64; i++) host_tw[i] = 0x9E3779B97F4A7C15ULL * (i + 1) >> (i & 1);
    cudaMemcpy(tw, host_tw, 512, cudaMemcpyHostToDevice);
    unsigned long long ref[2] = {0, 0};
    run<0>("A mad-chain folds (round 1)", io, tw, ref, false);
    run<1>("B (hi*45)<<40 - hi", io, tw, ref, true);
    run<2>("C (hi<<32) - hi + hi*c1<<32", io, tw, ref, true);
    run<3>("D round-1 mul, non-canonical sums", io, tw, ref, true);
    run<4>("E K = 0x2D00 fold (shipped fe_mul)", io, tw, ref, true);
    run<5>("F 4 pre-shifted copies (shipped NTT)", io, tw, ref, true);
    return 0;
}
