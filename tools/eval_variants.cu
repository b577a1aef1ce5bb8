// Throughput of the MiMC transition-constraint inner loop (k_eval_constraints, ZKB_AIR_MIMC) on register-resident data:
//   per (point, column):  a1 = cur + rc; a7 = a1^7 (2 squarings + 2 multiplications); ev = nxt - a7; t += coef * ev
// Variants:
//   0: as shipped before this experiment: squarings through fe_mul, every coef * ev reduced and added mod p
//   1: dedicated squaring (10 wide products instead of 16, off-diagonal sum doubled by a funnel shift)
//   2: lazy accumulation: t kept as a 288-bit integer sum of unreduced 256-bit products, reduced once per point
//   3: both
// All four must produce the same field elements (checksum printed).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/eval_variants tools/eval_variants.cu && ./tools/eval_variants
#include <cstdio>
#include "reduce_variants.cuh"
using namespace zkb;

#define COLS 8
// R = 0: shipped mad-chain folds; 1: variant B; 2: variant C (fewer IMADs, more ALU work)
template <int R> __device__ __forceinline__ fe red(const uint32_t w[8]) { return R == 0 ? fe_reduce256(w) : (R == 1 ? reduce_B(w) : reduce_C(w)); }
template <int R> __device__ __forceinline__ fe mulr(const fe& a, const fe& b) { uint32_t w[8]; mul_wide(a, b, w); return red<R>(w); }
template <int R> __device__ __forceinline__ fe sqrr(const fe& a) { uint32_t w[8]; sqr_wide(a, w); return red<R>(w); }
template <int V>
__global__ void __launch_bounds__(128) k(fe* io, const fe* coef, int iters) {
    fe cur[COLS], nxt[COLS];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int u = 0; u < COLS; u++) { cur[u] = fe_load(io + (size_t)t * 2 * COLS + 2 * u); nxt[u] = fe_load(io + (size_t)t * 2 * COLS + 2 * u + 1); }
    const fe rc = fe_load(coef + 32 + (threadIdx.x & 31));
    fe total = fe_zero();
    for (int i = 0; i < iters; i++) {
        fe tt = fe_zero();
        acc288 acc; acc288_zero(acc);
#pragma unroll
        for (int u = 0; u < COLS; u++) {
            const fe a1 = fe_add(cur[u], rc);
            fe a2, a4;
            constexpr int R = V >> 2;
            if (V & 1) { a2 = sqrr<R>(a1); a4 = sqrr<R>(a2); } else { a2 = mulr<R>(a1, a1); a4 = mulr<R>(a2, a2); }
            const fe a6 = mulr<R>(a4, a2), a7 = mulr<R>(a6, a1);
            const fe ev = fe_sub(nxt[u], a7);
            const fe cf = fe_ldg(coef + u);
            if (V & 2) acc288_mad(acc, cf, ev); else tt = fe_add(tt, fe_mul(cf, ev));
            cur[u] = a7;  // keep the data moving so nothing is hoisted
        }
        if (V & 2) tt = acc288_reduce(acc);
        total = fe_add(total, tt);
    }
    fe_store(io + (size_t)t * 2 * COLS, total);
}

template <int V> void run(const char* name, fe* io, fe* coef, unsigned long long* ref, bool check) {
    const int blocks = 148 * 8, threads = 128, iters = 400;
    cudaMemset(io, 0x3c, (size_t)blocks * threads * 2 * COLS * 16);
    k<V><<<blocks, threads>>>(io, coef, 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(io, 0x3c, (size_t)blocks * threads * 2 * COLS * 16);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(io, coef, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2];
    cudaMemcpy(h, io + 12345 * 2 * COLS, 16, cudaMemcpyDeviceToHost);
    const double items = (double)blocks * threads * COLS * iters;
    printf("%-44s %7.3f ms  %7.1f G (point,column)/s   checksum %016llx%016llx %s\n", name, ms, items / (ms * 1e-3) / 1e9, h[1], h[0],
           check ? ((h[0] == ref[0] && h[1] == ref[1]) ? "(matches 0)" : "(DIFFERS from 0!)") : "");
    if (!check) { ref[0] = h[0]; ref[1] = h[1]; }
}

int main() {
    fe *io, *coef;
    cudaMalloc(&io, (size_t)148 * 8 * 128 * 2 * COLS * 16);
    cudaMalloc(&coef, 64 * 16);
    unsigned long long host[128];
    for (int i = 0; i < 128; i++) host[i] = (0x9E3779B97F4A7C15ULL * (i + 1)) >> (i & 1);
    cudaMemcpy(coef, host, 1024, cudaMemcpyHostToDevice);
    unsigned long long ref[2] = {0, 0};
    run<0>("0 fe_mul squarings, reduced accumulation", io, coef, ref, false);
    run<1>("1 dedicated squaring", io, coef, ref, true);
    run<2>("2 lazy 288-bit accumulation", io, coef, ref, true);
    run<3>("3 both", io, coef, ref, true);
    run<7>("7 both + reduction B (shift/funnel)", io, coef, ref, true);
    run<11>("11 both + reduction C (limb shift + c1)", io, coef, ref, true);
    return 0;
}
