// Integer-pipe throughput micro-benchmark for sm_100a (MEASURED_PEAKS.json has no integer peak).
// Each kernel runs long independent dependency chains of one instruction class per thread; reported numbers are
// warp-instructions per cycle per SM sub-partition (SMSP) and thread-ops/s for the whole GPU at the observed clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak tools/int_peak.cu && ./tools/int_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void k(uint32_t* out, uint32_t a0, uint32_t b0, long long* cyc) {
    uint32_t x[CHAINS], y[CHAINS];
    unsigned long long w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = a0 + threadIdx.x + c; y[c] = b0 ^ (c * 77u); w[c] = x[c]; }
    long long t0 = clock64();
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(x[c]), "r"(y[c]));          // IMAD.WIDE.U32
            if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(a0));                // IMAD
            if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(a0));                // IMAD.HI
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c]));                                // IADD3 (or IMAD.IADD)
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(a0));            // LOP3
            if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y[c]));                      // SHF
            if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(x[c]) : "r"(y[c]));                       // PRMT
            if (OP == 8) {  // ALU + FMA mix, independent: LOP3 on one chain register, IMAD on the other
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(a0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(b0), "r"(a0));
            }
            if (OP == 9) {  // the multiplier's shape: IMAD.WIDE chain + carry adds
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(x[c]), "r"(y[c]));
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a0));
                asm volatile("addc.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(b0));
            }
            if (OP == 10) {  // two-source ALU op + two-source IMAD-class op
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));
                asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(b0));
            }
            if (OP == 7) { asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c])); asm volatile("addc.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(a0)); }  // IADD3 + IADD3.X
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= x[c] ^ y[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_iter, int sms, double mhz) {
    uint32_t* out; long long* cyc;
    const int blocks = sms * 2, threads = 512;  // 32 warps per SM = 8 per SMSP
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8);
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr_per_smsp = 8.0 * ITERS * CHAINS * ops_per_iter;  // 8 warps per SMSP
    const double total_thread_ops = (double)blocks * threads * ITERS * CHAINS * ops_per_iter;
    (void)warp_instr_per_smsp; (void)c;
    const double ipc = total_thread_ops / 32.0 / (ms * 1e-3) / (mhz * 1e6) / (sms * 4.0);
    printf("%-26s %8.2f Tops/s   %6.3f warp-instr/clk/SMSP at %.0f MHz (CUDA events, %0.3f ms)\n", name, total_thread_ops / (ms * 1e-3) / 1e12, ipc, mhz, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk / 1000);
    const int sms = p.multiProcessorCount; const double mhz = clk / 1000.0;
    run<0>("IMAD.WIDE.U32", 1, sms, mhz);
    run<1>("IMAD (lo)", 1, sms, mhz);
    run<2>("IMAD.HI", 1, sms, mhz);
    run<3>("IADD (add.u32)", 1, sms, mhz);
    run<4>("LOP3", 1, sms, mhz);
    run<5>("SHF", 1, sms, mhz);
    run<6>("PRMT", 1, sms, mhz);
    run<7>("IADD3 + IADD3.X pair", 2, sms, mhz);
    run<8>("LOP3 + IMAD (3 src each)", 2, sms, mhz);
    run<9>("IMAD.WIDE + IADD3 + .X", 3, sms, mhz);
    run<10>("SHF + IMUL (2 src each)", 2, sms, mhz);
    return 0;
}
