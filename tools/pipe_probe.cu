// Do carry-chained integer instructions on the ALU pipe (IADD3 / IADD3.X) and on the FMA pipe (IMAD.WIDE[.X] with carry out / in)
// overlap?  Independent chains per thread, 8 warps per SM sub-partition; reports warp-instructions per clock per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu && ./tools/pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define CH 6
template <int OP>
__global__ void k(uint32_t* out, uint32_t a0, uint32_t b0) {
    uint32_t x[CH], y[CH], lo[CH], hi[CH], u[CH], v[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { x[c] = a0 + threadIdx.x + c; y[c] = b0 ^ (c * 77u); lo[c] = x[c]; hi[c] = y[c]; u[c] = x[c] * 3; v[c] = y[c] * 5; }
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (OP == 0 || OP == 2 || OP == 4)   // ALU carry pair: IADD3 (carry out) + IADD3.X (carry in)
                asm volatile("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, %3;" : "+r"(x[c]), "+r"(y[c]) : "r"(a0), "r"(b0));
            if (OP == 1 || OP == 2)              // FMA carry pair: mad.lo.cc + madc.hi -> IMAD.WIDE with carry out, then a carry-in consumer
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\t madc.hi.cc.u32 %1, %2, %3, %1;\n\t addc.u32 %4, %4, 0;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(u[c]), "r"(v[c]), "r"(x[c]));
            if (OP == 3 || OP == 4)              // FMA without carries: mad.wide
            {
                unsigned long long w = ((unsigned long long)hi[c] << 32) | lo[c];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(u[c]), "r"(v[c]));
                lo[c] = (uint32_t)w; hi[c] = (uint32_t)(w >> 32);
            }
            if (OP == 5) {                       // two independent IMAD.WIDE.X-style chains (what a multiplier row looks like)
                asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\t madc.hi.cc.u32 %1, %4, %5, %1;\n\t madc.lo.cc.u32 %2, %4, %6, %2;\n\t madc.hi.u32 %3, %4, %6, %3;"
                             : "+r"(lo[c]), "+r"(hi[c]), "+r"(x[c]), "+r"(y[c]) : "r"(u[c]), "r"(v[c]), "r"(a0));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= x[c] ^ y[c] ^ lo[c] ^ hi[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP> void run(const char* name, double instr_per_iter, int sms, double mhz) {
    uint32_t* out;
    const int blocks = sms * 2, threads = 512;
    cudaMalloc(&out, blocks * threads * 4);
    k<OP><<<blocks, threads>>>(out, 3, 5);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)blocks * threads * ITERS * CH * instr_per_iter;
    // 8 warps per sub-partition each run ITERS * CH chain steps: issue cycles one chain step costs the sub-partition
    printf("%-58s %6.2f clk per chain step per SMSP   (%6.3f warp-instr/clk/SMSP at the nominal %.0f instr per step; %.3f ms)\n", name,
           (ms * 1e-3 * mhz * 1e6) / ((double)ITERS * CH * 8), total / 32.0 / (ms * 1e-3) / (mhz * 1e6) / (sms * 4.0), instr_per_iter, ms);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount; const double mhz = clk / 1000.0;
    run<0>("ALU carry pair (IADD3 + IADD3.X)", 2, sms, mhz);
    run<1>("FMA carry pair (IMAD.WIDE cc) + carry-in add", 2, sms, mhz);
    run<2>("both, independent registers", 4, sms, mhz);
    run<3>("IMAD.WIDE without carries", 1, sms, mhz);
    run<4>("ALU carry pair + carry-free IMAD.WIDE, independent", 3, sms, mhz);
    run<5>("IMAD.WIDE cc -> IMAD.WIDE.X chain (multiplier row)", 2, sms, mhz);
    return 0;
}
