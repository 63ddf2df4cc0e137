// Pipe-throughput microbenchmark for the K1 node step: which SASS ops share the (half-rate) ALU pipe and which run
// on the FMA pipes.  Each kernel runs ITER x 8 independent chains of ONE op per thread, full machine, 1024 threads/SM x 2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes.bin tools/ubench/pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
#define CHAINS 8
enum { OP_PRMT, OP_IMADHI, OP_IMAD, OP_FFMA, OP_FMNMX, OP_SEL, OP_VIMNMX, OP_LOP3, OP_MIX_PRMT_FFMA, OP_MIX_IMADHI_FFMA, OP_MIX_PRMT_IMADHI, OP_ISETP_SEL, OP_HADD2F32, OP_MIX_HADD2F32_FFMA, OP_MIX_HADD2F32_PRMT, N_OPS };
const char *names[] = {"PRMT", "IMAD.HI(c)", "IMAD(c)", "FFMA", "FMNMX", "SEL(pred)", "VIMNMX", "LOP3", "PRMT+FFMA 1:1", "IMAD.HI+FFMA 1:1", "PRMT+IMAD.HI 1:1", "ISETP+SEL", "HADD2.F32 (cvt.f32.f16)", "HADD2.F32+FFMA 1:1", "HADD2.F32+PRMT 1:1"};
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t c256, uint32_t magic, float fs, int iters) {
    uint32_t x[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x, f[i] = (float)x[i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7652;" : "+r"(x[i]) : "r"(magic));
            if (OP == OP_IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c256), "r"(magic));
            if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c256), "r"(magic));
            if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fs));
            if (OP == OP_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fs));
            if (OP == OP_SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %0, %2, p;}" : "+r"(x[i]) : "r"(c256), "r"(magic));
            if (OP == OP_VIMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(magic));
            if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(magic), "r"(c256));
            if (OP == OP_MIX_PRMT_FFMA) {
                asm volatile("prmt.b32 %0, %0, %1, 0x7652;" : "+r"(x[i]) : "r"(magic));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fs));
            }
            if (OP == OP_MIX_IMADHI_FFMA) {
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c256), "r"(magic));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fs));
            }
            if (OP == OP_MIX_PRMT_IMADHI) {
                asm volatile("prmt.b32 %0, %0, %1, 0x7652;" : "+r"(x[i]) : "r"(magic));
                uint32_t y = __float_as_uint(f[i]);
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(y) : "r"(c256), "r"(magic));
                f[i] = __uint_as_float(y);
            }
            if (OP == OP_HADD2F32 || OP == OP_MIX_HADD2F32_FFMA || OP == OP_MIX_HADD2F32_PRMT)
                asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, lo;}" : "+r"(x[i]));
            if (OP == OP_MIX_HADD2F32_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fs));
            if (OP == OP_MIX_HADD2F32_PRMT) {
                uint32_t y = __float_as_uint(f[i]);
                asm volatile("prmt.b32 %0, %0, %1, 0x7652;" : "+r"(y) : "r"(magic));
                f[i] = __uint_as_float(y);
            }
            if (OP == OP_ISETP_SEL) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.u32 %0, %0, %2, p;}" : "+r"(x[i]) : "r"(c256), "r"(magic));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc += x[i] + __float_as_uint(f[i]);
    if (acc == 0x12345678u) out[0] = acc;
}
template <int OP>
void run(uint32_t *d, int sms) {
    const int grid = sms * 8, block = 256;
    int per_iter = ((OP >= OP_MIX_PRMT_FFMA && OP <= OP_MIX_PRMT_IMADHI) || OP >= OP_MIX_HADD2F32_FFMA) ? 2 : (OP == OP_SEL || OP == OP_ISETP_SEL ? 2 : 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<OP><<<grid, block>>>(d, 256u, 0x4B000000u, 1.0000001f, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<OP><<<grid, block>>>(d, 256u, 0x4B000000u, 1.0000001f, ITER);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double warp_instr = (double)grid * (block / 32) * ITER * CHAINS * per_iter;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = best * 1e-3 * clk * 1e3;
    printf("%-20s %8.3f ms  %.2f warp-instr/cycle/SM (%d SASS op(s) per chain step; nominal clock %d kHz)\n", names[OP], best, warp_instr / cycles / sms, per_iter, clk);
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *d;
    cudaMalloc(&d, 4);
    run<OP_PRMT>(d, sms); run<OP_IMADHI>(d, sms); run<OP_IMAD>(d, sms); run<OP_FFMA>(d, sms); run<OP_FMNMX>(d, sms); run<OP_SEL>(d, sms);
    run<OP_VIMNMX>(d, sms); run<OP_LOP3>(d, sms); run<OP_MIX_PRMT_FFMA>(d, sms); run<OP_MIX_IMADHI_FFMA>(d, sms); run<OP_MIX_PRMT_IMADHI>(d, sms); run<OP_ISETP_SEL>(d, sms); run<OP_HADD2F32>(d, sms); run<OP_MIX_HADD2F32_FFMA>(d, sms); run<OP_MIX_HADD2F32_PRMT>(d, sms);
    return 0;
}
