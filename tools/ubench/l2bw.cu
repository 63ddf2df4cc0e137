// Memory-side roofs for K1: what the L2 / HBM hierarchy delivers for the two access shapes the
// traversal kernel uses, so that roofline.l2 in bench.py has a MEASURED denominator on this GPU.
//   stream   every lane reads consecutive 32-byte sectors (ld.global.cg.v8, bypassing L1) over a working set of W bytes,
//            repeatedly: W <= ~100 MB stays L2-resident (L2 -> SM bandwidth), W >> 126 MB is the HBM copy-read figure.
//   gather   every lane reads RECORDS of R bytes (64 = one quad / triangle record, 128 = one 8-wide node) at hashed,
//            independent addresses inside W — the shape of a node fetch by 32 divergent rays; G loads in flight per lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/l2bw.bin tools/ubench/l2bw.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld256(const void *p, uint32_t &acc) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    acc += a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ void ld256_nc(const void *p, uint32_t &acc) {  // through L1 (the path K1 uses)
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    acc += a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

__global__ void __launch_bounds__(256) k_stream(const char *buf, size_t sectors, int reps, uint32_t *out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < sectors; i += stride) ld256(buf + i * 32, acc);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int R, bool L1>
__global__ void __launch_bounds__(256) k_gather(const char *buf, uint32_t n_records, int iters, uint32_t *out) {
    uint32_t acc = 0;
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < 4; g++) {  // 4 independent records in flight per lane
            x = x * 1664525u + 1013904223u;
            const uint32_t rec = (uint32_t)(((uint64_t)(x ^ (x >> 15)) * n_records) >> 32);
            const char *p = buf + (size_t)rec * R;
#pragma unroll
            for (int s = 0; s < R / 32; s++) {
                if (L1) ld256_nc(p + s * 32, acc);
                else ld256(p + s * 32, acc);
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) {
    float ms;
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const size_t cap = (size_t)2 << 30;
    char *buf;
    uint32_t *out;
    cudaMalloc(&buf, cap);
    cudaMalloc(&out, 4);
    cudaMemset(buf, 1, cap);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    printf("{\"sms\": %d, \"clock_khz\": %d, \"results\": [\n", sms, clk);
    const size_t sets[] = {(size_t)16 << 20, (size_t)48 << 20, (size_t)96 << 20, (size_t)112 << 20, (size_t)432 << 20, (size_t)1400 << 20};
    bool first = true;
    for (size_t W : sets) {
        const size_t sectors = W / 32;
        const int reps = (int)(((size_t)8 << 30) / W) + 1;
        k_stream<<<sms * 8, 256>>>(buf, sectors, 1, out);
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            k_stream<<<sms * 8, 256>>>(buf, sectors, reps, out);
            cudaEventRecord(e1);
            best = fminf(best, time_ms(e0, e1));
        }
        printf("%s{\"shape\": \"stream\", \"set_mb\": %zu, \"gbs\": %.1f}", first ? "" : ",\n", W >> 20, (double)W * reps / (best * 1e-3) / 1e9);
        first = false;
    }
    auto gather = [&](auto kernel, int R, bool l1, size_t W) {
        const uint32_t n = (uint32_t)(W / R);
        const int iters = 256;
        kernel<<<sms * 8, 256>>>(buf, n, 8, out);
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            kernel<<<sms * 8, 256>>>(buf, n, iters, out);
            cudaEventRecord(e1);
            best = fminf(best, time_ms(e0, e1));
        }
        const double recs = (double)sms * 8 * 256 * iters * 4;
        printf(",\n{\"shape\": \"gather\", \"record_bytes\": %d, \"l1\": %s, \"set_mb\": %zu, \"grecords_s\": %.2f, \"gbs\": %.1f}", R, l1 ? "true" : "false", W >> 20,
               recs / (best * 1e-3) / 1e9, recs * R / (best * 1e-3) / 1e9);
    };
    for (size_t W : sets) {
        gather(k_gather<64, false>, 64, false, W);
        gather(k_gather<64, true>, 64, true, W);
        gather(k_gather<128, true>, 128, true, W);
    }
    printf("\n]}\n");
    return 0;
}
