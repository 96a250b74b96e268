// Warp-level tensor-core micro-benchmark for sm_100a (mma.sync, SASS HMMA): achieved MACs/clk/SM and TFLOP/s of
//   (a) m16n8k16 bf16 -> f32, NACC independent accumulators per warp,
//   (b) the same with R independent FFMAs issued per MMA (does the FMA pipe overlap the tensor pipe?),
//   (c) m16n8k8 tf32 -> f32.
// Register-only; no memory traffic in the timed loop.  Fixes the denominator for a tensor-core oscillator bank.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bin/hmma_peak tools/microbench/hmma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0: bf16 k16; MODE 1: tf32 k8.  R FFMAs (independent chains) per MMA.
template <int NACC, int R, int MODE>
__global__ void k_mma(float* out, int iters, float fa, float fb, long long* cyc) {
    float d[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = 0x3c003c00u + threadIdx.x + i;     // small bf16 pairs
    for (int i = 0; i < 2; i++) b[i] = 0x3c003c00u + 2 * threadIdx.x + i;
    for (int n = 0; n < NACC; n++) for (int i = 0; i < 4; i++) d[n][i] = 0.f;
    float x[R > 0 ? R * 4 : 1];
    for (int i = 0; i < (R > 0 ? R * 4 : 1); i++) x[i] = threadIdx.x * 1e-3f + i;
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int n = 0; n < NACC; n++) {
            if (MODE == 0) mma_bf16(d[n], a, b); else mma_tf32(d[n], a, b);
#pragma unroll
            for (int r = 0; r < R; r++) x[(n % 4) * (R > 0 ? R : 1) + r] = fmaf(x[(n % 4) * (R > 0 ? R : 1) + r], fa, fb);
        }
    }
    long long c1 = clock64();
    float s = 0;
    for (int n = 0; n < NACC; n++) for (int i = 0; i < 4; i++) s += d[n][i];
    for (int i = 0; i < (R > 0 ? R * 4 : 1); i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int NACC, int R, int MODE>
int run(const char* name, int ctas_per_sm, int warps, int iters, float* d_out, long long* d_cyc) {
    int dev = 0; cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    const int grid = pr.multiProcessorCount * ctas_per_sm, block = warps * 32;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_mma<NACC, R, MODE><<<grid, block>>>(d_out, 10, 1.0001f, 1e-6f, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_mma<NACC, R, MODE><<<grid, block>>>(d_out, iters, 1.0001f, 1e-6f, d_cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc = 0; CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    const double macs_per_mma = MODE == 0 ? 16.0 * 8 * 16 : 16.0 * 8 * 8;
    const double mmas = (double)grid * warps * iters * NACC;
    const double tflops = 2.0 * mmas * macs_per_mma / (ms * 1e-3) / 1e12;
    const double mac_clk_sm = (double)ctas_per_sm * warps * iters * NACC * macs_per_mma / (double)cyc;
    const double cyc_per_mma_smsp = (double)cyc / ((double)ctas_per_sm * warps * iters * NACC / 4.0);
    printf("{\"case\": \"%s\", \"nacc\": %d, \"ffma_per_mma\": %d, \"ctas_per_sm\": %d, \"warps\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"mac_per_clk_sm\": %.0f, \"cycles_per_mma_per_smsp\": %.2f}\n",
           name, NACC, R, ctas_per_sm, warps, ms, tflops, mac_clk_sm, cyc_per_mma_smsp);
    return 0;
}

int main() {
    float* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 148 * 8 * 1024 * sizeof(float) * 4));
    CK(cudaMalloc(&d_cyc, 148 * 64 * sizeof(long long)));
    const int it = 20000;
    if (run<8, 0, 0>("bf16 m16n8k16", 1, 4, it, d_out, d_cyc)) return 1;
    if (run<8, 0, 0>("bf16 m16n8k16", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<8, 0, 0>("bf16 m16n8k16", 2, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 0, 0>("bf16 m16n8k16", 1, 4, it, d_out, d_cyc)) return 1;
    if (run<16, 0, 0>("bf16 m16n8k16", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<32, 0, 0>("bf16 m16n8k16", 1, 4, it / 2, d_out, d_cyc)) return 1;
    if (run<32, 0, 0>("bf16 m16n8k16", 1, 8, it / 2, d_out, d_cyc)) return 1;
    if (run<16, 1, 0>("bf16 + ffma", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 2, 0>("bf16 + ffma", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 3, 0>("bf16 + ffma", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 4, 0>("bf16 + ffma", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 6, 0>("bf16 + ffma", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 4, 0>("bf16 + ffma", 1, 4, it, d_out, d_cyc)) return 1;
    if (run<8, 0, 1>("tf32 m16n8k8", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 0, 1>("tf32 m16n8k8", 1, 8, it, d_out, d_cyc)) return 1;
    if (run<16, 0, 1>("tf32 m16n8k8", 2, 8, it, d_out, d_cyc)) return 1;
    return 0;
}
