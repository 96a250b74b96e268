// FP32 FMA-pipe micro-benchmark for sm_100a: measures achieved FMA lanes/clk/SM for
//   (a) scalar FFMA with 3 register sources, (b) packed fma.rn.f32x2 (FFMA2),
//   (c) the oscillator inner loop candidates (magic-circle step, scalar and packed).
// Register-only; no memory traffic in the timed loop.  Used to fix the K1 roofline denominator.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

template <int CH>
__global__ void k_ffma(float* out, int iters, float a, float b, long long* cyc) {
    float x[CH];
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-3f + i;
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) x[i] = fmaf(x[i], a, b);
    }
    long long c1 = clock64();
    float s = 0; for (int i = 0; i < CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int CH>
__global__ void k_ffma2(float* out, int iters, float a, float b, long long* cyc) {
    uint64_t x[CH];
    uint64_t a2 = pack2(a, a * 1.0001f), b2 = pack2(b, b * 0.999f);
    for (int i = 0; i < CH; i++) x[i] = pack2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) x[i] = fma2(x[i], a2, b2);
    }
    long long c1 = clock64();
    float s = 0; for (int i = 0; i < CH; i++) { float p, q; unpack2(x[i], p, q); s += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// magic-circle oscillator, scalar: per sample  x -= eps*y ; y += eps*x ; e += e*rm1 ; acc_j += e*y   (4 FMA-pipe ops)
template <int NP, int T>
__global__ void k_osc_scalar(float* out, int iters, float eps0, float rm1, long long* cyc) {
    float x[NP], y[NP], e[NP], eps[NP];
    for (int p = 0; p < NP; p++) { x[p] = 1.f; y[p] = 0.f; e[p] = 1.f; eps[p] = eps0 * (1 + p + threadIdx.x); }
    float acc[T];
    for (int j = 0; j < T; j++) acc[j] = 0.f;
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < T; j++) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                x[p] = fmaf(-eps[p], y[p], x[p]);
                y[p] = fmaf(eps[p], x[p], y[p]);
                e[p] = fmaf(e[p], rm1, e[p]);
                acc[j] = fmaf(e[p], y[p], acc[j]);
            }
        }
    }
    long long c1 = clock64();
    float s = 0; for (int j = 0; j < T; j++) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// packed: each f32x2 register holds two partials
template <int NP, int T>
__global__ void k_osc_packed(float* out, int iters, float eps0, float rm1, long long* cyc) {
    uint64_t x[NP], y[NP], e[NP], eps[NP], neps[NP], r2[NP];
    for (int p = 0; p < NP; p++) {
        x[p] = pack2(1.f, 1.f); y[p] = pack2(0.f, 0.f); e[p] = pack2(1.f, 1.f);
        float e0 = eps0 * (1 + 2 * p + threadIdx.x), e1 = eps0 * (2 + 2 * p + threadIdx.x);
        eps[p] = pack2(e0, e1); neps[p] = pack2(-e0, -e1); r2[p] = pack2(rm1, rm1 * 1.01f);
    }
    uint64_t acc[T];
    for (int j = 0; j < T; j++) acc[j] = pack2(0.f, 0.f);
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < T; j++) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                x[p] = fma2(neps[p], y[p], x[p]);
                y[p] = fma2(eps[p], x[p], y[p]);
                e[p] = fma2(e[p], r2[p], e[p]);
                acc[j] = fma2(e[p], y[p], acc[j]);
            }
        }
    }
    long long c1 = clock64();
    float s = 0; for (int j = 0; j < T; j++) { float p, q; unpack2(acc[j], p, q); s += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <typename F>
int run(const char* name, F launch, int blocks, int threads, double fma_per_thread, float* d_out, long long* d_cyc) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc[4096]; CK(cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double mc = 0; for (int i = 0; i < blocks; i++) mc += (double)cyc[i]; mc /= blocks;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double total = fma_per_thread * (double)blocks * threads;
    double per_clk_sm = total / mc / sms;          // assumes blocks spread evenly, one wave
    printf("%-28s blocks=%4d thr=%4d  %.3f ms  %.3e FMA/s  mean_cyc=%.0f  FMA/clk/SM=%.1f  implied_MHz=%.0f\n",
           name, blocks, threads, ms, total / (ms * 1e-3), mc, per_clk_sm, mc / (ms * 1e-3) / 1e6);
    return 0;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sms=%d clock=%d kHz\n", prop.name, sms, prop.clockRate);
    float* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(float) * 4096 * 1024)); CK(cudaMalloc(&d_cyc, sizeof(long long) * 4096));
    const int iters = 4000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mult : {1, 2}) {
            int blocks = sms * mult;
            if (threads * mult > 2048) continue;
            run("ffma_3reg_ch8", [&] { k_ffma<8><<<blocks, threads>>>(d_out, iters, 1.0001f, 0.5f, d_cyc); }, blocks, threads, 8.0 * 8 * iters, d_out, d_cyc);
            run("ffma2_ch8", [&] { k_ffma2<8><<<blocks, threads>>>(d_out, iters, 1.0001f, 0.5f, d_cyc); }, blocks, threads, 2.0 * 8 * 8 * iters, d_out, d_cyc);
        }
    }
    for (int threads : {128, 256, 384, 512}) {
        int blocks = sms;
        run("osc_scalar_np2_T32", [&] { k_osc_scalar<2, 32><<<blocks, threads>>>(d_out, iters / 8, 1e-3f, -1e-4f, d_cyc); }, blocks, threads, 4.0 * 2 * 32 * (iters / 8), d_out, d_cyc);
        run("osc_scalar_np4_T32", [&] { k_osc_scalar<4, 32><<<blocks, threads>>>(d_out, iters / 8, 1e-3f, -1e-4f, d_cyc); }, blocks, threads, 4.0 * 4 * 32 * (iters / 8), d_out, d_cyc);
        run("osc_packed_np1_T32", [&] { k_osc_packed<1, 32><<<blocks, threads>>>(d_out, iters / 8, 1e-3f, -1e-4f, d_cyc); }, blocks, threads, 8.0 * 1 * 32 * (iters / 8), d_out, d_cyc);
        run("osc_packed_np2_T32", [&] { k_osc_packed<2, 32><<<blocks, threads>>>(d_out, iters / 8, 1e-3f, -1e-4f, d_cyc); }, blocks, threads, 8.0 * 2 * 32 * (iters / 8), d_out, d_cyc);
        if (threads <= 256) run("osc_packed_np2_T64", [&] { k_osc_packed<2, 64><<<blocks, threads>>>(d_out, iters / 16, 1e-3f, -1e-4f, d_cyc); }, blocks, threads, 8.0 * 2 * 64 * (iters / 16), d_out, d_cyc);
    }
    // sustained: 2 s of packed FMA to see the clock under load
    {
        int blocks = sms * 2, threads = 512;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < 200; r++) k_ffma2<8><<<blocks, threads>>>(d_out, iters * 4, 1.0001f, 0.5f, d_cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double total = 200.0 * 2 * 8 * 8 * iters * 4 * blocks * threads;
        printf("sustained ffma2: %.1f ms  %.3e FMA/s (%.1f TFLOP/s)\n", ms, total / (ms * 1e-3), 2 * total / (ms * 1e-3) / 1e12);
    }
    return 0;
}
