// Inner-loop candidates for the K1 oscillator bank (damped lifting-form resonator, 4 FMA-pipe ops / partial-sample):
//    x  = fma(-a, y, x);  t = fma(cm1, y, y);  y = fma(b, x, t);  acc[j] += y
// Lanes map to time phases, so the per-partial coefficients (a, b, cm1) are the same for every lane.
// Variants differ in where the coefficients live (vector regs from smem, constant bank -> uniform regs) and in
// how many independent time phases R a thread carries (operand-reuse-cache hits on the shared coefficient).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__constant__ float4 c_coef[1024];

// coefficients in shared memory (vector registers after LDS.128), R phases per thread, state in registers
template <int R, int T, int KP>
__global__ void __launch_bounds__(256) k_smem(float* out, const float4* __restrict__ coef, int tiles, long long* cyc) {
    __shared__ float4 s_coef[KP];
    for (int i = threadIdx.x; i < KP; i += blockDim.x) s_coef[i] = coef[i];
    __syncthreads();
    float x[KP][R], y[KP][R];
#pragma unroll
    for (int k = 0; k < KP; k++)
#pragma unroll
        for (int r = 0; r < R; r++) { x[k][r] = 1.f + threadIdx.x * 1e-3f + r; y[k][r] = 0.1f * k; }
    float tot = 0.f;
    long long c0 = clock64();
    for (int tile = 0; tile < tiles; tile++) {
        float acc[R][T];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < T; j++) acc[r][j] = 0.f;
#pragma unroll
        for (int k = 0; k < KP; k++) {
            float4 c = s_coef[k];
#pragma unroll
            for (int j = 0; j < T; j++) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    x[k][r] = fmaf(-c.x, y[k][r], x[k][r]);
                    float t = fmaf(c.z, y[k][r], y[k][r]);
                    y[k][r] = fmaf(c.y, x[k][r], t);
                    acc[r][j] += y[k][r];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < T; j++) tot += acc[r][j];
    }
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// state in shared memory (per thread column), partial loop NOT unrolled: small code, arbitrary K
template <int R, int T>
__global__ void __launch_bounds__(256) k_smem_state(float* out, const float4* __restrict__ coef, int K, int tiles, long long* cyc) {
    extern __shared__ float2 s_state[];   // [K][R][256]
    __shared__ float4 s_coef[64];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_coef[i] = coef[i];
    for (int k = 0; k < K; k++)
        for (int r = 0; r < R; r++) s_state[(k * R + r) * 256 + threadIdx.x] = make_float2(1.f + threadIdx.x * 1e-3f + r, 0.1f * k);
    __syncthreads();
    float tot = 0.f;
    long long c0 = clock64();
    for (int tile = 0; tile < tiles; tile++) {
        float acc[R][T];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < T; j++) acc[r][j] = 0.f;
        for (int k = 0; k < K; k++) {
            float4 c = s_coef[k];
            float x[R], y[R];
#pragma unroll
            for (int r = 0; r < R; r++) { float2 s = s_state[(k * R + r) * 256 + threadIdx.x]; x[r] = s.x; y[r] = s.y; }
#pragma unroll
            for (int j = 0; j < T; j++) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    x[r] = fmaf(-c.x, y[r], x[r]);
                    float t = fmaf(c.z, y[r], y[r]);
                    y[r] = fmaf(c.y, x[r], t);
                    acc[r][j] += y[r];
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) s_state[(k * R + r) * 256 + threadIdx.x] = make_float2(x[r], y[r]);
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < T; j++) tot += acc[r][j];
    }
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// coefficients from the constant bank with a uniform index (-> ULDC / uniform registers), state in smem
template <int R, int T>
__global__ void __launch_bounds__(256) k_const_state(float* out, int K, int tiles, long long* cyc) {
    extern __shared__ float2 s_state[];
    for (int k = 0; k < K; k++)
        for (int r = 0; r < R; r++) s_state[(k * R + r) * 256 + threadIdx.x] = make_float2(1.f + threadIdx.x * 1e-3f + r, 0.1f * k);
    __syncthreads();
    float tot = 0.f;
    long long c0 = clock64();
    for (int tile = 0; tile < tiles; tile++) {
        float acc[R][T];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < T; j++) acc[r][j] = 0.f;
        for (int k = 0; k < K; k++) {
            const float ca = c_coef[k].x, cb = c_coef[k].y, cc = c_coef[k].z;
            float x[R], y[R];
#pragma unroll
            for (int r = 0; r < R; r++) { float2 s = s_state[(k * R + r) * 256 + threadIdx.x]; x[r] = s.x; y[r] = s.y; }
#pragma unroll
            for (int j = 0; j < T; j++) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    x[r] = fmaf(-ca, y[r], x[r]);
                    float t = fmaf(cc, y[r], y[r]);
                    y[r] = fmaf(cb, x[r], t);
                    acc[r][j] += y[r];
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) s_state[(k * R + r) * 256 + threadIdx.x] = make_float2(x[r], y[r]);
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int j = 0; j < T; j++) tot += acc[r][j];
    }
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <typename F>
int run(const char* name, F launch, int blocks, int threads, double fma_per_thread, long long* d_cyc) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    static long long cyc[4096]; CK(cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double mc = 0; for (int i = 0; i < blocks; i++) mc += (double)cyc[i]; mc /= blocks;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double total = fma_per_thread * (double)blocks * threads;
    printf("%-34s blocks=%4d thr=%4d  %.3f ms  %.3e FMA/s  FMA/clk/SM=%.1f (of 128)\n",
           name, blocks, threads, ms, total / (ms * 1e-3), total / mc / sms * (blocks > sms ? (double)sms / blocks * ((blocks + sms - 1) / sms) : 1.0));
    return 0;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* d_out; long long* d_cyc; float4* d_coef;
    CK(cudaMalloc(&d_out, sizeof(float) * 4096 * 1024)); CK(cudaMalloc(&d_cyc, sizeof(long long) * 4096));
    float4 h_coef[1024];
    for (int i = 0; i < 1024; i++) h_coef[i] = make_float4(1e-3f * (i + 1), 1e-3f * (i + 1), -2e-4f, 0.f);
    CK(cudaMalloc(&d_coef, sizeof(h_coef))); CK(cudaMemcpy(d_coef, h_coef, sizeof(h_coef), cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_coef, h_coef, sizeof(h_coef)));
    const int tiles = 200;
    int blocks = sms;
    run("smem_coef R1 T32 KP8 regstate", [&] { k_smem<1, 32, 8><<<blocks, 256>>>(d_out, d_coef, tiles, d_cyc); }, blocks, 256, 4.0 * 1 * 32 * 8 * tiles, d_cyc);
    run("smem_coef R2 T32 KP8 regstate", [&] { k_smem<2, 32, 8><<<blocks, 256>>>(d_out, d_coef, tiles, d_cyc); }, blocks, 256, 4.0 * 2 * 32 * 8 * tiles, d_cyc);
    run("smem_coef R4 T16 KP8 regstate", [&] { k_smem<4, 16, 8><<<blocks, 256>>>(d_out, d_coef, tiles, d_cyc); }, blocks, 256, 4.0 * 4 * 16 * 8 * tiles, d_cyc);
    const int K = 32;
#define RUNSTATE(R, T, THR) { size_t sm = (size_t)K * R * 256 * sizeof(float2); \
        CK(cudaFuncSetAttribute(k_smem_state<R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        run("smem_coef R" #R " T" #T " smemstate", [&] { k_smem_state<R, T><<<blocks, THR, sm>>>(d_out, d_coef, K, tiles, d_cyc); }, blocks, THR, 4.0 * R * T * K * tiles, d_cyc); \
        CK(cudaFuncSetAttribute(k_const_state<R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        run("const_coef R" #R " T" #T " smemstate", [&] { k_const_state<R, T><<<blocks, THR, sm>>>(d_out, K, tiles, d_cyc); }, blocks, THR, 4.0 * R * T * K * tiles, d_cyc); }
    RUNSTATE(1, 32, 256)
    RUNSTATE(1, 64, 256)
    RUNSTATE(2, 32, 256)
    RUNSTATE(2, 64, 256)
    RUNSTATE(3, 32, 256)
    RUNSTATE(4, 16, 256)
    RUNSTATE(4, 32, 128)
    return 0;
}
