// What the operand generation of K1T costs per instruction class on sm_100a (register-only loops, 8 warps x 2 CTAs per SM):
//   split   = gm_split (cvt.rn.f16x2.f32, 2 x f16->f32, 2 FADD, cvt.rn.f16x2.f32)   rot = complex rotation (2 FMUL + 2 FFMA)
//   f2fp    = cvt.rn.f16x2.f32 alone          up = the two f16 -> f32 conversions alone
//   anchor  = gm_anchor (umulhi/imad, I2F, FMUL, 2 MUFU sin/cos, MUFU ex2, 3 FMUL)
// Prints warp-instructions (or calls) per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I libfriendship_b200/csrc -I include -o build/bin/gen_peak tools/microbench/gen_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void split(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void anchor(unsigned inc_lo, unsigned inc_hi, unsigned ph0_hi, float kappa, float amp, unsigned long long n, float& s, float& c) {
    const unsigned n_lo = (unsigned)n, n_hi = (unsigned)(n >> 32);
    const unsigned turns_hi = __umulhi(inc_lo, n_lo) + inc_hi * n_lo + inc_lo * n_hi + ph0_hi;
    const float th = (float)(int)turns_hi * 1.4629180792671596e-9f;
    __sincosf(th, &s, &c);
    float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-kappa * (float)n));
    e *= amp; s *= e; c *= e;
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters, float fa, float fb, long long* cyc) {
    float x[8], y[8];
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = 1.f - x[i] * 0.5f; }
    uint32_t acc = 0;
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { uint32_t h, l; split(x[i], y[i], h, l); acc ^= h + l; x[i] += fa; }
            if (MODE == 1) { const float nx = fmaf(x[i], fa, y[i] * fb), ny = fmaf(y[i], fa, -x[i] * fb); x[i] = nx; y[i] = ny; }
            if (MODE == 2) { const __half2 h = __floats2half2_rn(x[i], y[i]); acc ^= *reinterpret_cast<const uint32_t*>(&h); x[i] += fa; }
            if (MODE == 3) { uint32_t u = __float_as_uint(x[i]) ^ acc; const float2 f = __half22float2(*reinterpret_cast<__half2*>(&u)); x[i] += f.x; y[i] += f.y; }
            if (MODE == 4) { float s, c; anchor(__float_as_uint(x[i]), 77u + i, 5u, fb, fa, (unsigned long long)(it * 8 + i), s, c); x[i] += s; y[i] += c; }
            if (MODE == 5) { uint32_t h, l; split(x[i], y[i], h, l); acc ^= h + l; const float nx = fmaf(x[i], fa, y[i] * fb), ny = fmaf(y[i], fa, -x[i] * fb); x[i] = nx; y[i] = ny; }
        }
    }
    long long c1 = clock64();
    float s = 0; for (int i = 0; i < 8; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int MODE>
int run(const char* name, int iters, float* d_out, long long* d_cyc) {
    const int grid = 148 * 2, block = 256;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<grid, block>>>(d_out, 10, 1.0001f, 1e-3f, d_cyc); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k<MODE><<<grid, block>>>(d_out, iters, 1.0001f, 1e-3f, d_cyc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc = 0; CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    const double calls_per_sm = 2.0 * 8 * iters * 8;            // warp-level calls per SM (2 CTAs x 8 warps x iters x 8)
    printf("{\"case\": \"%s\", \"ms\": %.3f, \"cycles_per_warp_call_per_sm\": %.3f, \"thread_calls_per_clk_per_sm\": %.2f}\n", name, ms, (double)cyc / calls_per_sm, calls_per_sm * 32 / (double)cyc);
    return 0;
}
int main() {
    float* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 148 * 2 * 256 * sizeof(float))); CK(cudaMalloc(&d_cyc, 148 * 2 * sizeof(long long)));
    const int it = 4000;
    if (run<0>("split (2 cvt.f16x2 + 2 up + 2 FADD)", it, d_out, d_cyc)) return 1;
    if (run<1>("rotation (2 FMUL + 2 FFMA)", it, d_out, d_cyc)) return 1;
    if (run<2>("cvt.rn.f16x2.f32 alone", it, d_out, d_cyc)) return 1;
    if (run<3>("2 x f16 -> f32 alone", it, d_out, d_cyc)) return 1;
    if (run<4>("anchor (2 MUFU sincos + ex2)", it / 4, d_out, d_cyc)) return 1;
    if (run<5>("split + rotation", it, d_out, d_cyc)) return 1;
    return 0;
}
