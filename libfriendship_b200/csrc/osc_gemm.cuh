// osc_gemm.cuh — K1G: the oscillator bank of big voices as a matrix product on the tensor cores (included by osc.cu).
//
// After the attack ramps (t >= max_attack) a partial is  amp * rho^t * sin(w t + phi).  Cut time into blocks of 128
// samples, t = 128 b + j:
//     amp rho^t sin(w t + phi) = [E_b sin(T_b)] * [rho^j cos(w j)] + [E_b cos(T_b)] * [rho^j sin(w j)],
//     E_b = amp rho^(128 b),  T_b = 128 b w + phi,
// so a voice is  out[b, j] = sum_k A[b, k] W[k, j]  with k = (partial, sin | cos): a GEMM with M = blocks, N = 128,
// K = 2 x partials — 2 K flops per sample instead of the resonator's 4 FMA-pipe instructions per partial-sample, and on the
// tensor pipe.  Neither operand ever exists in memory: every thread GENERATES the fragments of A and W it owns in the
// mma.sync register layout, by complex rotation (4 FMA-pipe instructions per element pair) from anchors evaluated the way
// osc_group does it (phase in 64-bit fixed-point turns = exact range reduction, MUFU sin/cos/ex2).
//   * arithmetic: fp16 x fp16 -> fp32 (mma.sync.m16n8k16, SASS HMMA.16816.F32).  fp16 alone is 11 bits, so every operand
//     is split x = hi + lo (two fp16: 22 bits) and a step issues lo*hi + hi*lo + hi*hi (the lo*lo term is 2^-22 of the
//     product).  fp16 has a short exponent range too (spacing 2^-24 below 6e-5), so A is scaled per voice by a power of
//     two that puts the largest amplitude into [2^13, 2^14) and W by 2^10: a partial 2^-20 of the loudest one still has
//     all 22 bits, nothing overflows (the products add up in fp32), and the write-out multiplies the scale back out.
//   * accumulation: a tensor-core accumulator takes one rounding per MMA; thousands of them in a row drift.  The MMA
//     accumulators are therefore added into fp32 sums in shared memory (FADD, round to nearest) every GM_FLUSH steps
//     (256 partials) and cleared.
//   * tiling: a CTA = 4 warps = 128 blocks x 128 samples of one voice; a warp owns 64 x 64 (128 accumulator registers per
//     thread) and walks through ALL partials of the voice, 8 per step (k16): no partial-range planes, no reduce pass.
//     Rows and columns of a warp tile are anchored at absolute multiples of 64 blocks / 64 samples, partials are taken in
//     record order: a sample's value depends on the bank and its absolute time only, not on how a render is cut up.
// Bound: the legacy tensor path (HMMA) at 1,024 fp16 MACs/clk/SM (tools/microbench/hmma_peak.cu: 550 TFLOP/s on B200),
// shared with the ~3 generating instructions per MMA.  tcgen05 would need the operands in shared memory / TMEM; see DESIGN.md.
#pragma once

#include <cuda_fp16.h>

namespace frb {

constexpr int GM_N = 128;            // samples per block (GEMM N)
constexpr int GM_M = 128;            // blocks per CTA tile (GEMM M)
constexpr int GM_THREADS = 128;      // 4 warps, 2 x 2, warp tile 64 x 64
constexpr int GM_FLUSH = 32;         // k-steps (8 partials each) between flushes of the MMA accumulators
constexpr float GM_WSCALE = 1024.0f; // W is generated times 2^10 (see above)
constexpr int GT_REC_WORDS = 9;      // K1T's record: inc_lo, inc_hi, ph0_hi, kappa, amp, rho^1024 (cos, sin), rho^8 (cos, sin)
constexpr int GM_ROW_STRIDE = 8 * GM_N;   // samples between consecutive rows of a thread (rows g, g + 8, ...)
constexpr size_t GM_SMEM = (size_t)4 * 32 * 32 * sizeof(float4);   // fp32 sums: [warp][quad][lane]

struct OscGemmLaunch {
    const float4* anc; const uint4* ph; const float4* rot;
    const uint32_t* tc;              // K1T: the same record fields word-major per group of 16 (GT_REC_WORDS x 16 words)
    const uint32_t* grp_begin; const uint32_t* n_grp; const float2* vscale;
    const BufferDesc* bufdesc; uint32_t first_buf;
    unsigned long long lo, hi;       // absolute output window [lo, hi)
    unsigned long long tile0;        // absolute index of the first tile (tile = GM_M blocks of GM_N samples)
    int K;                           // records per group of the bank's layout (16)
};

struct GmRecs {                      // 8 consecutive records, staged per warp (double buffered)
    uint4 ph[8]; float4 anc[8]; float4 rot[8];
};

__device__ __forceinline__ void gm_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// (a, b) -> fp16 pairs hi, lo with hi + lo = (a, b) to 22 bits; a in the low half (the lower k index of the fragment)
__device__ __forceinline__ void gm_split(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// amp * 2^(-kappa n) * (sin, cos)(2 pi (inc n + ph0)): osc_group's anchor
__device__ __forceinline__ void gm_anchor(unsigned inc_lo, unsigned inc_hi, unsigned ph0_hi, float kappa, float amp,
                                          unsigned long long n, float& s, float& c) {
    const unsigned n_lo = (unsigned)n, n_hi = (unsigned)(n >> 32);
    const unsigned turns_hi = __umulhi(inc_lo, n_lo) + inc_hi * n_lo + inc_lo * n_hi + ph0_hi;
    const float th = (float)(int)turns_hi * 1.4629180792671596e-9f;               // * 2 pi / 2^32, in [-pi, pi)
    __sincosf(th, &s, &c);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-kappa * (float)n));
    e *= amp;
    s *= e; c *= e;
}
__device__ __forceinline__ void gm_cp16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(GM_THREADS, 2) osc_gemm_kernel(OscGemmLaunch p) {
    extern __shared__ float4 gm_sums[];                     // [4][32][32]
    __shared__ GmRecs gm_recs[4][2];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned g = lane >> 2, tq = lane & 3, wm = w >> 1, wn = w & 1;
    const unsigned v = blockIdx.y;
    const unsigned long long tile = p.tile0 + blockIdx.x;
    const unsigned long long row0 = tile * GM_M + 64u * wm + g;          // this thread's rows: row0 + 8 i, i < 8
    const unsigned col0 = 64u * wn + g;                                  // ... and columns: col0 + 8 i
    float4* mine = gm_sums + (size_t)w * 32 * 32 + lane;
#pragma unroll
    for (int r = 0; r < 32; r++) mine[r * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[4][8][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int u = 0; u < 4; u++) acc[i][j][u] = 0.f;

    const size_t rec0 = (size_t)p.grp_begin[v] * p.K;
    const unsigned n_steps = p.n_grp[v] * (unsigned)p.K / 8u;
    const float2 vs = p.vscale[v];                                       // {scale, 1 / scale}
    const unsigned long long nA = row0 * (unsigned long long)GM_N;       // sample index of this thread's first row

    auto stage = [&](unsigned step, unsigned buf) {                      // records of `step` -> this warp's buffer
        if (lane < 24) {
            const unsigned which = lane >> 3, k = lane & 7;
            const size_t r = rec0 + 8u * step + k;
            GmRecs& d = gm_recs[w][buf];
            if (which == 0) gm_cp16(&d.ph[k], p.ph + r);
            else if (which == 1) gm_cp16(&d.anc[k], p.anc + r);
            else gm_cp16(&d.rot[k], p.rot + r);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float4 s = mine[(i * 8 + j) * 32];
                s.x += acc[i][j][0]; s.y += acc[i][j][1]; s.z += acc[i][j][2]; s.w += acc[i][j][3];
                mine[(i * 8 + j) * 32] = s;
                acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
            }
    };

    if (n_steps) stage(0, 0);
    for (unsigned step = 0; step < n_steps; step++) {
        const unsigned buf = step & 1u;
        if (step + 1 < n_steps) stage(step + 1, buf ^ 1u); else asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        const GmRecs& R = gm_recs[w][buf];
        // ---- W fragments of the warp's 8 column tiles: B0 = partial tq, B1 = partial tq + 4; (cos, sin) = (k even, k odd)
        uint32_t bh[8][2], bl[8][2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const uint4 ph = R.ph[tq + 4 * q];
            const float kappa = R.anc[tq + 4 * q].z;
            const float4 rt = R.rot[tq + 4 * q];
            float ws, wc;
            gm_anchor(ph.x, ph.y, 0u, kappa, GM_WSCALE, col0, ws, wc);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                gm_split(wc, ws, bh[j][q], bl[j][q]);
                const float nc = fmaf(wc, rt.z, -ws * rt.w), ns = fmaf(ws, rt.z, wc * rt.w);
                wc = nc; ws = ns;
            }
        }
        // ---- A: anchors at the thread's first row, then 8 rows by rotation; (sin, cos) = (k even, k odd)
        float zs[2], zc[2], cr[2], ci[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const uint4 ph = R.ph[tq + 4 * q];
            const float4 an = R.anc[tq + 4 * q];
            const float4 rt = R.rot[tq + 4 * q];
            gm_anchor(ph.x, ph.y, ph.w, an.z, an.y * vs.x, nA, zs[q], zc[q]);
            cr[q] = rt.x; ci[q] = rt.y;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int h = 0; h < 2; h++) {                            // rows g + 16 i and g + 16 i + 8
                    gm_split(zs[q], zc[q], ah[2 * q + h], al[2 * q + h]);
                    const float ns = fmaf(zs[q], cr[q], zc[q] * ci[q]), nc = fmaf(zc[q], cr[q], -zs[q] * ci[q]);
                    zs[q] = ns; zc[q] = nc;
                }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                gm_mma(acc[i][j], al, bh[j]);
                gm_mma(acc[i][j], ah, bl[j]);
                gm_mma(acc[i][j], ah, bh[j]);
            }
        }
        if ((step % GM_FLUSH) == GM_FLUSH - 1) flush();
        __syncwarp();                                                    // everyone is done with R before it is restaged
    }
    flush();
    // ---- write-out: c0 c1 = (row g, cols 2 tq, 2 tq + 1), c2 c3 = (row g + 8, same columns) of every 16 x 8 tile
    const BufferDesc bd = p.bufdesc[p.first_buf + v];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 s = mine[(i * 8 + j) * 32];
            const unsigned long long r = tile * GM_M + 64u * wm + 16u * i + g;
            const unsigned c = 64u * wn + 8u * j + 2u * tq;
            const unsigned long long t0 = r * GM_N + c, t1 = t0 + (unsigned long long)GM_ROW_STRIDE;
            const float un = vs.y * (1.0f / GM_WSCALE);
            const float o[4] = {s.x * un, s.y * un, s.z * un, s.w * un};
            if (t0 >= p.lo && t0 + 2 <= p.hi) *reinterpret_cast<float2*>(bd.data + (t0 & bd.mask)) = make_float2(o[0], o[1]);
            else { if (t0 >= p.lo && t0 < p.hi) bd.data[t0 & bd.mask] = o[0]; if (t0 + 1 >= p.lo && t0 + 1 < p.hi) bd.data[(t0 + 1) & bd.mask] = o[1]; }
            if (t1 >= p.lo && t1 + 2 <= p.hi) *reinterpret_cast<float2*>(bd.data + (t1 & bd.mask)) = make_float2(o[2], o[3]);
            else { if (t1 >= p.lo && t1 < p.hi) bd.data[t1 & bd.mask] = o[2]; if (t1 + 1 >= p.lo && t1 + 1 < p.hi) bd.data[(t1 + 1) & bd.mask] = o[3]; }
        }
}

}  // namespace frb
