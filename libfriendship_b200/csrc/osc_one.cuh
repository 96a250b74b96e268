// osc_one.cuh — the oscillator of a ONE-partial voice (one exciter per voice: BASELINE configs[2]), shared by the bank
// kernel for such voices (osc_one_kernel, osc.cu) and by the fused exciter -> biquad -> comb kernel (scan.cu), so that
// both produce the same bits.
//
// Same resonator as the big-bank kernel (osc.cu: lifting form, 3 FFMA per sample, exact fixed-point phase at the anchor),
// but re-anchored every 8 samples at absolute multiples of 8: with one partial per voice the anchor's three MUFU
// operations per 8 samples are noise beside the voice's ring traffic, a group of 8 is exactly what one thread of the
// chain kernel owns, and any 8-aligned window can be produced without running a recurrence up to it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace frb {

// Device arrays of a bank whose voices have at most one partial each (record layout K = 1: one record per group).
struct OscOneSrc {
    const float4* hot;            // {a, b, cm1, k1}
    const float4* anc;            // {k2, amp, kappa, invA}
    const uint4* ph;              // {inc_lo, inc_hi, ph0_lo, ph0_hi}: 64-bit fixed-point turns per sample / at t = 0
    const uint32_t* grp_begin;    // per voice: its record
    const uint32_t* n_grp0;       // per voice: 1 if the partial is of class 0 (cos w >= 0)
    const uint32_t* n_grp;        // per voice: 0 (silent) or 1
    float max_attack;             // longest attack ramp of the bank, samples
};

struct OscOneVoice {              // one voice's record, as loaded
    float4 h, an;
    uint4 ph;
    unsigned flags;               // bit 0: the voice has a partial; bit 1: class 1 (runs in the alternating-sign domain)
};

__device__ __forceinline__ OscOneVoice osc_one_load(const OscOneSrc& s, unsigned v) {
    OscOneVoice o;
    o.h = make_float4(0.f, 0.f, 0.f, 0.f); o.an = o.h; o.ph = make_uint4(0u, 0u, 0u, 0u); o.flags = 0u;
    if (s.n_grp[v]) {
        const size_t r = s.grp_begin[v];
        o.h = __ldg(s.hot + r); o.an = __ldg(s.anc + r); o.ph = __ldg(s.ph + r);
        o.flags = 1u | (s.n_grp0[v] == 0 ? 2u : 0u);
    }
    return o;
}

// r[u] = out(n8 + u), u < 8, n8 a multiple of 8.  The anchor is osc_group's (osc.cu), operation for operation.
__device__ __forceinline__ void osc_one_group8(const float4 h, const float4 an, const uint4 ph, unsigned flags,
                                               unsigned long long n8, float max_attack, float (&r)[8]) {
    const float nf = (float)n8;
    const unsigned n_lo = (unsigned)n8, n_hi = (unsigned)(n8 >> 32);
    // top 32 bits of (inc * n + ph0) mod 2^64: the integer wrap-around is the exact range reduction
    const unsigned turns_hi = __umulhi(ph.x, n_lo) + ph.y * n_lo + ph.x * n_hi + ph.w;
    const float th = (float)(int)turns_hi * 1.4629180792671596e-9f;             // * 2 pi / 2^32, in [-pi, pi)
    float sn, cs;
    __sincosf(th, &sn, &cs);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(-an.z, nf)));     // amp * exp(-n / tau) = amp * 2^(-kappa n)
    e = __fmul_rn(e, an.y);
    float y = __fmul_rn(e, sn);
    float x = __fmul_rn(e, fmaf(h.w, sn, __fmul_rn(an.x, cs)));
#pragma unroll
    for (int u = 0; u < 8; u++) {
        r[u] = y;
        x = fmaf(-h.x, y, x);
        const float t = fmaf(h.z, y, y);
        y = fmaf(h.y, x, t);
    }
    if (nf < max_attack) {                                      // inside somebody's attack ramp: min(t / A, 1)
#pragma unroll
        for (int u = 0; u < 8; u++) r[u] = __fmul_rn(fminf(__fmul_rn(nf + (float)u, an.w), 1.0f), r[u]);
    }
    if (flags & 2u) { r[1] = -r[1]; r[3] = -r[3]; r[5] = -r[5]; r[7] = -r[7]; }   // class 1: (-1)^n y[n], n8 is even
    if (!(flags & 1u)) {
#pragma unroll
        for (int u = 0; u < 8; u++) r[u] = 0.0f;
    }
}

}  // namespace frb
