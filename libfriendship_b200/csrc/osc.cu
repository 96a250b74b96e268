// osc.cu — K1: oscillator bank for sm_100a (extension node FRB_KIND_OSCBANK; not in the reference).
//
//   out_v(t) = sum_p amp_p * min(t/A_p, 1) * exp(-t/tau_p) * sin(2 pi f_p t / sr + phi_p)
//
// This file: the resonator kernel K1 (attack ramps of every bank; banks of few or small voices), the definition pipeline and
// the launch policy.  Past the ramps a bank of big voices is a matrix product and runs on the tensor cores: osc_tc.cuh (K1T,
// tcgen05) / osc_gemm.cuh (K1G, mma.sync), see osc_gemm_wanted and launch_osc below.
// K1's bound: the FP32 FMA pipe (128 lanes/clk/SM; measured 125 on B200 with tools/microbench/fma_peak.cu).
//
// Design (measured choices: tools/microbench/, profiles/ncu_osc_r1*_summary.txt):
//  * One thread owns one time SEGMENT of L = 128 consecutive samples of one voice, and walks through the partials of
//    that voice in groups of K = 16.  Lanes of a warp therefore differ in time, not in partial: the per-partial
//    coefficients are warp-uniform, there is no cross-lane reduction at all, and the K recurrences of a group give
//    the scheduler K independent FMA chains per thread.  One warp per CTA.
//  * Each partial is a damped two-state resonator in lifting form, 3 FFMA + 1 FADD per partial-sample:
//        x <- x - a*y ;  y <- (1+cm1)*y + b*x ;  acc += y
//    with a*b = (1-rho)^2 + 4 rho sin^2(w'/2), cm1 = rho^2 - 1: the eigenvalues are rho*exp(+-i w'), so y is exactly
//    amp * rho^n * sin(w' n + phi) in exact arithmetic; the exponential decay costs nothing extra.
//    For cos w < 0 the resonator runs at w' = pi - w (keeps a*b small: Nyquist is as well conditioned as DC) on
//    the variables (-1)^n x, (-1)^n y, which obey the very same update; the kernel flips the sign of the odd samples
//    of the accumulator between the two classes instead of having a second loop.
//  * The coefficients of a group are staged in shared memory and read at constant addresses, so that ptxas keeps
//    them in UNIFORM registers: every hot FFMA then reads two vector registers, not three.  Register-file read
//    bandwidth, not the FMA pipe, was the limit before (27% dispatch stalls).  ptxas only does this with >= 16 uses
//    per loop iteration and ONE instance of the hot loop per kernel — hence the 16-sample unroll, the single
//    recurrence for both classes, and the separate (slower) kernel for the few segments inside the attack ramp.
//  * State is re-anchored exactly at the start of every segment (segments are aligned to absolute time, so the
//    result does not depend on how a render is cut into blocks): phase = inc*n + phase0 in 64-bit fixed-point
//    turns (exact range reduction by integer wrap-around), sin/cos and 2^x by MUFU.
//    With L = 128 the worst single-partial error is ~6e-6 of its amplitude (tools/osc_math_check.py).
//  * Coefficients come from an fp64 setup kernel at definition time.
//  * Per-thread accumulators for the L samples live in shared memory as float4 columns (conflict-free),
//    read-modify-written once per 16 samples per group.
//  * Voices with many partials are split into partial ranges over CTAs (fixed per bank: the summation order never
//    depends on the block size); the range planes are summed in fixed order by osc_reduce_kernel.
#include "osc.cuh"
#include "osc_one.cuh"
#include "osc_gemm.cuh"
#include "osc_tc.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace frb {

constexpr int OSC_K = 16;           // partials per group (independent FMA chains per thread) for ordinary banks
constexpr int OSC_K_SMALL = 4;      // banks whose voices have <= 8 partials: less padding
constexpr int OSC_K_ONE = 1;        // banks of one-partial voices (one exciter per voice, BASELINE configs[2]): no padding
constexpr int OSC_THREADS = 32;     // threads (= time segments) per CTA: one warp, so the per-group barrier couples no warps
constexpr int OSC_LMAX = 256;       // max segment length (shared memory: L * THREADS * 4 B)
constexpr uint32_t kOscMinGroupsPerCta = 8;   // a partial-range split leaves every CTA at least this many groups of K partials

struct OscBankDev {
    uint32_t n_voices = 0;
    uint64_t n_partials = 0;        // as defined by the user
    uint64_t n_records = 0;         // after per-(voice, class) padding to multiples of K
    int K = OSC_K;                  // group size of this bank's record layout
    double sample_rate = 48000.0;
    float max_attack = 0.0f;
    uint32_t max_groups = 0;        // largest number of groups in one voice
    uint32_t split = 1;             // partial-range split per voice (fixed per bank: keeps the summation order fixed)
    float4* d_hot = nullptr;        // {a, b, cm1, k1}
    float4* d_anc = nullptr;        // {k2, amp, kappa, invA}
    uint4* d_ph = nullptr;          // {inc_lo, inc_hi, ph0_lo, ph0_hi}
    float4* d_rot = nullptr;        // K1G (osc_gemm.cuh): {rho^1024 (cos, sin)(1024 w), rho^8 (cos, sin)(8 w)}
    uint32_t* d_tc = nullptr;       // K1T (osc_tc.cuh): per record group 9 words x 16 records (GT_REC_WORDS), what a stage reads
    uint64_t rec_cap = 0;           // records the four arrays above can hold
    float2* d_vscale = nullptr;     // K1G: per voice {2^k, 2^-k} with max |amp| 2^k in [2^13, 2^14)
    bool gemm = false;              // K1G renders this bank past its attack ramps (osc_gemm_wanted)
    uint32_t* d_grp_begin = nullptr;   // per voice: first group
    uint32_t* d_n_grp0 = nullptr;      // per voice: groups of class 0
    uint32_t* d_n_grp = nullptr;       // per voice: groups in total
    uint32_t voice_cap = 0;
    char* d_raw = nullptr;          // definition-time scratch (raw parameters, ranks); kept only by a re-defined bank
    size_t raw_cap = 0;
    mutable float* d_planes = nullptr; // scratch for split > 1: TWO buffers [split][n_voices][plane_len] (see launch_osc)
    mutable uint64_t planes_cap = 0;   // floats per buffer
    // The plane reduce of sub-block k runs on its own stream beside the main kernel of sub-block k + 1 (launch_osc):
    mutable cudaStream_t red = nullptr;
    mutable cudaStream_t main2 = nullptr;          // odd sub-blocks' main kernels: their head fills the even ones' tail
    mutable cudaEvent_t ev_enter = nullptr;
    mutable cudaEvent_t ev_main[2] = {nullptr, nullptr}, ev_red[2] = {nullptr, nullptr};
    mutable bool red_pending[2] = {false, false};
    mutable cudaStream_t side = nullptr;   // the (tiny, slow) attack-ramp kernel runs beside the main kernel
    mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    mutable cudaStream_t gside = nullptr;  // K1T / K1G: the resonator kernels of the ramp region run beside the matrix-product kernel
    mutable cudaEvent_t ev_gfork = nullptr, ev_gjoin = nullptr;
    ~OscBankDev() {
        if (side) cudaStreamDestroy(side);
        if (gside) cudaStreamDestroy(gside);
        if (ev_gfork) cudaEventDestroy(ev_gfork);
        if (ev_gjoin) cudaEventDestroy(ev_gjoin);
        if (red) cudaStreamDestroy(red);
        if (main2) cudaStreamDestroy(main2);
        if (ev_enter) cudaEventDestroy(ev_enter);
        for (int i = 0; i < 2; i++) { if (ev_main[i]) cudaEventDestroy(ev_main[i]); if (ev_red[i]) cudaEventDestroy(ev_red[i]); }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        cudaFree(d_hot); cudaFree(d_anc); cudaFree(d_ph); cudaFree(d_rot); cudaFree(d_tc); cudaFree(d_vscale);
        cudaFree(d_grp_begin); cudaFree(d_n_grp0); cudaFree(d_n_grp); cudaFree(d_planes); cudaFree(d_raw);
    }
    // a bank that replaces `src` under the same key takes over every device allocation of it (cudaMalloc / cudaFree cost
    // milliseconds each beside a busy context; a re-definition per render must not pay them)
    void take_buffers(OscBankDev& src) {
        auto mv = [](auto& a, auto& b2) { a = b2; b2 = {}; };
        mv(d_hot, src.d_hot); mv(d_anc, src.d_anc); mv(d_ph, src.d_ph); mv(d_rot, src.d_rot); mv(d_tc, src.d_tc); mv(rec_cap, src.rec_cap);
        mv(d_vscale, src.d_vscale);
        mv(d_grp_begin, src.d_grp_begin); mv(d_n_grp0, src.d_n_grp0); mv(d_n_grp, src.d_n_grp); mv(voice_cap, src.voice_cap);
        mv(d_raw, src.d_raw); mv(raw_cap, src.raw_cap);
        mv(d_planes, src.d_planes); mv(planes_cap, src.planes_cap);
        mv(side, src.side); mv(ev_fork, src.ev_fork); mv(ev_join, src.ev_join);
        mv(gside, src.gside); mv(ev_gfork, src.ev_gfork); mv(ev_gjoin, src.ev_gjoin);
        mv(red, src.red); mv(main2, src.main2); mv(ev_enter, src.ev_enter);
        for (int i = 0; i < 2; i++) { mv(ev_main[i], src.ev_main[i]); mv(ev_red[i], src.ev_red[i]); red_pending[i] = false; }
    }
};

OscBankInfo osc_info(const OscBankDev& b) { return OscBankInfo{b.n_voices, b.n_partials}; }
bool osc_usable(const OscBankDev& b) { return b.rec_cap >= std::max<uint64_t>(b.n_records, 1) && b.voice_cap >= std::max<uint32_t>(b.n_voices, 1); }

// Definition-time pipeline, all on the device (host work is O(n_voices)): a bank definition is the per-render "input" of a
// synthesis graph, so its upload is on the end-to-end path (bench.py e2e re-defines the bank every step).
//   1. osc_rank_kernel   — resonator class of every partial, its stable rank inside (voice, class), per-voice class-0 count,
//                          longest attack ramp
//   2. host              — group tables from the n_voices counts (prefix over voices, padding to multiples of K)
//   3. osc_fill_kernel   — every record := the silent padding record
//   4. osc_setup_kernel  — raw parameters -> resonator records (fp64), scattered to their (voice, class, rank) position

__device__ __forceinline__ double osc_turns(double freq, double sample_rate) {
    double fr = freq / sample_rate;
    return fr - floor(fr);                                  // turns per sample in [0, 1)
}
__device__ __forceinline__ bool osc_class1(double fr) { return fr > 0.25 && fr < 0.75; }   // cos w < 0

constexpr unsigned OSC_RANK_CLASS1 = 0x80000000u;
constexpr int OSC_RANK_THREADS = 1024;

// One CTA walks one voice in tiles of 1024 partials (ballot + warp totals: a block-wide stable partition).
__global__ void __launch_bounds__(OSC_RANK_THREADS)
osc_rank_kernel(unsigned n_voices, const unsigned long long* __restrict__ voice_offsets, double sample_rate,
                const double* __restrict__ freq, const float* __restrict__ amp, const float* __restrict__ attack,
                unsigned* __restrict__ rank, unsigned* __restrict__ cnt0, unsigned* __restrict__ max_attack_bits) {
    __shared__ unsigned warp_tot[OSC_RANK_THREADS / 32];
    const unsigned tid = threadIdx.x, w = tid >> 5, ln = tid & 31;
    float my_att = 0.f;
    for (unsigned v = blockIdx.x; v < n_voices; v += gridDim.x) {
        const unsigned long long lo = voice_offsets[v], hi = voice_offsets[v + 1];
        unsigned running0 = 0;                              // class-0 partials of this voice before the tile
        for (unsigned long long tile = lo; tile < hi; tile += OSC_RANK_THREADS) {
            const unsigned long long p = tile + tid;
            const bool valid = p < hi;
            const bool c1 = valid && osc_class1(osc_turns(freq[valid ? p : lo], sample_rate));
            const bool c0 = valid && !c1;
            const unsigned bal = __ballot_sync(0xffffffffu, c0);
            if (ln == 0) warp_tot[w] = __popc(bal);
            __syncthreads();
            unsigned before = 0, total = 0;
            for (unsigned k = 0; k < OSC_RANK_THREADS / 32; k++) {
                const unsigned t = warp_tot[k];
                before += (k < w) ? t : 0u;
                total += t;
            }
            __syncthreads();
            const unsigned n0_before = running0 + before + __popc(bal & ((1u << ln) - 1u));
            if (valid) {
                rank[p] = c1 ? (OSC_RANK_CLASS1 | (unsigned)((p - lo) - n0_before)) : n0_before;
                const float at = attack[p];
                if (at > 0.f && amp[p] != 0.f) my_att = fmaxf(my_att, at);
            }
            running0 += total;
        }
        if (tid == 0) cnt0[v] = running0;
    }
    for (int o = 16; o; o >>= 1) my_att = fmaxf(my_att, __shfl_xor_sync(0xffffffffu, my_att, o));
    if (ln == 0 && my_att > 0.f) atomicMax(max_attack_bits, __float_as_uint(my_att));   // positive floats order like their bits
}

__global__ void osc_fill_kernel(uint64_t n, float4* __restrict__ hot, float4* __restrict__ anc, uint4* __restrict__ ph,
                                float4* __restrict__ rot, uint32_t* __restrict__ tc) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        hot[i] = make_float4(0.f, 0.f, 0.f, 0.f);           // a resonator that stays at 0
        anc[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
        ph[i] = make_uint4(0u, 0u, 0u, 0u);
        if (rot) rot[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tc) {
#pragma unroll
            for (int wd = 0; wd < GT_REC_WORDS; wd++) tc[(i / 16) * (GT_REC_WORDS * 16) + wd * 16 + (i % 16)] = 0u;
        }
    }
}

// K1G / K1T: per voice the power of two that brings the largest amplitude into [2^13, 2^14) (fp16 operands, osc_gemm.cuh)
__global__ void osc_vscale_kernel(unsigned n_voices, const unsigned long long* __restrict__ voice_offsets,
                                  const float* __restrict__ amp, float2* __restrict__ vscale) {
    __shared__ float warp_max[8];
    for (unsigned v = blockIdx.x; v < n_voices; v += gridDim.x) {
        float m = 0.f;
        for (unsigned long long p = voice_offsets[v] + threadIdx.x; p < voice_offsets[v + 1]; p += blockDim.x) {
            const float a = fabsf(amp[p]);
            if (isfinite(a)) m = fmaxf(m, a);
        }
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < 8; k++) m = fmaxf(m, warp_max[k]);
            int e = 0;
            if (m > 0.f) frexpf(m, &e);                      // m = f 2^e, f in [0.5, 1)
            e = max(-100, min(100, e));
            vscale[v] = make_float2(exp2f((float)(14 - e)), exp2f((float)(e - 14)));
        }
    }
}

// blockIdx.y strides the voices, blockIdx.x the partials of a voice.
__global__ void osc_setup_kernel(unsigned n_voices, const unsigned long long* __restrict__ voice_offsets, int K,
                                 const uint32_t* __restrict__ grp_begin, const uint32_t* __restrict__ n_grp0,
                                 const unsigned* __restrict__ rank, double sample_rate, const double* __restrict__ freq,
                                 const float* __restrict__ amp, const float* __restrict__ phase,
                                 const float* __restrict__ attack, const float* __restrict__ tau,
                                 float4* __restrict__ hot, float4* __restrict__ anc, uint4* __restrict__ ph,
                                 float4* __restrict__ rot, uint32_t* __restrict__ tc) {
    const double PI = 3.14159265358979323846;
    for (unsigned v = blockIdx.y; v < n_voices; v += gridDim.y) {
    const unsigned long long vlo = voice_offsets[v], vhi = voice_offsets[v + 1];
    const uint64_t base0 = (uint64_t)grp_begin[v] * K, base1 = ((uint64_t)grp_begin[v] + n_grp0[v]) * K;
    for (unsigned long long p = vlo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < vhi;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        const float am = amp[p];
        if (am == 0.0f) continue;                           // silent partial: its record keeps the padding pattern
        const unsigned rk = rank[p];
        const uint64_t i = (rk & OSC_RANK_CLASS1) ? base1 + (rk & ~OSC_RANK_CLASS1) : base0 + rk;
        const double fr = osc_turns(freq[p], sample_rate);
        const bool cls1 = osc_class1(fr);
        double wp;                                          // w' in [-pi/2, pi/2]
        if (cls1) wp = PI - 2.0 * PI * fr;
        else wp = (fr <= 0.25) ? 2.0 * PI * fr : 2.0 * PI * (fr - 1.0);
        const float tf = tau[p];
        const bool decays = tf > 0.0f && isfinite(tf);
        const double one_m_rho = decays ? -expm1(-1.0 / (double)tf) : 0.0;
        const double rho = 1.0 - one_m_rho;
        const double sh = sin(0.5 * wp);
        const double ab = one_m_rho * one_m_rho + 4.0 * rho * sh * sh;
        float a32 = 0.f, b32 = 0.f, k1 = 0.f, k2 = 0.f;
        if (ab > 0.0) {
            // a = b = sqrt(ab) up to rounding; try neighbours of a so that fl(a)*fl(b) is as close to ab as possible
            // (the product sets the realised frequency: its rounding error is the systematic phase drift).
            const float a0 = (float)sqrt(ab);
            double best = 1e300;
            for (int d = -3; d <= 3; d++) {
                float ac = __int_as_float(__float_as_int(a0) + d);
                if (!(ac > 0.f)) continue;
                float bc = (float)(ab / (double)ac);
                double err = fabs((double)ac * (double)bc - ab);
                if (err < best) { best = err; a32 = ac; b32 = bc; }
            }
            const double bd = (double)b32;
            k1 = (float)((one_m_rho + 2.0 * rho * sh * sh) / bd);          // (1 - rho cos w') / b
            k2 = (float)((cls1 ? -1.0 : 1.0) * rho * sin(wp) / bd);
        }
        const float cm1 = (float)(-(one_m_rho) * (1.0 + rho));             // rho^2 - 1
        hot[i] = make_float4(a32, b32, cm1, k1);
        const float kappa = decays ? (float)(1.4426950408889634 / (double)tf) : 0.0f;
        const float at = attack[p];
        const float invA = (at > 0.0f) ? 1.0f / at : __int_as_float(0x7f800000);
        anc[i] = make_float4(k2, am, kappa, invA);
        const unsigned long long inc = __double2ull_rn(fr * 18446744073709551616.0);
        double p0 = (double)phase[p] / (2.0 * PI);
        p0 -= floor(p0);
        const unsigned long long ph0 = __double2ull_rn(p0 * 18446744073709551616.0);
        ph[i] = make_uint4((unsigned)(inc & 0xffffffffull), (unsigned)(inc >> 32), (unsigned)(ph0 & 0xffffffffull), (unsigned)(ph0 >> 32));
        if (rot) {                                          // K1G: N samples on = a rotation by N w and a decay by rho^N
            float r4[4];
            const double steps[2] = {8.0 * 128.0, 8.0};     // rows of a thread are 8 blocks apart, its columns 8 samples
            for (int k = 0; k < 2; k++) {
                double fn = fr * steps[k];
                fn -= floor(fn);
                const double R = decays ? exp(-steps[k] / (double)tf) : 1.0;
                double sn, cs;
                sincos(2.0 * PI * fn, &sn, &cs);
                r4[2 * k] = (float)(R * cs); r4[2 * k + 1] = (float)(R * sn);
            }
            rot[i] = make_float4(r4[0], r4[1], r4[2], r4[3]);
            if (tc) {                                       // the same numbers, word-major per group of 16 records
                uint32_t* g = tc + (i / 16) * (GT_REC_WORDS * 16) + (i % 16);
                g[0 * 16] = (unsigned)(inc & 0xffffffffull); g[1 * 16] = (unsigned)(inc >> 32); g[2 * 16] = (unsigned)(ph0 >> 32);
                g[3 * 16] = __float_as_uint(kappa); g[4 * 16] = __float_as_uint(am);
                g[5 * 16] = __float_as_uint(r4[0]); g[6 * 16] = __float_as_uint(r4[1]);
                g[7 * 16] = __float_as_uint(r4[2]); g[8 * 16] = __float_as_uint(r4[3]);
            }
        }
    }
    }
}

// Which banks K1G renders (past their attack ramps): a property of the bank alone, so that a sample's value never depends on
// how a render is cut into calls.  The tensor-core kernel wants a few hundred partials per voice (a step is 8 partials,
// a tile 16,384 samples of one voice) and enough voices to fill the machine with tiles; everything else stays on the
// resonator kernel.  FRB_OSC_GEMM (measurement / test knob): unset or 1 = this rule, on tcgen05 (K1T); 0 = never;
// 2 / 3 = every bank with the 16-record layout, on mma.sync (K1G) / tcgen05; 5 = this rule on mma.sync.
static int osc_gemm_mode() {
    static const int mode = [] { const char* e = getenv("FRB_OSC_GEMM"); return e ? atoi(e) : 1; }();
    return mode;
}
static bool osc_gemm_wanted(int K, uint32_t n_voices, uint64_t n_records) {
    const int mode = osc_gemm_mode();
    if (mode == 0 || K != OSC_K || n_voices == 0) return false;
    if (mode == 2 || mode == 3) return true;
    return n_voices >= 4 && n_records / n_voices >= 512;
}
// which of the two matrix-product kernels: tcgen05 (K1T, osc_tc.cuh) unless the knob asks for mma.sync (K1G, osc_gemm.cuh)
static bool osc_gemm_tcgen05() { const int m = osc_gemm_mode(); return m != 2 && m != 5; }

std::shared_ptr<OscBankDev> osc_create(const frb_oscbank_desc* d, cudaStream_t stream, std::string* err,
                                       const std::shared_ptr<OscBankDev>& recycle, uint32_t shard_rank, uint32_t shard_world,
                                       bool allow_tensor) {
    std::shared_ptr<OscBankDev> b;
    bool stolen = false;
    auto fail = [&](const std::string& m) {
        if (err) *err = m;
        if (stolen) recycle->take_buffers(*b);              // hand the allocations back (osc_usable tells if they survived)
        return std::shared_ptr<OscBankDev>();
    };
    if (!d->voice_offsets || (d->n_partials && (!d->freq_hz || !d->amp || !d->phase || !d->attack || !d->tau)))
        return fail("oscbank: null array");
    if (!(d->sample_rate > 0.0)) return fail("oscbank: sample_rate must be positive");
    if (d->voice_offsets[0] != 0 || d->voice_offsets[d->n_voices] != d->n_partials) return fail("oscbank: voice_offsets must span [0, n_partials]");
    // A renderer that shards voices over several devices (multi.cu) defines on every device the COMPACT bank of the
    // voices that device owns (global voice v = shard_rank + lane * shard_world becomes lane `lane`; the flattener maps
    // the graph's lane numbers the same way): only their parameters travel — straight from the caller's arrays, one
    // copy per run of owned voices — and nothing downstream knows about the other ranks' voices.
    if (shard_world == 0) shard_world = 1;
    for (uint32_t v = 0; v < d->n_voices; v++)
        if (d->voice_offsets[v + 1] < d->voice_offsets[v]) return fail("oscbank: voice_offsets must be non-decreasing");
    const uint32_t nv = d->n_voices > shard_rank ? (d->n_voices - shard_rank + shard_world - 1) / shard_world : 0;
    std::vector<uint64_t> vo((size_t)nv + 1, 0);
    struct Run { uint64_t src, dst, n; };
    std::vector<Run> runs;
    uint64_t mx = 0;
    for (uint32_t l = 0; l < nv; l++) {
        const uint32_t v = shard_rank + l * shard_world;
        const uint64_t len = d->voice_offsets[v + 1] - d->voice_offsets[v];
        vo[l + 1] = vo[l] + len;
        if (len) {
            if (!runs.empty() && runs.back().src + runs.back().n == d->voice_offsets[v]) runs.back().n += len;
            else runs.push_back(Run{d->voice_offsets[v], vo[l], len});
        }
        mx = std::max<uint64_t>(mx, len);
    }
    if (mx >= (1ull << 31)) return fail("oscbank: more than 2^31 partials in one voice");
    b = std::make_shared<OscBankDev>();
    b->n_voices = nv;
    b->n_partials = vo[nv];
    b->sample_rate = d->sample_rate;
    b->K = (mx <= 1) ? OSC_K_ONE : (mx <= 8) ? OSC_K_SMALL : OSC_K;
    const int K = b->K;
    const uint64_t np = vo[nv];
    auto upload = [&](void* dst, const void* src, size_t esz) -> cudaError_t {
        for (const Run& r : runs) {
            cudaError_t e = cudaMemcpyAsync((char*)dst + r.dst * esz, (const char*)src + r.src * esz, r.n * esz, cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };

#define OC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string("oscbank: ") + cudaGetErrorString(e_)); } while (0)
    const bool trace = getenv("FRB_TRACE") != nullptr;      // tuning aid: host-clock phases of a definition on stderr
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(stream);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[frb] define_oscbank %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    // raw parameters + scratch in one allocation: [freq f64][amp][phase][attack][tau][rank u32][offsets u64][cnt0 u32][max_attack u32]
    if (recycle) {                                          // nothing of the old bank may still be in flight
        OC(cudaStreamSynchronize(stream));
        b->take_buffers(*recycle);
        stolen = true;
    }
    const uint64_t npp = (np + 3) & ~3ull;                  // keeps every sub-array 16-byte aligned
    const uint64_t nvp = ((uint64_t)nv + 4) & ~3ull;
    const size_t raw_bytes = npp * (sizeof(double) + 5 * 4) + nvp * (sizeof(uint64_t) + 4) + 16;
    if (b->raw_cap < raw_bytes) {
        cudaFree(b->d_raw); b->d_raw = nullptr; b->raw_cap = 0;
        OC(cudaMalloc(&b->d_raw, raw_bytes));
        b->raw_cap = raw_bytes;
    }
    char* const d_raw = b->d_raw;
    double* d_freq = reinterpret_cast<double*>(d_raw);
    float* d_amp = reinterpret_cast<float*>(d_freq + npp);
    float* d_phase = d_amp + npp; float* d_attack = d_phase + npp; float* d_tau = d_attack + npp;
    unsigned* d_rank = reinterpret_cast<unsigned*>(d_tau + npp);
    unsigned long long* d_offs = reinterpret_cast<unsigned long long*>(d_rank + npp);
    unsigned* d_cnt0 = reinterpret_cast<unsigned*>(d_offs + nvp);
    unsigned* d_maxatt = d_cnt0 + nvp;
    lap("alloc raw");
    OC(cudaMemcpyAsync(d_offs, vo.data(), ((size_t)nv + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    OC(cudaMemsetAsync(d_maxatt, 0, 4, stream));
    if (np) {
        OC(upload(d_freq, d->freq_hz, sizeof(double)));
        OC(upload(d_amp, d->amp, sizeof(float)));
        OC(upload(d_attack, d->attack, sizeof(float)));
    }
    lap("h2d freq/amp/att");
    std::vector<uint32_t> cnt0(nv, 0);
    uint32_t max_attack_bits = 0;
    if (nv) {
        osc_rank_kernel<<<std::min<uint32_t>(nv, 148 * 2), OSC_RANK_THREADS, 0, stream>>>(nv, d_offs, d->sample_rate, d_freq, d_amp, d_attack, d_rank, d_cnt0, d_maxatt);
        OC(cudaGetLastError());
        OC(cudaMemcpyAsync(cnt0.data(), d_cnt0, (size_t)nv * 4, cudaMemcpyDeviceToHost, stream));
    }
    OC(cudaMemcpyAsync(&max_attack_bits, d_maxatt, 4, cudaMemcpyDeviceToHost, stream));
    if (np) {   // the remaining parameters travel while the host waits for the counts
        OC(upload(d_phase, d->phase, sizeof(float)));
        OC(upload(d_tau, d->tau, sizeof(float)));
    }
    OC(cudaStreamSynchronize(stream));
    lap("rank + h2d rest");
    memcpy(&b->max_attack, &max_attack_bits, 4);

    // group by (voice, class), every class segment padded to a multiple of K with silent records
    std::vector<uint32_t> grp_begin(nv), n_grp0(nv), n_grp(nv);
    uint64_t groups = 0;
    for (uint32_t v = 0; v < nv; v++) {
        const uint64_t len = vo[v + 1] - vo[v];
        const uint64_t g0 = (cnt0[v] + K - 1) / K, g1 = (len - cnt0[v] + K - 1) / K;
        if (groups + g0 + g1 >= (1ull << 32)) return fail("oscbank: too many partial groups");
        grp_begin[v] = (uint32_t)groups;
        n_grp0[v] = (uint32_t)g0;
        n_grp[v] = (uint32_t)(g0 + g1);
        groups += g0 + g1;
        b->max_groups = std::max(b->max_groups, n_grp[v]);
    }
    const uint64_t n = groups * K;
    b->n_records = n;
    // Partial-range split: one CTA is one warp working through (groups / split) groups for 32 segments (4,096 samples).
    // All CTAs of a launch cost the same, so the only imbalance is the partially filled last wave (148 SMs x ~12
    // resident CTAs).  CTAs per launch = (samples / 4096) x voices x split and the split planes cost
    // split x voices x samples x 8 B of HBM traffic, so the split is kept as SMALL as wave balance allows —
    // voices x split >= 1024..2048 (measured sweep, tools/split_sweep.sh: 64 voices peak at 2048, 8 voices at 1024;
    // 4096 costs 4% at 8 voices in plane traffic, 256 costs 3% in tail) — and long blocks do the rest.
    // Fixed per bank, so the order of summation (and hence the result) never depends on the block size of a render.
    {
        const uint32_t live_voices = std::max<uint32_t>(nv, 1);
        uint64_t target = live_voices >= 32 ? 2048 : 1024;
        uint32_t min_groups = kOscMinGroupsPerCta;
        if (const char* ev = getenv("FRB_OSC_SPLIT_TARGET")) target = std::max<uint64_t>(1, strtoull(ev, nullptr, 10));   // tuning aids
        if (const char* ev = getenv("FRB_OSC_MIN_GROUPS")) min_groups = (uint32_t)std::max<uint64_t>(1, strtoull(ev, nullptr, 10));
        uint32_t s = 1;
        while (s < 512 && (uint64_t)live_voices * s < target && b->max_groups / (s * 2) >= min_groups) s *= 2;
        b->split = s;
    }

    const uint64_t nn = std::max<uint64_t>(n, 1);
    b->gemm = allow_tensor && osc_gemm_wanted(K, nv, n);
    if (b->rec_cap < nn || (b->gemm && !b->d_rot)) {
        cudaFree(b->d_hot); cudaFree(b->d_anc); cudaFree(b->d_ph); cudaFree(b->d_rot); cudaFree(b->d_tc);
        b->d_hot = b->d_anc = b->d_rot = nullptr; b->d_ph = nullptr; b->d_tc = nullptr; b->rec_cap = 0;
        OC(cudaMalloc(&b->d_hot, nn * sizeof(float4))); OC(cudaMalloc(&b->d_anc, nn * sizeof(float4))); OC(cudaMalloc(&b->d_ph, nn * sizeof(uint4)));
        if (b->gemm) {
            OC(cudaMalloc(&b->d_rot, nn * sizeof(float4)));
            OC(cudaMalloc(&b->d_tc, ((nn + 15) / 16) * GT_REC_WORDS * 16 * sizeof(uint32_t)));
        }
        b->rec_cap = nn;
    }
    const uint32_t nv1 = std::max<uint32_t>(nv, 1);
    if (b->voice_cap < nv1 || (b->gemm && !b->d_vscale)) {
        cudaFree(b->d_grp_begin); cudaFree(b->d_n_grp0); cudaFree(b->d_n_grp); cudaFree(b->d_vscale);
        b->d_grp_begin = b->d_n_grp0 = b->d_n_grp = nullptr; b->d_vscale = nullptr; b->voice_cap = 0;
        OC(cudaMalloc(&b->d_grp_begin, nv1 * sizeof(uint32_t))); OC(cudaMalloc(&b->d_n_grp0, nv1 * sizeof(uint32_t))); OC(cudaMalloc(&b->d_n_grp, nv1 * sizeof(uint32_t)));
        if (b->gemm) OC(cudaMalloc(&b->d_vscale, nv1 * sizeof(float2)));
        b->voice_cap = nv1;
    }
    if (nv) {
        OC(cudaMemcpyAsync(b->d_grp_begin, grp_begin.data(), nv * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        OC(cudaMemcpyAsync(b->d_n_grp0, n_grp0.data(), nv * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        OC(cudaMemcpyAsync(b->d_n_grp, n_grp.data(), nv * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    }
    lap("alloc records");
    if (n) {
        const unsigned fb = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
        osc_fill_kernel<<<fb, 256, 0, stream>>>(n, b->d_hot, b->d_anc, b->d_ph, b->gemm ? b->d_rot : nullptr, b->gemm ? b->d_tc : nullptr);
        OC(cudaGetLastError());
        lap("fill");
        const unsigned gy = std::min<uint32_t>(nv, 32768);
        const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((mx + 255) / 256, std::max<uint64_t>(1, 148ull * 16 / gy)));
        osc_setup_kernel<<<dim3(gx, gy), 256, 0, stream>>>(nv, d_offs, K, b->d_grp_begin, b->d_n_grp0, d_rank, d->sample_rate, d_freq, d_amp,
                                                           d_phase, d_attack, d_tau, b->d_hot, b->d_anc, b->d_ph, b->gemm ? b->d_rot : nullptr, b->gemm ? b->d_tc : nullptr);
        OC(cudaGetLastError());
        if (b->gemm) {
            osc_vscale_kernel<<<std::min<uint32_t>(nv, 148 * 4), 256, 0, stream>>>(nv, d_offs, d_amp, b->d_vscale);
            OC(cudaGetLastError());
        }
    }
    if (!recycle) {   // a first definition is usually the only one: give the scratch (28 B per partial) back
        OC(cudaStreamSynchronize(stream));
        lap("setup");
        cudaFree(b->d_raw); b->d_raw = nullptr; b->raw_cap = 0;
        lap("free raw");
    } else {
        // a re-definition (the per-render input of a synthesis graph): the caller's arrays were consumed by the copies
        // the wait above covered, and the fill / setup kernels are ordered before any render on this stream — the host
        // goes on to enqueue that render instead of waiting for them
        lap("setup (enqueued)");
    }
#undef OC
    return b;
}

// ------------------------------------------------------------------------------------------------------------------
struct OscLaunch {
    const float4* hot; const float4* anc; const uint4* ph;
    const uint32_t* grp_begin; const uint32_t* n_grp0; const uint32_t* n_grp;
    const BufferDesc* bufdesc; uint32_t first_buf;
    unsigned long long lo, hi;       // absolute output window [lo, hi)
    unsigned long long seg0;         // absolute index (t / L) of the first segment
    unsigned nseg;                   // segments per voice
    int L;                           // segment length, multiple of 16
    unsigned split;                  // partial-range split
    float* planes;                   // split > 1: [split][n_voices][plane_len], plane time 0 == seg0 * L
    unsigned long long plane_len;
    unsigned long long plane_off;    // offset of this launch's first segment inside a plane
    unsigned n_voices;
    float max_attack;
};

// One group of K partials over one segment.  ATTACK: apply min(t*invA, 1).
// Both resonator classes run this same recurrence: for a class-1 partial (cos w < 0) the variables
// x_p[n] = (-1)^n x[n], y_p[n] = (-1)^n y[n] obey exactly the class-0 update with the same (a, b, cm1), and segments
// start at even n, so the anchor state needs no change; the kernel flips the sign of the odd samples of the
// accumulator column between the class-0 and class-1 groups instead (osc_kernel).  One instance of the hot loop
// per kernel matters: ptxas' uniform-register budget is per function (2 instances -> 75% promoted, 4 -> none).
// The records come from shared memory at warp-uniform addresses: the compiler keeps a, b, cm1 in uniform
// registers, so the hot loop's FFMAs read two vector registers each instead of three (register-file read
// bandwidth, not the FMA pipe, was the limiter with per-thread coefficient registers: 27% dispatch stalls).
template <int K, bool ATTACK>
__device__ __forceinline__ void osc_group(const float4* hot, const float4* anc, const uint4* ph,
                                          unsigned long long n0, int L, float4* s_acc, unsigned nthr) {
    float x[K], y[K], a[K], b[K], cm1[K], invA[K];
    const float nf = (float)n0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float4 h = hot[k];
        const float4 an = anc[k];
        const uint4 p = ph[k];
        a[k] = h.x; b[k] = h.y; cm1[k] = h.z;
        invA[k] = an.w;
        // exact phase: 64-bit fixed-point turns, wrap-around == range reduction
        // top 32 bits of (inc * n0 + ph0) mod 2^64; the carry out of the low words is dropped (<= 2^-32 turns)
        const unsigned n0_lo = (unsigned)n0, n0_hi = (unsigned)(n0 >> 32);
        const unsigned turns_hi = __umulhi(p.x, n0_lo) + p.y * n0_lo + p.x * n0_hi + p.w;
        const float th = (float)(int)turns_hi * 1.4629180792671596e-9f;           // * 2 pi / 2^32, in [-pi, pi)
        float s, c;
        __sincosf(th, &s, &c);
        // amp * exp(-n/tau) = amp * 2^(-kappa n).  f32 product: the exponent's rounding error X*6e-8 (X = kappa n)
        // gives a relative error X*4e-8 of an envelope that is itself 2^-X of the amplitude — far below 1e-5.
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-an.z * nf));
        e *= an.y;
        y[k] = e * s;
        x[k] = e * fmaf(h.w, s, an.x * c);
    }
    // 16 samples per loop iteration.  ptxas only moves the loop-invariant, warp-uniform coefficients into uniform
    // registers when each is used often enough per iteration (measured on the SASS: 4 samples/iteration -> none,
    // 8 -> about half, 16 -> all of a, b, cm1).  With uniform-register coefficients every hot FFMA reads two vector
    // registers instead of three; the register-file read ports, not the FMA pipe, were the limiter before
    // (27% dispatch stalls, 61% FMA-pipe utilisation in profiles/ncu_osc_r1a_summary.txt).
    constexpr int NV = 4;
    for (int jb = 0; jb < L / (4 * NV); jb++) {
        float4 av[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) av[q] = s_acc[(NV * jb + q) * nthr];
#pragma unroll
        for (int q = 0; q < NV; q++) {
            float r[4] = {av[q].x, av[q].y, av[q].z, av[q].w};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float sum = r[u];
                float tf = 0.f;
                if (ATTACK) tf = nf + (float)(jb * 4 * NV + q * 4 + u);
                if (ATTACK) {
#pragma unroll
                    for (int k = 0; k < K; k++) sum = fmaf(fminf(tf * invA[k], 1.0f), y[k], sum);
                } else {
                    // pairwise tree over the K partials (fixed order): dependency depth log2(K)+1 instead of K
                    float part[K];
#pragma unroll
                    for (int k = 0; k < K; k++) part[k] = y[k];
#pragma unroll
                    for (int w = 1; w < K; w *= 2)
#pragma unroll
                        for (int k = 0; k + w < K; k += 2 * w) part[k] += part[k + w];
                    sum += part[0];
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    x[k] = fmaf(-a[k], y[k], x[k]);
                    const float t = fmaf(cm1[k], y[k], y[k]);
                    y[k] = fmaf(b[k], x[k], t);
                }
                r[u] = sum;
            }
            av[q] = make_float4(r[0], r[1], r[2], r[3]);
        }
#pragma unroll
        for (int q = 0; q < NV; q++) s_acc[(NV * jb + q) * nthr] = av[q];
    }
}

template <int K>
struct OscStage {            // one group's records, staged in shared memory (double buffered)
    float4 hot[K];
    float4 anc[K];
    uint4 ph[K];
};

template <int K, bool ATTACK>
__global__ void __launch_bounds__(OSC_THREADS) osc_kernel(OscLaunch p) {
    extern __shared__ float4 s_acc_all[];
    __shared__ OscStage<K> stage;                            // static: constant addresses -> uniform-register loads
    const unsigned nthr = blockDim.x;
    // Banks of small voices (K < 16) are bound by their ring writes, not by the FMA pipe: they use an odd column stride
    // and a cooperative, coalesced write-out.  The big-bank kernel keeps stride nthr and per-thread columns: its output
    // is 0.2% of its traffic, and the other layout costs it 1.8% (ptxas allocates 70 registers instead of 78 and the
    // hot loop picks up register-bank dispatch stalls: ncu r1i vs r1f).  The register allocation of that kernel is
    // fragile — even replacing `live` below by an equivalent test at the end renumbers the hot loop's registers and
    // costs 1% — so the K = 16 path is kept instruction-for-instruction as measured in r1f (checked by a SASS diff).
    constexpr bool COOP = K < OSC_K;
    const unsigned cstride = COOP ? nthr + 1 : nthr;         // column stride in float4
    const int L = p.L;
    float4* s_acc = s_acc_all + threadIdx.x;                 // column of this thread: s_acc[j4 * cstride]
    const unsigned seg = blockIdx.x * nthr + threadIdx.x;
    const unsigned v = blockIdx.y;
    const unsigned sp = blockIdx.z;
    const bool live = seg < p.nseg;                          // dead threads compute too (dropped at the end)
    const unsigned long long n0 = (p.seg0 + seg) * (unsigned long long)L;
    for (int j4 = 0; j4 < L / 4; j4++) s_acc[j4 * cstride] = make_float4(0.f, 0.f, 0.f, 0.f);

    const unsigned ng = p.n_grp[v], ng0 = p.n_grp0[v], gb = p.grp_begin[v];
    const unsigned g_lo = (unsigned)((unsigned long long)ng * sp / p.split);
    const unsigned g_hi = (unsigned)((unsigned long long)ng * (sp + 1) / p.split);
    const unsigned g_mid = min(max(ng0, g_lo), g_hi);        // first class-1 group of this split

    auto flip_odd = [&]() {                                  // acc[j] *= (-1)^j
        for (int j4 = 0; j4 < L / 4; j4++) {
            float4 a4 = s_acc[j4 * cstride];
            a4.y = -a4.y; a4.w = -a4.w;
            s_acc[j4 * cstride] = a4;
        }
    };
    for (unsigned g = g_lo; g < g_hi; g++) {
        __syncthreads();                                     // everyone is done with the previous group's records
        if (threadIdx.x < K) {                               // threads 0..K-1 stage one record each
            const size_t r = (size_t)(gb + g) * K + threadIdx.x;
            stage.hot[threadIdx.x] = __ldg(p.hot + r);
            stage.anc[threadIdx.x] = __ldg(p.anc + r);
            stage.ph[threadIdx.x] = __ldg(p.ph + r);
        }
        if (g == g_mid) flip_odd();                          // class-0 sums -> alternating-sign domain
        __syncthreads();
        osc_group<K, ATTACK>(stage.hot, stage.anc, stage.ph, n0, L, s_acc, cstride);
    }
    if (g_mid < g_hi) flip_odd();                            // back: out[j] = acc0[j] + (-1)^j acc1[j]
    if constexpr (!COOP) {
        if (!live) return;
        if (p.split == 1) {
            const BufferDesc bd = p.bufdesc[p.first_buf + v];
            for (int j4 = 0; j4 < L / 4; j4++) {
                const unsigned long long t = n0 + 4ull * j4;
                const float4 acc = s_acc[j4 * nthr];
                if (t >= p.lo && t + 4 <= p.hi) {
                    *reinterpret_cast<float4*>(bd.data + (t & bd.mask)) = acc;
                } else {
                    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (t + u >= p.lo && t + u < p.hi) bd.data[(t + u) & bd.mask] = r[u];
                }
            }
        } else {
            float* plane = p.planes + ((size_t)sp * p.n_voices + v) * p.plane_len + p.plane_off + (size_t)seg * L;
            for (int j4 = 0; j4 < L / 4; j4++) *reinterpret_cast<float4*>(plane + 4 * j4) = s_acc[j4 * nthr];
        }
        return;
    }
    // Write-out, cooperative: the CTA's 32 segments are contiguous in time, so quad q of the CTA (4 samples) lives in
    // column q / (L/4), row q % (L/4); consecutive lanes take consecutive quads and every store instruction covers 512
    // contiguous bytes.  (One thread writing its own column = 32 lanes 4*L bytes apart: 5.4 ms instead of 1.9 ms for
    // 4,096 one-partial voices x 480,000 samples.)  The odd column stride keeps these row-major reads conflict-free.
    __syncthreads();
    const int Q = L / 4;
    const unsigned seg_base = blockIdx.x * nthr;
    const unsigned n_live = min(nthr, p.nseg > seg_base ? p.nseg - seg_base : 0u);
    const unsigned long long cta_n0 = (p.seg0 + seg_base) * (unsigned long long)L;
    if (p.split == 1) {
        const BufferDesc bd = p.bufdesc[p.first_buf + v];
        for (unsigned q = threadIdx.x; q < n_live * Q; q += nthr) {
            const unsigned sg = q / Q, j4 = q - sg * Q;
            const float4 acc = s_acc_all[j4 * cstride + sg];
            const unsigned long long t = cta_n0 + 4ull * q;
            if (t >= p.lo && t + 4 <= p.hi) {
                *reinterpret_cast<float4*>(bd.data + (t & bd.mask)) = acc;
            } else {
                const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (t + u >= p.lo && t + u < p.hi) bd.data[(t + u) & bd.mask] = r[u];
            }
        }
    } else {
        float* plane = p.planes + ((size_t)sp * p.n_voices + v) * p.plane_len + p.plane_off + (size_t)seg_base * L;
        for (unsigned q = threadIdx.x; q < n_live * Q; q += nthr) {
            const unsigned sg = q / Q, j4 = q - sg * Q;
            *reinterpret_cast<float4*>(plane + 4ull * q) = s_acc_all[j4 * cstride + sg];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Banks of ONE-partial voices, unsplit (one exciter per voice: BASELINE configs[2]).  Nothing of the multi-partial
// machinery: no accumulator columns (there is nothing to accumulate), no record staging, no sign-flip passes, no
// separate attack kernel.  The oscillator is osc_one_group8 (osc_one.cuh): groups of 8 samples anchored at absolute
// multiples of 8 — the very function the fused exciter -> biquad -> comb kernel (scan.cu) calls, so the two paths
// agree bit for bit.  The kernel is bound by its ring writes (osc_kernel<1> spent 76% of its warp time waiting for the
// 48-byte record at the head of each CTA, with only 12 one-warp CTAs per SM: 17 KB of accumulator columns each): one
// warp per 4,096-sample span of a voice; a thread hands over 32 samples at a time through a 4.5 KB transposition tile
// (32 CTAs per SM) and the warp writes them out as 128-byte rows.
constexpr int ONE_SUB = 32;                  // samples per thread between write-outs (4 groups of 8)
constexpr int ONE_ROW = ONE_SUB + 4;         // tile row stride in floats: 128-bit accesses by row and by column are conflict-free
constexpr int ONE_ROUNDS = 4;                // write-outs per CTA
constexpr int ONE_SPAN = OSC_THREADS * ONE_SUB * ONE_ROUNDS;   // samples per CTA, aligned to absolute multiples of it

__global__ void __launch_bounds__(OSC_THREADS) osc_one_kernel(OscOneSrc src, const BufferDesc* __restrict__ bufdesc, uint32_t first_buf,
                                                              unsigned long long span0, unsigned long long lo, unsigned long long hi) {
    __shared__ __align__(16) float s_t[OSC_THREADS * ONE_ROW];
    const unsigned lane = threadIdx.x, v = blockIdx.y;
    const OscOneVoice o = osc_one_load(src, v);
    const BufferDesc bd = bufdesc[first_buf + v];
    const unsigned long long base = (span0 + blockIdx.x) * (unsigned long long)ONE_SPAN;
    for (int rd = 0; rd < ONE_ROUNDS; rd++) {
        const unsigned long long rb = base + (unsigned long long)rd * (OSC_THREADS * ONE_SUB);   // this round: 1,024 samples
        if (rb >= hi || rb + OSC_THREADS * ONE_SUB <= lo) continue;                               // warp-uniform
#pragma unroll
        for (int g = 0; g < ONE_SUB / 8; g++) {
            float r[8];
            osc_one_group8(o.h, o.an, o.ph, o.flags, rb + (unsigned long long)lane * ONE_SUB + 8u * g, src.max_attack, r);
            *reinterpret_cast<float4*>(s_t + lane * ONE_ROW + 8 * g) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(s_t + lane * ONE_ROW + 8 * g + 4) = make_float4(r[4], r[5], r[6], r[7]);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < ONE_SUB / 4; it++) {                   // 8 quads per row: 4 rows per instruction
            const unsigned row = 4u * it + (lane >> 3), qq = lane & 7u;
            const float4 q = *reinterpret_cast<const float4*>(s_t + row * ONE_ROW + 4 * qq);
            const unsigned long long t = rb + (unsigned long long)row * ONE_SUB + 4 * qq;
            if (t >= lo && t + 4 <= hi) {
                *reinterpret_cast<float4*>(bd.data + (t & bd.mask)) = q;
            } else {
                const float rr[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (t + u >= lo && t + u < hi) bd.data[(t + u) & bd.mask] = rr[u];
            }
        }
        __syncwarp();
    }
}

bool osc_one_source(const OscBankDev& b, OscOneSrc* out) {
    if (b.K != OSC_K_ONE || b.split != 1) return false;
    if (out) *out = OscOneSrc{b.d_hot, b.d_anc, b.d_ph, b.d_grp_begin, b.d_n_grp0, b.d_n_grp, b.max_attack};
    return true;
}

// sums the split planes in fixed order into the voices' rings
__global__ void osc_reduce_kernel(OscLaunch p) {
    const unsigned v = blockIdx.y;
    const BufferDesc bd = p.bufdesc[p.first_buf + v];
    const unsigned long long base = p.seg0 * (unsigned long long)p.L;
    const unsigned long long n4 = p.plane_len / 4;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        float4 s = *reinterpret_cast<const float4*>(p.planes + (size_t)v * p.plane_len + 4 * i);
        for (unsigned sp = 1; sp < p.split; sp++) {
            const float4 q = *reinterpret_cast<const float4*>(p.planes + ((size_t)sp * p.n_voices + v) * p.plane_len + 4 * i);
            s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
        }
        const unsigned long long t = base + 4 * i;
        if (t >= p.lo && t + 4 <= p.hi) {
            *reinterpret_cast<float4*>(bd.data + (t & bd.mask)) = s;
        } else {
            const float r[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (t + u >= p.lo && t + u < p.hi) bd.data[(t + u) & bd.mask] = r[u];
        }
    }
}

static cudaError_t launch_osc_range(const OscBankDev& b, const BufferDesc* d_bufdesc, uint32_t first_buf, uint64_t lo, uint64_t hi,
                                    uint32_t anchor, int sm_count, cudaStream_t stream, uint64_t* n_launches, unsigned which = 0,
                                    bool overlap_reduce = false);

// per device, called when a renderer is created on it: the accumulator columns need up to 32 KB of dynamic shared memory
cudaError_t osc_init_device() {
    const int mx = OSC_LMAX * (OSC_THREADS + 1) * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(osc_kernel<OSC_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_kernel<OSC_K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_kernel<OSC_K_SMALL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_kernel<OSC_K_SMALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_kernel<OSC_K_ONE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_kernel<OSC_K_ONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(osc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM);
    return e;
}

static cudaError_t launch_osc_fma(const OscBankDev& b, const BufferDesc* d_bufdesc, uint32_t first_buf, uint64_t lo, uint64_t hi,
                                  uint32_t anchor, int sm_count, cudaStream_t stream, uint64_t* n_launches);

cudaError_t launch_osc(const OscBankDev& b, const BufferDesc* d_bufdesc, uint32_t first_buf, uint64_t lo, uint64_t hi,
                       uint32_t anchor, int sm_count, cudaStream_t stream, uint64_t* n_launches, uint64_t* n_tensor_launches) {
    if (n_launches) *n_launches = 0;
    if (n_tensor_launches) *n_tensor_launches = 0;
    if (hi <= lo || b.n_voices == 0) return cudaSuccess;
    if (!b.gemm) return launch_osc_fma(b, d_bufdesc, first_buf, lo, hi, anchor, sm_count, stream, n_launches);
    // K1G from the first 128-sample block past every attack ramp; the resonator kernels (with their ramp instance) before it
    const uint64_t ramp_end = ((uint64_t)std::ceil((double)std::max(b.max_attack, 0.0f)) + GM_N - 1) / GM_N * GM_N;
    // The ramp region is a few hundred samples — a handful of one-warp CTAs per voice walking every partial, 1.2 ms on cfg4 —
    // and writes other samples than the matrix-product kernel: it runs on a stream of its own beside it.
    const bool beside = lo < ramp_end && hi > ramp_end;
    if (beside && !b.gside) {
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        cudaError_t e = cudaStreamCreateWithPriority(&b.gside, cudaStreamNonBlocking, hi_prio);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_gfork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_gjoin, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    if (lo < ramp_end) {
        cudaStream_t rs = stream;
        if (beside) { cudaEventRecord(b.ev_gfork, stream); cudaStreamWaitEvent(b.gside, b.ev_gfork, 0); rs = b.gside; }
        cudaError_t e = launch_osc_fma(b, d_bufdesc, first_buf, lo, std::min(hi, ramp_end), anchor, sm_count, rs, n_launches);
        if (e != cudaSuccess) return e;
        if (beside) cudaEventRecord(b.ev_gjoin, b.gside);
    }
    if (hi > ramp_end) {
        OscGemmLaunch p;
        p.anc = b.d_anc; p.ph = b.d_ph; p.rot = b.d_rot; p.tc = b.d_tc;
        p.grp_begin = b.d_grp_begin; p.n_grp = b.d_n_grp; p.vscale = b.d_vscale;
        p.bufdesc = d_bufdesc; p.first_buf = first_buf;
        p.lo = std::max(lo, ramp_end); p.hi = hi;
        p.K = b.K;
        const uint64_t tile_len = (uint64_t)GM_M * GM_N;
        p.tile0 = p.lo / tile_len;
        const uint64_t n_tiles = (hi + tile_len - 1) / tile_len - p.tile0;
        for (uint64_t t = 0; t < n_tiles; t += 65535) {                  // (grid.x is the wide dimension: tiles there, voices in y)
            OscGemmLaunch q = p;
            q.tile0 = p.tile0 + t;
            const unsigned nx = (unsigned)std::min<uint64_t>(n_tiles - t, 65535);
            if (osc_gemm_tcgen05()) osc_tc_kernel<<<dim3(nx, b.n_voices), GT_THREADS, GT_SMEM, stream>>>(q);
            else osc_gemm_kernel<<<dim3(nx, b.n_voices), GM_THREADS, GM_SMEM, stream>>>(q);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            if (n_launches) (*n_launches)++;
            if (n_tensor_launches) (*n_tensor_launches)++;
        }
    }
    if (beside) cudaStreamWaitEvent(stream, b.ev_gjoin, 0);
    return cudaSuccess;
}

static cudaError_t launch_osc_fma(const OscBankDev& b, const BufferDesc* d_bufdesc, uint32_t first_buf, uint64_t lo, uint64_t hi,
                                  uint32_t anchor, int sm_count, cudaStream_t stream, uint64_t* n_launches) {
    if (hi <= lo || b.n_voices == 0) return cudaSuccess;
    // split banks stage their partial-range planes in scratch: bound it by rendering 64 Ki samples at a time
    if (b.split == 1) return launch_osc_range(b, d_bufdesc, first_buf, lo, hi, anchor, sm_count, stream, n_launches);
    // ... in sub-blocks whose planes stay under 1 GiB (at least 64 Ki samples)
    uint64_t sub = 1ull << 20;
    while (sub > (1ull << 16) && sub * b.split * b.n_voices > (1ull << 28)) sub >>= 1;
    // The plane reduce is HBM-bound (it reads split x voices x samples x 4 B) and the main kernel FMA-bound, so the reduce
    // of sub-block k runs on a stream of its own BESIDE the main kernel of sub-block k + 1, the two alternating between two
    // plane buffers: only the last sub-block's reduce is left on the critical path (8 voices per GPU, split 128: 0.35 ms of
    // reduce per 31 ms step were all exposed before — the whole of the 8-GPU scaling loss, profiles/r2f_timeline_n8.json).
    static const bool overlap_env = [] { const char* e = getenv("FRB_OSC_REDUCE_OVERLAP"); return !e || e[0] != '0'; }();   // measurement knobs
    static const int main_streams = [] { const char* e = getenv("FRB_OSC_MAIN_STREAMS"); return e ? atoi(e) : 2; }();
    static const uint64_t min_ranges = [] { const char* e = getenv("FRB_OSC_MIN_RANGES"); return e ? strtoull(e, nullptr, 10) : 1ull; }();
    // (more, shorter sub-blocks would shrink the one reduce left on the critical path, but every launch costs a ramp and a
    // tail: 8 sub-blocks instead of 2 measured 31.04 vs 30.73 ms on an 8-voice shard; FRB_OSC_MIN_RANGES forces them)
    if (overlap_env) while (sub > (1ull << 16) && (hi - lo) / sub < min_ranges) sub >>= 1;
    const bool overlap = overlap_env && (lo / sub != (hi - 1) / sub);     // a single sub-block has nothing to overlap with
    if (overlap && !b.red) {
        cudaError_t e = cudaStreamCreateWithFlags(&b.red, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b.main2, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_enter, cudaEventDisableTiming);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&b.ev_main[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_red[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) return e;
    }
    // Main kernels of consecutive sub-blocks alternate between the caller's stream and a second one (they write different
    // plane buffers): the head of sub-block k + 1 fills the SMs the last CTAs of sub-block k leave idle, so only the
    // last launch of a render has a tail.
    if (overlap) {
        // both plane buffers up front, for the longest sub-block: nothing is reallocated while another stream uses it
        int L = anchor ? (int)anchor : 128;
        L = std::max(16, std::min(OSC_LMAX, (L / 16) * 16));
        const uint64_t longest = std::min<uint64_t>(sub, hi - lo);
        const uint64_t need = (uint64_t)b.split * b.n_voices * ((longest + L - 1) / L + 1) * L;
        if (need > b.planes_cap) {
            cudaStreamSynchronize(stream); cudaStreamSynchronize(b.red); cudaStreamSynchronize(b.main2);
            cudaFree(b.d_planes); b.d_planes = nullptr; b.planes_cap = 0;
            cudaError_t e = cudaMalloc(&b.d_planes, 2 * need * sizeof(float));
            if (e != cudaSuccess) return e;
            b.planes_cap = need;
        }
    }
    const bool two_mains = overlap && main_streams >= 2;
    if (two_mains) { cudaEventRecord(b.ev_enter, stream); cudaStreamWaitEvent(b.main2, b.ev_enter, 0); }
    unsigned k = 0;
    for (uint64_t c0 = lo; c0 < hi; k++) {
        const uint64_t c1 = std::min(hi, (c0 / sub + 1) * sub);
        cudaStream_t ms = (two_mains && (k & 1u)) ? b.main2 : stream;
        cudaError_t e = launch_osc_range(b, d_bufdesc, first_buf, c0, c1, anchor, sm_count, ms, n_launches, k & 1u, overlap);
        if (e != cudaSuccess) return e;
        c0 = c1;
    }
    for (int i = 0; i < 2; i++)                                   // the rings are complete when both reduce chains are
        if (b.red_pending[i]) { cudaStreamWaitEvent(stream, b.ev_red[i], 0); b.red_pending[i] = false; }
    return cudaGetLastError();
}

static cudaError_t launch_osc_range(const OscBankDev& b, const BufferDesc* d_bufdesc, uint32_t first_buf, uint64_t lo, uint64_t hi,
                                    uint32_t anchor, int sm_count, cudaStream_t stream, uint64_t* n_launches, unsigned which,
                                    bool overlap_reduce) {
    OscOneSrc one;
    if (osc_one_source(b, &one)) {                           // one-partial voices: their own kernel, anchors every 8 samples
        const uint64_t span0 = lo / ONE_SPAN, n_span = (hi + ONE_SPAN - 1) / ONE_SPAN - span0;
        osc_one_kernel<<<dim3((unsigned)n_span, b.n_voices), OSC_THREADS, 0, stream>>>(one, d_bufdesc, first_buf, span0, lo, hi);
        if (n_launches) (*n_launches)++;
        return cudaGetLastError();
    }
    int L = anchor ? (int)anchor : 128;
    L = std::max(16, std::min(OSC_LMAX, (L / 16) * 16));
    OscLaunch p;
    p.hot = b.d_hot; p.anc = b.d_anc; p.ph = b.d_ph;
    p.grp_begin = b.d_grp_begin; p.n_grp0 = b.d_n_grp0; p.n_grp = b.d_n_grp;
    p.bufdesc = d_bufdesc; p.first_buf = first_buf;
    p.lo = lo; p.hi = hi;
    p.L = L;
    p.seg0 = lo / L;
    p.nseg = (unsigned)((hi + L - 1) / L - p.seg0);
    p.split = b.split;
    p.n_voices = b.n_voices;
    p.max_attack = b.max_attack;
    p.plane_len = (unsigned long long)p.nseg * L;
    p.planes = nullptr;
    p.plane_off = 0;
    if (p.split > 1) {
        uint64_t need = (uint64_t)p.split * b.n_voices * p.plane_len;
        if (need > b.planes_cap) {
            if (b.d_planes) {
                cudaStreamSynchronize(stream);
                if (b.red) cudaStreamSynchronize(b.red);
                cudaFree(b.d_planes); b.d_planes = nullptr; b.planes_cap = 0;
            }
            cudaError_t e = cudaMalloc(&b.d_planes, 2 * need * sizeof(float));
            if (e != cudaSuccess) return e;
            b.planes_cap = need;
        }
        p.planes = b.d_planes + (size_t)which * b.planes_cap;
        // this buffer was last read by the reduce of two sub-blocks ago
        if (overlap_reduce && b.red_pending[which]) { cudaStreamWaitEvent(stream, b.ev_red[which], 0); b.red_pending[which] = false; }
    }
    const unsigned threads = OSC_THREADS;
    // L/4 rows of float4 columns: threads + 1 of them for the small-voice kernels (odd stride), threads for K = 16
    const size_t smem = (size_t)L * (threads + (b.K < OSC_K ? 1 : 0)) * sizeof(float);
    // Segments that start below max_attack need the attack ramp: they get their own (slower) kernel so that each
    // kernel contains exactly one instance of the hot loop.
    unsigned n_att = 0;
    if (b.max_attack > 0.f) {
        unsigned long long first_plain = ((unsigned long long)std::ceil((double)b.max_attack) + L - 1) / L;   // first segment with n0 >= max_attack
        if (first_plain > p.seg0) n_att = (unsigned)std::min<unsigned long long>(first_plain - p.seg0, p.nseg);
    }
    const unsigned nseg_total = p.nseg;
    const bool fork = n_att && nseg_total > n_att;
    if (fork && !b.side) {
        cudaError_t e = cudaStreamCreateWithFlags(&b.side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    if (fork) cudaEventRecord(b.ev_fork, stream);            // what the side stream must wait for: everything before this range
    if (nseg_total > n_att) {
        OscLaunch q = p;
        q.seg0 = p.seg0 + n_att;
        q.nseg = nseg_total - n_att;
        q.plane_off = (unsigned long long)n_att * L;
        dim3 grid((q.nseg + threads - 1) / threads, b.n_voices, p.split);
        if (b.K == OSC_K) osc_kernel<OSC_K, false><<<grid, threads, smem, stream>>>(q);
        else if (b.K == OSC_K_SMALL) osc_kernel<OSC_K_SMALL, false><<<grid, threads, smem, stream>>>(q);
        else osc_kernel<OSC_K_ONE, false><<<grid, threads, smem, stream>>>(q);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    if (n_att) {
        // The attack region is a few hundred samples: at the main kernel's L it is a handful of live threads in one warp
        // per (voice, split) walking every group of the split — the longest CTA of the launch (the ramp loop is slower),
        // and for a bank of few voices, whose launch is a single wave, the whole kernel's duration.  It therefore runs in
        // SHORTER segments (32 samples, anchored at absolute multiples of 32 like every segment): four times the live
        // threads, a quarter of the critical path.  A function of the bank and of absolute time only, so results still
        // do not depend on how a render is cut into blocks.
        OscLaunch q = p;
        const int La = (L % 32 == 0 && !getenv("FRB_OSC_ATTACK_SAME_L")) ? 32 : L;
        const unsigned r = (unsigned)(L / La);
        q.L = La;
        q.seg0 = p.seg0 * r;
        q.nseg = n_att * r;
        cudaStream_t st = stream;
        if (fork) {   // a handful of warps walking every partial: beside the main kernel (launched above), not before it
            cudaStreamWaitEvent(b.side, b.ev_fork, 0);
            st = b.side;
        }
        const size_t smem_a = (size_t)La * (threads + (b.K < OSC_K ? 1 : 0)) * sizeof(float);
        dim3 grid((q.nseg + threads - 1) / threads, b.n_voices, p.split);
        if (b.K == OSC_K) osc_kernel<OSC_K, true><<<grid, threads, smem_a, st>>>(q);
        else if (b.K == OSC_K_SMALL) osc_kernel<OSC_K_SMALL, true><<<grid, threads, smem_a, st>>>(q);
        else osc_kernel<OSC_K_ONE, true><<<grid, threads, smem_a, st>>>(q);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    if (fork) {
        cudaEventRecord(b.ev_join, b.side);
        cudaStreamWaitEvent(stream, b.ev_join, 0);
    }
    cudaError_t e = cudaSuccess;
    if (p.split > 1) {
        unsigned long long n4 = p.plane_len / 4;
        unsigned bx = (unsigned)std::min<unsigned long long>((n4 + 255) / 256, (unsigned long long)sm_count * 4);
        cudaStream_t rs = stream;
        if (overlap_reduce) {                                        // beside the next sub-block's main kernel
            cudaEventRecord(b.ev_main[which], stream);
            cudaStreamWaitEvent(b.red, b.ev_main[which], 0);
            rs = b.red;
        }
        osc_reduce_kernel<<<dim3(bx, b.n_voices), 256, 0, rs>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (overlap_reduce) { cudaEventRecord(b.ev_red[which], b.red); b.red_pending[which] = true; }
        if (n_launches) (*n_launches)++;
    }
    return cudaSuccess;
}

}  // namespace frb
