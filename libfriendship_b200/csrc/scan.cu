// scan.cu — K4: Direct Form linear recurrences for sm_100a (extension nodes FRB_KIND_DIRECTFORM / FRB_KIND_FBDELAY;
// not in the reference: SURVEY.md F2 — the reference rejects cycles, so feedback only exists as these nodes).
//
// Bound: HBM.  Algorithmic traffic = 8 B per lane-sample for each node (read x, write y): each kernel reads its
// input ring once and writes its output ring once; no intermediate touches HBM.  A biquad that feeds a comb lane for
// lane (and nothing else) runs fused with it: 8 B per lane-sample for the pair (dfcomb_kernel below).
//
// DirectForm (biquad, Direct Form I), per lane:  y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2].
//   One WARP owns one lane and walks the block in tiles of 32 threads x 8 samples.  Inside a tile the recurrence is
//   evaluated as a parallel prefix scan over the recurrence's transfer matrices (biquad_tile): every thread runs its
//   8 samples from a zero state (lane 0 from the carry), the 2-vectors of end states are combined with a Kogge-Stone
//   scan whose operator is "multiply by A^(8*2^k) and add" (A = [[-a1, -a2], [1, 0]]; the powers come from an fp64
//   setup at definition time), then every thread re-runs its 8 samples from its true initial state.  The carry
//   between tiles and between launches is just the last two samples of x and y, re-read from the rings, so
//   consecutive blocks continue exactly.
//
// FbDelay, per lane:  y[n] = x[n] + g y[n-D].  D independent first-order recurrences with stride D: one CTA per
//   lane steps through time D samples at a time, all D phases in parallel; every output is computed in the same
//   order and with the same two roundings (mul, add) as a sequential f32 evaluation, so it is bit-exact to one.
#include "scan.cuh"
#include "osc_one.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace frb {

constexpr int FB_THREADS = 128;
constexpr int DF_CTA_THREADS = 128;   // 4 lanes per CTA, one warp each
constexpr int DF_PER_THREAD = 8;     // samples per thread of a tile (the exciter-fused chain kernel uses 16, see there)
constexpr int DF_LEVELS = 8;      // stored powers A^(8*2^k), k < 8 (the warp scan uses k < 5)

struct DirectFormDev {
    uint32_t n_lanes = 0;
    float* d_coef = nullptr;      // [n_lanes][5]  b0 b1 b2 a1 a2
    float* d_pow = nullptr;       // [n_lanes][DF_LEVELS][4]  A^(8 * 2^k), row-major 2x2
    ~DirectFormDev() { cudaFree(d_coef); cudaFree(d_pow); }
};
struct FbDelayDev {
    uint32_t n_lanes = 0;
    uint64_t max_delay = 0;
    uint64_t min_delay = ~0ull;
    uint32_t* d_delay = nullptr;
    float* d_gain = nullptr;
    uint32_t* d_order = nullptr;  // fused chain: which lane the w-th warp of the launch takes (see chain_lane_order)
    ~FbDelayDev() { cudaFree(d_delay); cudaFree(d_gain); cudaFree(d_order); }
};

struct ChainStateDev {            // biquad carry of a fused DirectForm -> FbDelay chain: {x[n-1], x[n-2], y[n-1], y[n-2]} at `time`
    float4* d_state = nullptr;
    uint32_t n_lanes = 0;
    uint64_t time = 0;
    ~ChainStateDev() { cudaFree(d_state); }
};

uint32_t directform_lanes(const DirectFormDev& f) { return f.n_lanes; }
uint32_t fbdelay_lanes(const FbDelayDev& f) { return f.n_lanes; }
uint64_t fbdelay_max_delay(const FbDelayDev& f) { return f.max_delay; }
uint64_t fbdelay_min_delay(const FbDelayDev& f) { return f.min_delay; }

std::shared_ptr<DirectFormDev> directform_create(const frb_directform_desc* d, cudaStream_t stream, std::string* err) {
    auto fail = [&](const std::string& m) { if (err) *err = m; return std::shared_ptr<DirectFormDev>(); };
    if (d->n_lanes && (!d->b0 || !d->b1 || !d->b2 || !d->a1 || !d->a2)) return fail("directform: null array");
    auto f = std::make_shared<DirectFormDev>();
    f->n_lanes = d->n_lanes;
    const size_t n = std::max<uint32_t>(d->n_lanes, 1);
    std::vector<float> coef(n * 5, 0.f), pw(n * DF_LEVELS * 4, 0.f);
    for (uint32_t l = 0; l < d->n_lanes; l++) {
        coef[l * 5 + 0] = d->b0[l]; coef[l * 5 + 1] = d->b1[l]; coef[l * 5 + 2] = d->b2[l];
        coef[l * 5 + 3] = d->a1[l]; coef[l * 5 + 4] = d->a2[l];
        // A = [[-a1, -a2], [1, 0]] acts on (y[n-1], y[n-2]); powers in fp64
        double m[4] = {-(double)d->a1[l], -(double)d->a2[l], 1.0, 0.0};
        double p[4] = {1, 0, 0, 1};
        for (int i = 0; i < DF_PER_THREAD; i++) {   // p = A^8
            double q[4] = {m[0] * p[0] + m[1] * p[2], m[0] * p[1] + m[1] * p[3], m[2] * p[0] + m[3] * p[2], m[2] * p[1] + m[3] * p[3]};
            for (int j = 0; j < 4; j++) p[j] = q[j];
        }
        for (int k = 0; k < DF_LEVELS; k++) {
            for (int j = 0; j < 4; j++) pw[((size_t)l * DF_LEVELS + k) * 4 + j] = (float)p[j];
            double q[4] = {p[0] * p[0] + p[1] * p[2], p[0] * p[1] + p[1] * p[3], p[2] * p[0] + p[3] * p[2], p[2] * p[1] + p[3] * p[3]};
            for (int j = 0; j < 4; j++) p[j] = q[j];
        }
    }
    if (cudaMalloc(&f->d_coef, coef.size() * sizeof(float)) != cudaSuccess || cudaMalloc(&f->d_pow, pw.size() * sizeof(float)) != cudaSuccess)
        return fail("directform: out of device memory");
    cudaMemcpyAsync(f->d_coef, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(f->d_pow, pw.data(), pw.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail("directform: upload failed");
    return f;
}

// The fused chain kernel gives every lane one warp for the whole launch and all its CTAs are resident at once, so a launch
// lasts as long as the busiest SM.  A comb shorter than a tile costs more per sample (taps out of the tile itself: a
// barrier per chunk; below 128 narrower chunks), and lanes of one class tend to be neighbours (cfg3: D = 100 + v mod 900
// puts 156 short combs in a row = 39 whole CTAs, two of them on some SMs and none on others: 3.25 against 3.06 ms with
// uniform long delays).  Lanes are therefore sorted by cost class and dealt out across the CTAs, warp w of CTA c taking
// sorted[w * n_ctas + (c - 37 w) mod n_ctas]: every CTA gets at most one lane more of a class than any other.  A lane's
// arithmetic does not depend on which warp runs it.  cfg3, ring-fed: 3.27 -> 3.125 ms (3.08 with uniform long delays);
// FRB_K4_LANE_ORDER=0 keeps the identity (measurement knob; profiles/k4_lane_order_ab_r2.jsonl).
static std::vector<uint32_t> chain_lane_order(const uint32_t* delay, uint32_t n_lanes) {
    std::vector<uint32_t> order(n_lanes);
    for (uint32_t i = 0; i < n_lanes; i++) order[i] = i;
    static const bool on = [] { const char* e = getenv("FRB_K4_LANE_ORDER"); return !(e && e[0] == '0'); }();
    if (!on || n_lanes == 0) return order;
    auto cls = [&](uint32_t l) { const uint32_t D = delay[l]; return D < 64 ? 0 : D < 128 ? 1 : D < 256 ? 2 : 3; };
    bool mixed = false;
    for (uint32_t i = 1; i < n_lanes && !mixed; i++) mixed = cls(i) != cls(0);
    if (!mixed) return order;                                      // one class: nothing to balance, neighbours stay together
    std::vector<uint32_t> sorted = order;
    std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) { return cls(a) < cls(b); });
    const uint32_t per_cta = DF_CTA_THREADS / 32, n_ctas = (n_lanes + per_cta - 1) / per_cta;
    const uint32_t rot = 37;                                       // a CTA's lanes (their rings) not a power of two apart: 0.5%
    // slots of the last CTA beyond n_lanes do not exist: walk the (w, c) grid in sorted order and skip them
    uint32_t i = 0;
    for (uint32_t w = 0; w < per_cta; w++)
        for (uint32_t k = 0; k < n_ctas; k++) {
            const uint32_t c = (k + rot * w) % n_ctas;
            const uint32_t slot = c * per_cta + w;
            if (slot < n_lanes) order[slot] = sorted[i++];
        }
    return order;
}

std::shared_ptr<FbDelayDev> fbdelay_create(const frb_fbdelay_desc* d, cudaStream_t stream, std::string* err) {
    auto fail = [&](const std::string& m) { if (err) *err = m; return std::shared_ptr<FbDelayDev>(); };
    if (d->n_lanes && (!d->delay || !d->gain)) return fail("fbdelay: null array");
    auto f = std::make_shared<FbDelayDev>();
    f->n_lanes = d->n_lanes;
    for (uint32_t i = 0; i < d->n_lanes; i++) {
        if (d->delay[i] < 1) return fail("fbdelay: delay must be >= 1");
        f->max_delay = std::max<uint64_t>(f->max_delay, d->delay[i]);
        f->min_delay = std::min<uint64_t>(f->min_delay, d->delay[i]);
    }
    const size_t n = std::max<uint32_t>(d->n_lanes, 1);
    if (cudaMalloc(&f->d_delay, n * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&f->d_gain, n * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&f->d_order, n * sizeof(uint32_t)) != cudaSuccess)
        return fail("fbdelay: out of device memory");
    const std::vector<uint32_t> order = chain_lane_order(d->delay, d->n_lanes);
    if (d->n_lanes) {
        cudaMemcpyAsync(f->d_delay, d->delay, d->n_lanes * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(f->d_gain, d->gain, d->n_lanes * sizeof(float), cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(f->d_order, order.data(), d->n_lanes * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail("fbdelay: upload failed");
    return f;
}

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ring_at(const BufferDesc& b, long long t) {
    return t >= 0 ? b.data[(unsigned long long)t & b.mask] : 0.0f;    // every signal is 0 before t = 0
}

// The biquad over one tile of 32 threads x 8 samples, shared by directform_kernel and dfcomb_kernel (same bits).
//   x[0..9]: this thread's samples x[2..9] preceded by the two before them (x[0] = x[t0-2], x[1] = x[t0-1]);
//   (cy1, cy2) = (y[tb-1], y[tb-2]), the carry into the tile (read by lane 0 only);  P(k) = A^(8*2^k), row-major 2x2.
// Every thread runs its 8 samples from a zero state — lane 0 from the carry — the end states are combined by a
// Kogge-Stone scan with the operator "v_i <- v_i + A^(8*2^k) v_(i-2^k)", which leaves in lane i the true state at the end
// of its run; its left neighbour's is the state it re-runs from.  Out: yv[0..7], (y1, y2) = (y[t0+7], y[t0+6]).
template <int S, class PF>
__device__ __forceinline__ void biquad_tile(const float (&x)[S + 2], float cy1, float cy2, PF P, float b0, float b1,
                                            float b2, float a1, float a2, unsigned wl, float (&yv)[S],
                                            float& y1, float& y2) {
    float u[S];
#pragma unroll
    for (int j = 0; j < S; j++) u[j] = fmaf(b2, x[j], fmaf(b1, x[j + 1], b0 * x[j + 2]));
    const float s1 = wl ? 0.f : cy1, s2 = wl ? 0.f : cy2;
    float e1 = s1, e2 = s2;
    // the y[n-2] term is folded in first: ONE FMA per step on the recurrence's critical path
#pragma unroll
    for (int j = 0; j < S; j++) {
        const float y = fmaf(-a1, e1, fmaf(-a2, e2, u[j]));
        e2 = e1; e1 = y;
    }
    float v1 = e1, v2 = e2;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const float4 Pk = P(k);
        float o1 = __shfl_up_sync(0xffffffffu, v1, 1u << k), o2 = __shfl_up_sync(0xffffffffu, v2, 1u << k);
        const bool on = wl >= (1u << k);                            // lanes below 2^k have nothing to their left at this level
        o1 = on ? o1 : 0.f; o2 = on ? o2 : 0.f;
        const float n1 = fmaf(Pk.x, o1, fmaf(Pk.y, o2, v1));
        v2 = fmaf(Pk.z, o1, fmaf(Pk.w, o2, v2));
        v1 = n1;
    }
    const float p1 = __shfl_up_sync(0xffffffffu, v1, 1), p2 = __shfl_up_sync(0xffffffffu, v2, 1);
    y1 = wl ? p1 : s1;                                              // (y[t0-1], y[t0-2])
    y2 = wl ? p2 : s2;
#pragma unroll
    for (int j = 0; j < S; j++) {
        const float y = fmaf(-a1, y1, fmaf(-a2, y2, u[j]));
        yv[j] = y; y2 = y1; y1 = y;
    }
}

// One WARP per lane (4 lanes per CTA): tiles of 32 threads x 8 samples, warp-level scan only, no block barriers.
__global__ void __launch_bounds__(DF_CTA_THREADS)
directform_kernel(const float* __restrict__ coef, const float* __restrict__ pw, const BufferDesc* __restrict__ bufdesc,
                  const uint32_t* __restrict__ in_bufs, uint32_t first_out_buf, unsigned n_lanes,
                  unsigned long long lo, unsigned long long hi) {
    const unsigned wl = threadIdx.x & 31;
    const unsigned lane = blockIdx.x * (DF_CTA_THREADS / 32) + (threadIdx.x >> 5);
    if (lane >= n_lanes) return;                                   // whole warp exits together
    const BufferDesc xin = bufdesc[in_bufs[lane]];
    const BufferDesc yout = bufdesc[first_out_buf + lane];
    const float b0 = coef[lane * 5 + 0], b1 = coef[lane * 5 + 1], b2 = coef[lane * 5 + 2];
    const float a1 = coef[lane * 5 + 3], a2 = coef[lane * 5 + 4];
    // A^(8*2^k), k < 5, in registers (warp-uniform)
    float P[5][4];
#pragma unroll
    for (int k = 0; k < 5; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) P[k][j] = pw[((size_t)lane * DF_LEVELS + k) * 4 + j];

    // carries: the last two outputs and inputs before the tile
    float y1c = ring_at(yout, (long long)lo - 1), y2c = ring_at(yout, (long long)lo - 2);
    float x1c = ring_at(xin, (long long)lo - 1), x2c = ring_at(xin, (long long)lo - 2);
    const bool vec = (lo % 4 == 0);                                // 128-bit accesses when the block start is aligned
    constexpr int TILE = 32 * DF_PER_THREAD;
    for (unsigned long long tb = lo; tb < hi; tb += TILE) {
        const unsigned long long t0 = tb + (unsigned long long)wl * DF_PER_THREAD;
        float x[DF_PER_THREAD + 2];
        if (vec && t0 + DF_PER_THREAD <= hi) {
#pragma unroll
            for (int jq = 0; jq < DF_PER_THREAD / 4; jq++) {
                const float4 v = *reinterpret_cast<const float4*>(xin.data + ((t0 + 4 * jq) & xin.mask));
                x[2 + 4 * jq] = v.x; x[3 + 4 * jq] = v.y; x[4 + 4 * jq] = v.z; x[5 + 4 * jq] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < DF_PER_THREAD; j++) x[2 + j] = (t0 + j < hi) ? xin.data[(t0 + j) & xin.mask] : 0.0f;
        }
        // two samples of history from the previous thread (lane 0: from the carry)
        const float px1 = __shfl_up_sync(0xffffffffu, x[DF_PER_THREAD + 1], 1), px2 = __shfl_up_sync(0xffffffffu, x[DF_PER_THREAD], 1);
        x[1] = wl ? px1 : x1c;
        x[0] = wl ? px2 : x2c;
        float yv[DF_PER_THREAD], y1, y2;
        biquad_tile<DF_PER_THREAD>(x, y1c, y2c, [&](int k) { return make_float4(P[k][0], P[k][1], P[k][2], P[k][3]); }, b0, b1, b2, a1, a2, wl, yv, y1, y2);
        if (vec && t0 + DF_PER_THREAD <= hi) {
#pragma unroll
            for (int jq = 0; jq < DF_PER_THREAD / 4; jq++)
                *reinterpret_cast<float4*>(yout.data + ((t0 + 4 * jq) & yout.mask)) = make_float4(yv[4 * jq], yv[4 * jq + 1], yv[4 * jq + 2], yv[4 * jq + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < DF_PER_THREAD; j++)
                if (t0 + j < hi) yout.data[(t0 + j) & yout.mask] = yv[j];
        }
        // carries for the next tile come from the last thread
        y1c = __shfl_sync(0xffffffffu, y1, 31); y2c = __shfl_sync(0xffffffffu, y2, 31);
        x1c = __shfl_sync(0xffffffffu, x[DF_PER_THREAD + 1], 31); x2c = __shfl_sync(0xffffffffu, x[DF_PER_THREAD], 31);
    }
}

// Phase j of the delay line (t = lo + j + k D) only ever depends on itself, so one thread carries y[t - D] in a
// register and walks k; the D phases of a lane run in parallel with no barrier at all.  Loads of x are independent of
// the recurrence and are issued four steps ahead.
__global__ void __launch_bounds__(FB_THREADS)
fbdelay_kernel(const uint32_t* __restrict__ delay, const float* __restrict__ gain, const BufferDesc* __restrict__ bufdesc,
               const uint32_t* __restrict__ in_bufs, uint32_t first_out_buf, unsigned long long lo, unsigned long long hi) {
    const unsigned lane = blockIdx.y;
    const BufferDesc xin = bufdesc[in_bufs[lane]];
    const BufferDesc yout = bufdesc[first_out_buf + lane];
    const unsigned long long D = delay[lane];
    const float g = gain[lane];
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < D;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long t = lo + j;
        if (t >= hi) break;
        float yp = (t >= D) ? yout.data[(t - D) & yout.mask] : 0.0f;
        for (; t + 3 * D < hi; t += 4 * D) {
            const float x0 = xin.data[t & xin.mask], x1 = xin.data[(t + D) & xin.mask];
            const float x2 = xin.data[(t + 2 * D) & xin.mask], x3 = xin.data[(t + 3 * D) & xin.mask];
            // two roundings per step, like a sequential f32 evaluation: x + (g * y[n-D])
            const float y0 = __fadd_rn(x0, __fmul_rn(g, yp));
            const float y1 = __fadd_rn(x1, __fmul_rn(g, y0));
            const float y2 = __fadd_rn(x2, __fmul_rn(g, y1));
            const float y3 = __fadd_rn(x3, __fmul_rn(g, y2));
            yout.data[t & yout.mask] = y0; yout.data[(t + D) & yout.mask] = y1;
            yout.data[(t + 2 * D) & yout.mask] = y2; yout.data[(t + 3 * D) & yout.mask] = y3;
            yp = y3;
        }
        for (; t < hi; t += D) {
            yp = __fadd_rn(xin.data[t & xin.mask], __fmul_rn(g, yp));
            yout.data[t & yout.mask] = yp;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Fused chain: DirectForm lane l -> FbDelay lane l, when nothing else reads the biquad's output.  8 B per lane-sample of
// HBM traffic (read x once, write z once) instead of 16 B for the two separate kernels: the biquad's y never leaves the
// SM.  One warp per lane, tiles of 32 threads x S samples (S = 8; 16 when the input is an oscillator, see the kernel):
//   * tiles travel between HBM and shared memory as quads (4 samples) in coalesced rows — lane i moves quads i, 32 + i,
//     ... : 512 contiguous bytes per instruction — with cp.async (no registers held): the x tile of the NEXT iteration
//     and this tile's comb taps are in flight during the biquad;
//   * the biquad is biquad_tile, shared with directform_kernel (same bits at S = 8); a thread's own samples are consecutive quads
//     of the tile, which sits in shared memory skewed (quad q at q + (q >> 3)) so that both this pattern and the
//     coalesced one are free of bank conflicts; y goes back into the same tile;
//   * the comb z[n] = y[n] + g z[n - D] runs as a second pass over the tile in the COALESCED mapping, in chunks of 32 E
//     samples (E = 4, 2 or 1 samples per thread; chunk <= D, so a chunk never reads its own outputs).  All taps come from
//     one window W[r] = z[tb - D - sh + r], r < tile + 4 (sh = (tb - D) & 3 makes the window 16-byte aligned): the part
//     before the tile (r < D + sh) is copied from the output ring (L2 hits), and every chunk stores its z into the
//     window at r + D + sh as well (an aligned quad, since D + sh = tb = 0 mod 4), where later chunks of the same tile
//     find it.  Short delays therefore cost one barrier per chunk, not a pass per D samples.  With E = 4 the chunk's z
//     goes straight from registers to the ring (512 contiguous bytes).  Same two roundings per sample as
//     fbdelay_kernel: given the same y the bits are identical;
//   * the biquad carry {x[n-1], x[n-2], y[n-1], y[n-2]} at the end of the block is kept in a per-lane state array (the
//     y ring that directform_kernel re-reads it from does not exist here).
template <int S>
struct ChainGeom {                                 // tile geometry for S samples per thread
    static constexpr int TILE = 32 * S;            // samples per tile
    static constexpr int TQ = TILE / 4;            // quads per tile
    static constexpr int QUADS = TQ + TQ / 8;      // ... skewed: quad q at position q + (q >> 3)
    static constexpr int WQ = TQ + 2;              // tap window: r < tile + 4 used, one more quad is read (and ignored)
    static constexpr int TPQ = S / 4;              // quads per thread
};
constexpr int CH_MIN_DELAY = 32;                   // shorter combs use the separate kernels
constexpr int kChainBulkDefault = 0;               // see dfcomb_kernel: which of the chain's tiles travel by bulk copy (UBLKCP)

__device__ __forceinline__ unsigned ch_qpos(unsigned q) { return q + (q >> 3); }
__device__ __forceinline__ unsigned ch_fpos(unsigned m) { return m + ((m >> 5) << 2); }   // sample m of a tile (float index)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {   // L2 only (.cg): coherent with this
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"                         // warp's earlier global stores
                 :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// ---- bulk copies (TMA engine, SASS UBLKCP) completing on an mbarrier (SASS SYNCS): the copy engine moves the tile, the
// LSU pipe — this kernel's limiter with per-thread cp.async (LDGSTS: 17 of 84 shared-memory wavefronts per tile) — does not.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\t"
                 "bra WAIT_LOOP;\n\tWAIT_DONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// what the generic proxy did to memory (y into the tile, z into the window and the ring) is ordered before what the
// async proxy does next (the bulk copies that overwrite the tile buffers / read the ring)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// `count` floats (a multiple of 4) from ring position `pos` (a multiple of 4) of a ring of `mask + 1` floats into shared
// memory: one bulk copy, two where the ring wraps (both parts stay multiples of 16 bytes)
__device__ __forceinline__ void bulk_ring_g2s(float* smem_dst, const float* ring, unsigned pos, unsigned mask, unsigned count, unsigned long long* bar) {
    pos &= mask;
    const unsigned first = min(count, mask + 1u - pos);
    bulk_g2s(smem_dst, ring + pos, first * 4u, bar);
    if (first < count) bulk_g2s(smem_dst + first, ring, (count - first) * 4u, bar);
}

template <int S>
struct __align__(16) ChainWarpSmem {               // what one warp (= one lane of the chain) keeps in shared memory
    float4 x[2][ChainGeom<S>::QUADS];              // x tiles in flight (double buffered); the current one becomes y
    float4 w[ChainGeom<S>::WQ];                    // tap window of the current tile
    float4 pw[5];                                  // A^(S * 2^k), k < 5 (read by broadcast)
    float4 c;                                      // carry between fast tiles {x[tb-1], x[tb-2], y[tb-1], y[tb-2]}
    float4 rec_h, rec_an;                          // EXC: the lane's one-partial oscillator record (osc_one.cuh)
    uint4 rec_ph;
    unsigned long long bar_x[2], bar_w;            // BULK: mbarriers of the two x tiles and of the tap window
    unsigned long long pad_;
};

// z of 4 samples: a[0..7] are two consecutive quads of the tap window, the taps are a[SH .. SH + 3]
template <int SH>
__device__ __forceinline__ float4 comb_quad(const float4 y, const float4 a0, const float4 a1, float g) {
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    return make_float4(__fadd_rn(y.x, __fmul_rn(g, a[SH])), __fadd_rn(y.y, __fmul_rn(g, a[SH + 1])),
                       __fadd_rn(y.z, __fmul_rn(g, a[SH + 2])), __fadd_rn(y.w, __fmul_rn(g, a[SH + 3])));
}

// Combs shorter than 128 samples: chunks of 32 E samples (<= D), E consecutive samples per thread.  y is read ahead of
// the chunk loop, z goes straight to the ring (32 E contiguous floats per instruction) and into the tap window for the
// chunks that follow; one barrier per chunk.
template <int CH_TILE, int E, bool SKEW>
__device__ __forceinline__ void comb_small(const float* tf, float* wf, float* zdata, unsigned zb, unsigned zm, unsigned Du,
                                           unsigned sh, float g, unsigned wl) {
    constexpr int NC = CH_TILE / (32 * E);
    float y[NC][E];
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
        for (int e = 0; e < E; e++) { const unsigned m = 32u * E * c + E * wl + e; y[c][e] = tf[SKEW ? ch_fpos(m) : m]; }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const unsigned m = 32u * E * c + E * wl;
        float z[E];
#pragma unroll
        for (int e = 0; e < E; e++) z[e] = __fadd_rn(y[c][e], __fmul_rn(g, wf[m + sh + e]));
        const unsigned r = m + Du + sh;                             // a multiple of E: D + sh = 0 (mod 4)
        if (E == 2) {
            *reinterpret_cast<float2*>(zdata + ((zb + m) & zm)) = make_float2(z[0], z[E - 1]);
            if (c + 1 < NC && r < (unsigned)CH_TILE + 4u) *reinterpret_cast<float2*>(wf + r) = make_float2(z[0], z[E - 1]);
        } else {
            zdata[(zb + m) & zm] = z[0];
            if (c + 1 < NC && r < (unsigned)CH_TILE + 4u) wf[r] = z[0];
        }
        if (c + 1 < NC) __syncwarp();
    }
}

// EXC: the biquad's input is not a ring but a one-partial oscillator voice (exc_voice[lane] of the bank `exc`), evaluated
// in registers by osc_one_group8 — the function osc_one_kernel fills that voice's ring with when the chain is not fused,
// so the bits are the same — and the chain's only HBM traffic is its output: 4 B per lane-sample.
// BULK: the x tile and the tap window travel by bulk copy (one elected lane, an mbarrier per buffer) instead of per-thread
// cp.async.  A bulk copy lands linearly, so the ring-fed instance keeps its x tile unskewed (own-sample reads then take
// a 2-way bank conflict: +4 wavefronts per tile against the 17 the LDGSTS copies cost).
// BULK = 0: cp.async for both; 1: the x tile by bulk copy, the tap window by cp.async; 2: both by bulk copy.  The window
// is z this very warp stored to the ring a moment ago: reading it through the async proxy needs the full
// fence.proxy.async (SASS: MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC) every tile, where the x tile (written by an earlier kernel)
// needs only the shared-memory flavour for the buffer it overwrites.
template <bool EXC, int BULK>
__global__ void __launch_bounds__(DF_CTA_THREADS, 7)
dfcomb_kernel(const float* __restrict__ coef, const float* __restrict__ pw, const uint32_t* __restrict__ delay,
              const float* __restrict__ gain, const BufferDesc* __restrict__ bufdesc, const uint32_t* __restrict__ in_bufs,
              uint32_t first_out_buf, float4* __restrict__ state, unsigned n_lanes, unsigned long long lo, unsigned long long hi,
              OscOneSrc exc, const uint32_t* __restrict__ exc_voice, const uint32_t* __restrict__ order) {
    // Samples per thread.  16 halves the per-sample cost of the scan and of everything that happens once per tile, which
    // pays when the kernel is bound by issue slots (EXC: 3.20 -> 2.81 ms on cfg3) and not when it is bound by the LSU
    // pipe (ring input: 3.26 -> 3.43 ms); the biquad's tiles — hence its roundings — therefore differ between the two.
    constexpr int SPT = EXC ? 16 : DF_PER_THREAD;
    constexpr int PSHIFT = EXC ? 1 : 0;                            // stored powers are A^(8 * 2^k): A^(SPT * 2^k) = entry k + PSHIFT
    constexpr int CH_TILE = ChainGeom<SPT>::TILE, CH_TQ = ChainGeom<SPT>::TQ, CH_QUADS = ChainGeom<SPT>::QUADS, CH_TPQ = ChainGeom<SPT>::TPQ;
    constexpr int NW = DF_CTA_THREADS / 32;
    constexpr bool BX = BULK >= 1 && !EXC;                         // x tile by bulk copy
    constexpr bool BW = BULK == 2;                                 // tap window by bulk copy
    constexpr bool SKEW = !BX;                                     // layout of the fast tiles' x / y buffer
    auto XQ = [](unsigned q) { return SKEW ? ch_qpos(q) : q; };
    __shared__ ChainWarpSmem<SPT> s_all[NW];
    const unsigned wl = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const unsigned slot = blockIdx.x * NW + wi;
    if (slot >= n_lanes) return;                                   // whole warp exits together
    const unsigned lane = order[slot];                             // chain_lane_order: cost classes dealt out across the CTAs
    ChainWarpSmem<SPT>& S = s_all[wi];
    BufferDesc xin = {nullptr, 0};
    unsigned exc_flags = 0;
    if (EXC) {
        const OscOneVoice o = osc_one_load(exc, exc_voice[lane]);
        if (wl == 0) { S.rec_h = o.h; S.rec_an = o.an; S.rec_ph = o.ph; }
        exc_flags = o.flags;
    } else {
        xin = bufdesc[in_bufs[lane]];
    }
    // EXC: the 8 samples from n8 (a multiple of 8) of this lane's exciter; the record is read by broadcast
    auto exc8 = [&](unsigned long long n8, float (&r)[8]) {
        osc_one_group8(S.rec_h, S.rec_an, S.rec_ph, exc_flags, n8, exc.max_attack, r);
    };
    const BufferDesc zout = bufdesc[first_out_buf + lane];
    const float b0 = coef[lane * 5 + 0], b1 = coef[lane * 5 + 1], b2 = coef[lane * 5 + 2];
    const float a1 = coef[lane * 5 + 3], a2 = coef[lane * 5 + 4];
    const unsigned Du = delay[lane];
    const float g = gain[lane];
    if (wl < 5) S.pw[wl] = reinterpret_cast<const float4*>(pw + (size_t)lane * DF_LEVELS * 4)[wl + PSHIFT];
    __syncwarp();
    const float4* P = S.pw;                                         // warp-uniform: read by broadcast
    auto Pk = [&](int k) { return P[k]; };

    float x1c = 0.f, x2c = 0.f, y1c = 0.f, y2c = 0.f;               // every signal is 0 before t = 0
    if (lo > 0) { const float4 st = state[lane]; x1c = st.x; x2c = st.y; y1c = st.z; y2c = st.w; }

    const unsigned sh = (unsigned)(((long long)lo - (long long)Du) & 3);   // (tb - D) & 3 of every tile: tb = lo (mod 4)
    // Any tile: ragged ends, an unaligned block start, the first D + 4 samples after t = 0, rings beyond 2^32 floats.
    // Same biquad_tile; the comb walks the tile in chunks of 32 samples (<= D) with its taps straight from the ring.
    auto slow_tile = [&](unsigned long long tb) {
        const unsigned long long t0 = tb + (unsigned long long)wl * SPT;
        float* tf = reinterpret_cast<float*>(S.x[0]);
        {
            float x[SPT + 2], yv[SPT], y1, y2;
            if (EXC) {
                // the (at most two) groups of 8 this thread's samples fall into, through a private strip of the tile buffers
                constexpr int NG = SPT / 8;               // whole groups per thread; one more when unaligned
                static_assert(SPT % 8 == 0 && 2 * CH_QUADS * 4 >= 32 * 8 * (NG + 1), "exciter scratch");
                float* sc = reinterpret_cast<float*>(S.x[0]) + 8u * (NG + 1) * wl;
                const unsigned off = (unsigned)(t0 & 7ull);
#pragma unroll
                for (int gi = 0; gi <= NG; gi++) {
                    if (gi < NG || off) {
                        float r[8];
                        exc8(t0 - off + 8ull * gi, r);
#pragma unroll
                        for (int u = 0; u < 8; u++) sc[8 * gi + u] = r[u];
                    }
                }
#pragma unroll
                for (int j = 0; j < SPT; j++) x[2 + j] = (t0 + j < hi) ? sc[off + j] : 0.0f;
                __syncwarp();                                       // the strips overlap the tile y is about to go into
            } else {
#pragma unroll
                for (int j = 0; j < SPT; j++) x[2 + j] = (t0 + j < hi) ? xin.data[(t0 + j) & xin.mask] : 0.0f;
            }
            const float px1 = __shfl_up_sync(0xffffffffu, x[SPT + 1], 1), px2 = __shfl_up_sync(0xffffffffu, x[SPT], 1);
            x[1] = wl ? px1 : x1c;
            x[0] = wl ? px2 : x2c;
            biquad_tile<SPT>(x, y1c, y2c, Pk, b0, b1, b2, a1, a2, wl, yv, y1, y2);
            if (tb + CH_TILE >= hi) {                               // the block's last two samples: carry of the next launch
#pragma unroll
                for (int j = 0; j < SPT; j++) {
                    if (t0 + j + 1 == hi) { state[lane].x = x[2 + j]; state[lane].z = yv[j]; }
                    if (t0 + j + 2 == hi) { state[lane].y = x[2 + j]; state[lane].w = yv[j]; }
                }
                if (wl == 0 && tb + 1 == hi) { state[lane].y = x1c; state[lane].w = y1c; }   // hi - 2 lies before the tile
            }
#pragma unroll
            for (int j = 0; j < SPT; j++) tf[ch_fpos(wl * SPT + j)] = yv[j];
            y1c = __shfl_sync(0xffffffffu, y1, 31); y2c = __shfl_sync(0xffffffffu, y2, 31);
            x1c = __shfl_sync(0xffffffffu, x[SPT + 1], 31); x2c = __shfl_sync(0xffffffffu, x[SPT], 31);
        }
        __syncwarp();
        for (unsigned m = wl; m < (unsigned)CH_TILE; m += 32) {
            const unsigned long long t = tb + m;
            if (t < hi) {
                float tap = 0.0f;
                if (m >= Du) tap = tf[ch_fpos(m - Du)];
                else if (t >= Du) tap = __ldcg(zout.data + ((t - Du) & zout.mask));
                const float z = __fadd_rn(tf[ch_fpos(m)], __fmul_rn(g, tap));
                tf[ch_fpos(m)] = z;
                zout.data[t & zout.mask] = z;
            }
            __syncwarp();                                           // the next chunk (and the next tile's taps) read these
        }
    };

    const bool can_fast = (((xin.mask | zout.mask) >> 32) == 0) && (lo % (EXC ? 8 : 4) == 0);   // EXC: a thread = one group of 8
    unsigned long long tb = lo;
    while (tb < hi && !(can_fast && tb >= (unsigned long long)Du + 4 && tb + CH_TILE <= hi)) { slow_tile(tb); tb += CH_TILE; }
    if (tb < hi) {
        // Fast tiles: full, 16-byte aligned, every tap at a time >= 0, 32-bit ring indices.
        const unsigned long long n_fast = (hi - tb) / CH_TILE;
        const unsigned xm = (unsigned)xin.mask, zm = (unsigned)zout.mask;
        unsigned xb = (unsigned)(tb & xin.mask);                                // ring index of the tile's first sample
        unsigned zb = (unsigned)(tb & zout.mask);
        unsigned qb = (unsigned)((tb - Du - sh) & zout.mask);                   // ... of the window's first quad
        const unsigned dq = (Du + sh) >> 2;                                     // window quad of the tile's first sample
        float* wf = reinterpret_cast<float*>(S.w);
        if (wl == 0) S.c = make_float4(x1c, x2c, y1c, y2c);
        if (BX || BW) {
            if (wl == 0) {
                mbar_init(&S.bar_x[0], 1); mbar_init(&S.bar_x[1], 1); mbar_init(&S.bar_w, 1);
                asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
                if (BW) fence_proxy_async();                        // the slow tiles above wrote these buffers and the ring
                else fence_proxy_async_smem();
                if (BX) {
                    mbar_expect_tx(&S.bar_x[0], CH_TILE * 4u);
                    bulk_ring_g2s(reinterpret_cast<float*>(S.x[0]), xin.data, xb, xm, CH_TILE, &S.bar_x[0]);
                }
            }
            __syncwarp();
        }
        if (!BX && !EXC) {
#pragma unroll
            for (int c = 0; c < CH_TQ / 32; c++)
                cp_async16(&S.x[0][ch_qpos(32u * c + wl)], xin.data + ((xb + 128u * c + 4u * wl) & xm));
        }
        if (!BX) cp_async_commit();                                 // group "x of tile 0" (empty with EXC)
        for (unsigned long long k = 0; k < n_fast; k++) {
            const unsigned cur = (unsigned)k & 1u;
            float4* xs = S.x[cur];
            // ---- taps of tile k
            if (!BW) {
#pragma unroll
                for (int c = 0; c < CH_TQ / 32; c++)
                    cp_async16(&S.w[32u * c + wl], zout.data + ((qb + 128u * c + 4u * wl) & zm));
                if (wl == 0) cp_async16(&S.w[CH_TQ], zout.data + ((qb + (unsigned)CH_TILE) & zm));
                cp_async_commit();                                  // group "taps of tile k"
            }
            // ---- x of tile k + 1
            xb = (xb + CH_TILE) & xm;
            if (BX || BW) {
                if (wl == 0) {
                    if (BW) fence_proxy_async();                    // tile k - 1: y / z written by the generic proxy, z stored to the ring
                    else fence_proxy_async_smem();                  // tile k - 1: y written into the buffer the copy below overwrites
                    if (BW) {
                        mbar_expect_tx(&S.bar_w, (CH_TILE + 4) * 4u);
                        bulk_ring_g2s(wf, zout.data, qb, zm, CH_TILE + 4, &S.bar_w);
                    }
                    if (BX && k + 1 < n_fast) {
                        mbar_expect_tx(&S.bar_x[cur ^ 1], CH_TILE * 4u);
                        bulk_ring_g2s(reinterpret_cast<float*>(S.x[cur ^ 1]), xin.data, xb, xm, CH_TILE, &S.bar_x[cur ^ 1]);
                    }
                }
            }
            if (!BX) {
                if (!EXC && k + 1 < n_fast) {
#pragma unroll
                    for (int c = 0; c < CH_TQ / 32; c++)
                        cp_async16(&S.x[cur ^ 1][ch_qpos(32u * c + wl)], xin.data + ((xb + 128u * c + 4u * wl) & xm));
                }
                cp_async_commit();                                  // group "x of tile k + 1" (possibly empty)
            }
            // ---- x of tile k has landed
            if (BX) mbar_wait(&S.bar_x[cur], (unsigned)(k >> 1) & 1u);
            else if (!EXC) {
                if (BW) cp_async_wait<1>(); else cp_async_wait<2>();   // groups still allowed in flight: x of k + 1 (and the taps of k)
                __syncwarp();
            }
            {
                float x[SPT + 2], yv[SPT], y1, y2;
                const float4 cr = S.c;                              // broadcast read
                if (EXC) {
#pragma unroll
                    for (int gi = 0; gi < SPT / 8; gi++) {
                        float r[8];
                        exc8(tb + (unsigned long long)SPT * wl + 8ull * gi, r);
#pragma unroll
                        for (int j = 0; j < 8; j++) x[2 + 8 * gi + j] = r[j];
                    }
                    const float px1 = __shfl_up_sync(0xffffffffu, x[SPT + 1], 1), px2 = __shfl_up_sync(0xffffffffu, x[SPT], 1);
                    x[0] = wl ? px2 : cr.y;
                    x[1] = wl ? px1 : cr.x;
                } else {
#pragma unroll
                    for (int jq = 0; jq < CH_TPQ; jq++) {
                        const float4 v = xs[XQ(CH_TPQ * wl + jq)];
                        x[2 + 4 * jq] = v.x; x[3 + 4 * jq] = v.y; x[4 + 4 * jq] = v.z; x[5 + 4 * jq] = v.w;
                    }
                    // history (x[t0-2], x[t0-1]) = the quad before this thread's own; lane 0 reads the carry instead
                    const float2 vh = reinterpret_cast<const float2*>(xs + XQ(wl ? CH_TPQ * wl - 1u : 0u))[1];
                    x[0] = wl ? vh.x : cr.y;
                    x[1] = wl ? vh.y : cr.x;
                }
                biquad_tile<SPT>(x, cr.z, cr.w, Pk, b0, b1, b2, a1, a2, wl, yv, y1, y2);
                __syncwarp();                                       // everyone has read its x (and the carry): y may overwrite them
                if (wl == 31) {
                    const float4 c = make_float4(x[SPT + 1], x[SPT], y1, y2);
                    S.c = c;
                    if (k + 1 == n_fast && tb + CH_TILE == hi) state[lane] = c;
                }
#pragma unroll
                for (int jq = 0; jq < CH_TPQ; jq++)
                    xs[XQ(CH_TPQ * wl + jq)] = make_float4(yv[4 * jq], yv[4 * jq + 1], yv[4 * jq + 2], yv[4 * jq + 3]);
            }
            if (BW) mbar_wait(&S.bar_w, (unsigned)k & 1u);          // the taps have landed
            else if (BX) cp_async_wait<0>();                        // (the only cp.async group of the tile)
            else cp_async_wait<1>();
            __syncwarp();                                           // ... and y is where the coalesced mapping finds it
            if (Du >= 128u) {                                       // warp-uniform
                const bool inside = Du < (unsigned)CH_TILE;         // some taps are outputs of this very tile
#pragma unroll
                for (int c = 0; c < CH_TQ / 32; c++) {
                    const unsigned q = 32u * c + wl;
                    const float4 t0 = S.w[q], y = xs[XQ(q)];
                    float4 t1 = t0;
                    if (sh) t1 = S.w[q + 1];
                    float4 z;
                    switch (sh) {                                   // warp-uniform
                        case 0: z = comb_quad<0>(y, t0, t1, g); break;
                        case 1: z = comb_quad<1>(y, t0, t1, g); break;
                        case 2: z = comb_quad<2>(y, t0, t1, g); break;
                        default: z = comb_quad<3>(y, t0, t1, g); break;
                    }
                    *reinterpret_cast<float4*>(zout.data + ((zb + 4u * q) & zm)) = z;
                    if (inside && c + 1 < CH_TQ / 32) {
                        if (q + dq <= (unsigned)CH_TQ) S.w[q + dq] = z;
                        __syncwarp();
                    }
                }
            } else if (Du >= 64u) {
                comb_small<CH_TILE, 2, SKEW>(reinterpret_cast<const float*>(xs), wf, zout.data, zb, zm, Du, sh, g, wl);
            } else {
                comb_small<CH_TILE, 1, SKEW>(reinterpret_cast<const float*>(xs), wf, zout.data, zb, zm, Du, sh, g, wl);
            }
            __syncwarp();                                           // the next tile's taps read these stores
            qb += CH_TILE; zb += CH_TILE;
            tb += CH_TILE;
        }
        { const float4 c = S.c; x1c = c.x; x2c = c.y; y1c = c.z; y2c = c.w; }   // for the ragged tail, if any
        cp_async_wait<0>();
        __syncwarp();
        while (tb < hi) { slow_tile(tb); tb += CH_TILE; }
    }
}

bool chain_fusable(const DirectFormDev& df, const FbDelayDev& fb) {
    return df.n_lanes == fb.n_lanes && df.n_lanes > 0 && fb.min_delay >= (uint64_t)CH_MIN_DELAY;
}

std::shared_ptr<ChainStateDev> chain_state_create(uint32_t n_lanes) {
    auto s = std::make_shared<ChainStateDev>();
    s->n_lanes = n_lanes;
    if (cudaMalloc(&s->d_state, std::max<uint32_t>(n_lanes, 1) * sizeof(float4)) != cudaSuccess) return nullptr;
    return s;
}

cudaError_t launch_dfcomb(const DirectFormDev& df, const FbDelayDev& fb, ChainStateDev& st, const BufferDesc* d_bufdesc,
                          const uint32_t* d_in_bufs, uint32_t first_out_buf, uint64_t lo, uint64_t hi, cudaStream_t stream,
                          uint64_t* n_launches, const OscOneSrc* exciter, const uint32_t* d_exc_voice) {
    if (n_launches) *n_launches = 0;
    if (hi <= lo || df.n_lanes == 0) return cudaSuccess;
    if (lo != 0 && lo != st.time) return cudaErrorInvalidValue;    // the renderer restarts recurrences at t = 0 or continues
    const unsigned per_cta = DF_CTA_THREADS / 32;
    const unsigned grid = (df.n_lanes + per_cta - 1) / per_cta;
    // measurement knob: FRB_K4_BULK=0 / 1 selects the cp.async (LDGSTS) / bulk-copy (UBLKCP + mbarrier) instance
    static const int bulk_env = [] { const char* e = getenv("FRB_K4_BULK"); return e ? atoi(e) : -1; }();
    const int bulk = bulk_env < 0 ? kChainBulkDefault : bulk_env;
    auto go = [&](auto kernel, const uint32_t* in_bufs, const OscOneSrc& ex, const uint32_t* ev) {
        kernel<<<grid, DF_CTA_THREADS, 0, stream>>>(df.d_coef, df.d_pow, fb.d_delay, fb.d_gain, d_bufdesc, in_bufs, first_out_buf,
                                                    st.d_state, df.n_lanes, lo, hi, ex, ev, fb.d_order);
    };
    if (exciter) {
        if (bulk == 2) go(dfcomb_kernel<true, 2>, nullptr, *exciter, d_exc_voice);
        else go(dfcomb_kernel<true, 0>, nullptr, *exciter, d_exc_voice);        // no x tile to copy: 1 == 0
    } else {
        if (bulk == 2) go(dfcomb_kernel<false, 2>, d_in_bufs, OscOneSrc{}, nullptr);
        else if (bulk == 1) go(dfcomb_kernel<false, 1>, d_in_bufs, OscOneSrc{}, nullptr);
        else go(dfcomb_kernel<false, 0>, d_in_bufs, OscOneSrc{}, nullptr);
    }
    st.time = hi;
    if (n_launches) *n_launches = 1;
    return cudaGetLastError();
}

cudaError_t launch_directform(const DirectFormDev& f, const BufferDesc* d_bufdesc, const uint32_t* d_in_bufs,
                              uint32_t first_out_buf, uint64_t lo, uint64_t hi, int, cudaStream_t stream, uint64_t* n_launches) {
    if (n_launches) *n_launches = 0;
    if (hi <= lo || f.n_lanes == 0) return cudaSuccess;
    const unsigned per_cta = DF_CTA_THREADS / 32;
    directform_kernel<<<(f.n_lanes + per_cta - 1) / per_cta, DF_CTA_THREADS, 0, stream>>>(f.d_coef, f.d_pow, d_bufdesc, d_in_bufs, first_out_buf, f.n_lanes, lo, hi);
    if (n_launches) *n_launches = 1;
    return cudaGetLastError();
}

cudaError_t launch_fbdelay(const FbDelayDev& f, const BufferDesc* d_bufdesc, const uint32_t* d_in_bufs,
                           uint32_t first_out_buf, uint64_t lo, uint64_t hi, int, cudaStream_t stream, uint64_t* n_launches) {
    if (n_launches) *n_launches = 0;
    if (hi <= lo || f.n_lanes == 0) return cudaSuccess;
    const unsigned bx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((f.max_delay + FB_THREADS - 1) / FB_THREADS, 64));
    fbdelay_kernel<<<dim3(bx, f.n_lanes), FB_THREADS, 0, stream>>>(f.d_delay, f.d_gain, d_bufdesc, d_in_bufs, first_out_buf, lo, hi);
    if (n_launches) *n_launches = 1;
    return cudaGetLastError();
}

}  // namespace frb
