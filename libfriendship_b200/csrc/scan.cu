// scan.cu — K4: Direct Form linear recurrences for sm_100a (extension nodes FRB_KIND_DIRECTFORM / FRB_KIND_FBDELAY;
// not in the reference: SURVEY.md F2 — the reference rejects cycles, so feedback only exists as these nodes).
//
// Bound: HBM.  Algorithmic traffic = 8 B per lane-sample for each node (read x, write y): each kernel reads its
// input ring once and writes its output ring once; no intermediate touches HBM.
//
// DirectForm (biquad, Direct Form I), per lane:  y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2].
//   One CTA owns one lane and walks the block in tiles of 256 threads x 8 samples.  Inside a tile the recurrence is
//   evaluated as a parallel prefix scan over the recurrence's transfer matrices: every thread runs its 8 samples
//   from a zero state, the 2-vectors of end states are combined with a Kogge-Stone scan whose operator is
//   "multiply by A^(8*2^k) and add" (A = [[-a1, -a2], [1, 0]]; the powers come from an fp64 setup at definition
//   time), then every thread re-runs its 8 samples from its true initial state.  The carry between tiles and
//   between launches is just the last two samples of x and y, re-read from the rings, so consecutive blocks
//   continue exactly.
//
// FbDelay, per lane:  y[n] = x[n] + g y[n-D].  D independent first-order recurrences with stride D: one CTA per
//   lane steps through time D samples at a time, all D phases in parallel; every output is computed in the same
//   order and with the same two roundings (mul, add) as a sequential f32 evaluation, so it is bit-exact to one.
#include "scan.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace frb {

constexpr int FB_THREADS = 128;
constexpr int DF_CTA_THREADS = 128;   // 4 lanes per CTA, one warp each
constexpr int DF_PER_THREAD = 8;
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
    uint32_t* d_delay = nullptr;
    float* d_gain = nullptr;
    ~FbDelayDev() { cudaFree(d_delay); cudaFree(d_gain); }
};

uint32_t directform_lanes(const DirectFormDev& f) { return f.n_lanes; }
uint32_t fbdelay_lanes(const FbDelayDev& f) { return f.n_lanes; }
uint64_t fbdelay_max_delay(const FbDelayDev& f) { return f.max_delay; }

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

std::shared_ptr<FbDelayDev> fbdelay_create(const frb_fbdelay_desc* d, cudaStream_t stream, std::string* err) {
    auto fail = [&](const std::string& m) { if (err) *err = m; return std::shared_ptr<FbDelayDev>(); };
    if (d->n_lanes && (!d->delay || !d->gain)) return fail("fbdelay: null array");
    auto f = std::make_shared<FbDelayDev>();
    f->n_lanes = d->n_lanes;
    for (uint32_t i = 0; i < d->n_lanes; i++) {
        if (d->delay[i] < 1) return fail("fbdelay: delay must be >= 1");
        f->max_delay = std::max<uint64_t>(f->max_delay, d->delay[i]);
    }
    const size_t n = std::max<uint32_t>(d->n_lanes, 1);
    if (cudaMalloc(&f->d_delay, n * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&f->d_gain, n * sizeof(float)) != cudaSuccess)
        return fail("fbdelay: out of device memory");
    if (d->n_lanes) {
        cudaMemcpyAsync(f->d_delay, d->delay, d->n_lanes * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(f->d_gain, d->gain, d->n_lanes * sizeof(float), cudaMemcpyHostToDevice, stream);
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail("fbdelay: upload failed");
    return f;
}

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ring_at(const BufferDesc& b, long long t) {
    return t >= 0 ? b.data[(unsigned long long)t & b.mask] : 0.0f;    // every signal is 0 before t = 0
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
    // A^(8*2^k), k = 0..5, in registers (warp-uniform)
    float P[6][4];
#pragma unroll
    for (int k = 0; k < 6; k++)
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
            const float4 v0 = *reinterpret_cast<const float4*>(xin.data + (t0 & xin.mask));
            const float4 v1 = *reinterpret_cast<const float4*>(xin.data + ((t0 + 4) & xin.mask));
            x[2] = v0.x; x[3] = v0.y; x[4] = v0.z; x[5] = v0.w; x[6] = v1.x; x[7] = v1.y; x[8] = v1.z; x[9] = v1.w;
        } else {
#pragma unroll
            for (int j = 0; j < DF_PER_THREAD; j++) x[2 + j] = (t0 + j < hi) ? xin.data[(t0 + j) & xin.mask] : 0.0f;
        }
        // two samples of history from the previous thread (lane 0: from the carry)
        const float px1 = __shfl_up_sync(0xffffffffu, x[9], 1), px2 = __shfl_up_sync(0xffffffffu, x[8], 1);
        x[1] = wl ? px1 : x1c;
        x[0] = wl ? px2 : x2c;
        float u[DF_PER_THREAD];
#pragma unroll
        for (int j = 0; j < DF_PER_THREAD; j++) u[j] = fmaf(b2, x[j], fmaf(b1, x[j + 1], b0 * x[j + 2]));
        // zero-state run: end state e = (y[7], y[6])
        float e1 = 0.f, e2 = 0.f;
#pragma unroll
        for (int j = 0; j < DF_PER_THREAD; j++) {
            const float y = fmaf(-a2, e2, fmaf(-a1, e1, u[j]));
            e2 = e1; e1 = y;
        }
        // inclusive scan over the warp: v_i <- v_i + A^(8*2^k) v_(i-2^k)
        float v1 = e1, v2 = e2;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float o1 = __shfl_up_sync(0xffffffffu, v1, 1u << k), o2 = __shfl_up_sync(0xffffffffu, v2, 1u << k);
            if (wl >= (1u << k)) {
                v1 += P[k][0] * o1 + P[k][1] * o2;
                v2 += P[k][2] * o1 + P[k][3] * o2;
            }
        }
        // state entering this thread = A^(8*wl) carry + exclusive prefix
        float p1 = __shfl_up_sync(0xffffffffu, v1, 1), p2 = __shfl_up_sync(0xffffffffu, v2, 1);
        if (wl == 0) { p1 = 0.f; p2 = 0.f; }
        float h1 = y1c, h2 = y2c;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            if (wl & (1u << k)) {
                const float n1 = P[k][0] * h1 + P[k][1] * h2;
                const float n2 = P[k][2] * h1 + P[k][3] * h2;
                h1 = n1; h2 = n2;
            }
        }
        float y1 = h1 + p1, y2 = h2 + p2;                           // (y[t0-1], y[t0-2])
        float yv[DF_PER_THREAD];
#pragma unroll
        for (int j = 0; j < DF_PER_THREAD; j++) {
            const float y = fmaf(-a2, y2, fmaf(-a1, y1, u[j]));
            yv[j] = y; y2 = y1; y1 = y;
        }
        if (vec && t0 + DF_PER_THREAD <= hi) {
            *reinterpret_cast<float4*>(yout.data + (t0 & yout.mask)) = make_float4(yv[0], yv[1], yv[2], yv[3]);
            *reinterpret_cast<float4*>(yout.data + ((t0 + 4) & yout.mask)) = make_float4(yv[4], yv[5], yv[6], yv[7]);
        } else {
#pragma unroll
            for (int j = 0; j < DF_PER_THREAD; j++)
                if (t0 + j < hi) yout.data[(t0 + j) & yout.mask] = yv[j];
        }
        // carries for the next tile come from the last thread
        y1c = __shfl_sync(0xffffffffu, y1, 31); y2c = __shfl_sync(0xffffffffu, y2, 31);
        x1c = __shfl_sync(0xffffffffu, x[9], 31); x2c = __shfl_sync(0xffffffffu, x[8], 31);
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
