// osc_tc.cuh — K1T: the matrix-product oscillator bank of osc_gemm.cuh on the 5th-generation tensor cores (included by osc.cu).
//
// Same mathematics, same operand values (anchors, rotations, fp16 hi/lo split, per-voice scale) as K1G; what changes is
// where the product runs: tcgen05.mma (SASS UTCHMMA), issued by one thread, accumulator in tensor memory, both operands
// read from SHARED memory — so the threads generate A and W into shared memory in the canonical K-major layout of a
// UMMA operand instead of into mma.sync register fragments, every element once per CTA (K1G generates each twice).
//   * CTA = 256 threads = one tile of 128 blocks x 128 samples of one voice; accumulator 128 lanes x 128 columns of TMEM.
//   * stage = 16 partials (one record group) = two k16 steps x {A hi, A lo, W hi, W lo} x 4 KB, double buffered.
//     Per stage: all warps generate (warp = a quad of partials x a 64-row half, one anchor + 8 rotations per thread and
//     operand; a store instruction covers 128 contiguous bytes), fence.proxy.async, __syncthreads, thread 0 issues the six
//     MMAs (lo*hi, hi*lo, hi*hi per k16 step) and commits them to the buffer's mbarrier; the MMAs run while the next
//     stage is generated, and a buffer is regenerated only after its mbarrier has flipped.
//   * operand tile (128 rows x 16 fp16, K-major, no swizzle): element (r, k) at (k / 8) * 2048 + r * 16 + (k % 8) * 2 bytes:
//     core matrices of 8 rows x 16 bytes, 128 bytes apart along M/N (SBO), 2,048 bytes apart along K (LBO).
//   * accumulation: the tensor core adds into its fp32 accumulator with truncation — measured: 1,536 MMAs in a row into one
//     accumulator leave 1.2e-5 of full scale, all of one sign — so the MMAs of GT_FLUSH stages (24 MMAs: at most 1.4e-6)
//     form a chunk, chunks alternate between two TMEM accumulators, and while chunk c + 1 is being issued the threads
//     drain chunk c with tcgen05.ld into fp32 sums in REGISTERS (thread = one row x 64 columns; FADD, round to nearest).
//     The drain sits between a stage's generation and its barrier, a whole stage after the chunk's last commit: no wait.
//   * write-out: every thread stores its row's 64 samples from those registers.
// Bound: generating the operands (about 250 instructions per thread and stage against 384 tensor-pipe cycles) and the
// shared-memory bandwidth they and the MMAs' operand reads share; DESIGN.md.
#pragma once

#include "osc_gemm.cuh"

namespace frb {

constexpr int GT_THREADS = 256;
constexpr int GT_TILE_BYTES = 4096;                       // one operand tile: 128 rows x 16 fp16
constexpr int GT_STAGE_BYTES = 2 * 4 * GT_TILE_BYTES;     // two k16 steps x {A hi, A lo, W hi, W lo}
constexpr int GT_TMEM_COLS = 256;                        // two accumulators of 128 columns, used in turn
constexpr unsigned GT_FLUSH = 4;                          // stages (6 MMAs each) between drains of the TMEM accumulator
constexpr size_t GT_SMEM = 2 * (size_t)GT_STAGE_BYTES + 1024;   // + slack to align the buffers to 1 KB

struct __align__(16) GtRecs { uint32_t w[GT_REC_WORDS][16]; };           // one record group, word-major: a warp's load of one word is ONE
                                                          // shared-memory wavefront (a 16-byte record load is four, and the loads
                                                          // contend with the tensor core's operand reads: ncu, 24% of all wavefronts)

__device__ __forceinline__ unsigned gt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = gt_smem_u32(bar);
    for (unsigned spin = 0; spin < (1u << 24); spin++) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();                                             // a lost arrival must not hang the device
}
// K-major, no swizzle: LBO (K direction) 2,048 B, SBO (M/N direction) 128 B, descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long gt_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3ffffu) >> 4) | ((unsigned long long)(2048u >> 4) << 16) |
           ((unsigned long long)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void gt_mma(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(GT_THREADS, 2) osc_tc_kernel(OscGemmLaunch p) {
    extern __shared__ unsigned char gt_raw[];
    __shared__ __align__(8) unsigned long long bar_done[2];
    __shared__ unsigned tmem_base_s;
    __shared__ GtRecs recs[2];
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned g = lane >> 2, tq = lane & 3;
    const unsigned hh = w & 3, half64 = w >> 2;           // this warp: partials tq + 4 hh of the group, rows / columns 64 half64 + g + 8 i
    const unsigned v = blockIdx.y;
    const unsigned long long tile = p.tile0 + blockIdx.x;
    unsigned char* const buf0 = gt_raw + ((1024u - (gt_smem_u32(gt_raw) & 1023u)) & 1023u);
    const unsigned buf0_u32 = gt_smem_u32(buf0);

    if (tid == 0) {
        for (int i = 0; i < 2; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(gt_smem_u32(&bar_done[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(gt_smem_u32(&tmem_base_s)), "r"(GT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    const unsigned n_stages = p.n_grp[v];                 // K == 16: one record group per stage
    auto stage_recs = [&](unsigned s, unsigned b) {       // 36 threads x 16 bytes: the records of stage s
        if (tid < GT_REC_WORDS * 4) {
            const uint32_t* src = p.tc + ((size_t)p.grp_begin[v] + s) * (GT_REC_WORDS * 16) + 4u * tid;
            gm_cp16(&recs[b].w[0][0] + 4u * tid, src);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (n_stages) stage_recs(0, 0);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_d = tmem_base_s;
    const float2 vs = p.vscale[v];
    // instruction descriptor: D f32 (bits 4-5 = 1), A and B f16 (0), both K-major (0), N >> 3 at bit 17, M >> 4 at bit 24
    const unsigned idesc = (1u << 4) | ((unsigned)(GM_N >> 3) << 17) | ((unsigned)(GM_M >> 4) << 24);
    const unsigned long long nA = (tile * GM_M + 64u * half64 + g) * (unsigned long long)GM_N;   // sample index of this thread's first row
    const unsigned cW = 64u * half64 + g;                                                        // ... and its first column
    // byte offset of this thread's element (row / column r0 + 8 i) inside an operand tile: k = 2 tq (+1) of k-half hh & 1
    const unsigned el_off = (hh & 1u) * 2048u + (64u * half64 + g) * 16u + tq * 4u;
    const unsigned step_off = (hh >> 1) * 4u * GT_TILE_BYTES;

    float sums[64];
#pragma unroll
    for (int k = 0; k < 64; k++) sums[k] = 0.f;

    auto drain = [&](unsigned which) {                                  // sums += accumulator `which` (its MMAs have completed)
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (unsigned c = 0; c < 4; c++) {
            uint32_t r[16];
            const unsigned taddr = tmem_d + ((32u * (w & 3u)) << 16) + which * 128u + 64u * (w >> 2) + 16u * c;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int k = 0; k < 16; k++) sums[16 * c + k] += __uint_as_float(r[k]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // ordered before the MMA that restarts this accumulator
    };

    for (unsigned s = 0; s < n_stages; s++) {
        const unsigned b = s & 1u;
        if (s + 1 < n_stages) stage_recs(s + 1, b ^ 1u);
        if (s >= 2) gt_wait(&bar_done[b], ((s >> 1) - 1u) & 1u);       // the MMAs that read this buffer two stages ago are done
        unsigned char* const base = buf0 + (size_t)b * GT_STAGE_BYTES + step_off + el_off;
        const unsigned P = tq + 4u * hh;
        const uint4 ph = make_uint4(recs[b].w[0][P], recs[b].w[1][P], 0u, recs[b].w[2][P]);
        const float4 an = make_float4(0.f, __uint_as_float(recs[b].w[4][P]), __uint_as_float(recs[b].w[3][P]), 0.f);
        const float4 rt = make_float4(__uint_as_float(recs[b].w[5][P]), __uint_as_float(recs[b].w[6][P]),
                                      __uint_as_float(recs[b].w[7][P]), __uint_as_float(recs[b].w[8][P]));
        {   // A: (sin, cos) at rows 64 half64 + g + 8 i
            float zs, zc;
            gm_anchor(ph.x, ph.y, ph.w, an.z, an.y * vs.x, nA, zs, zc);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t hi, lo;
                gm_split(zs, zc, hi, lo);
                *reinterpret_cast<uint32_t*>(base + i * 128) = hi;
                *reinterpret_cast<uint32_t*>(base + GT_TILE_BYTES + i * 128) = lo;
                const float ns = fmaf(zs, rt.x, zc * rt.y), nc = fmaf(zc, rt.x, -zs * rt.y);
                zs = ns; zc = nc;
            }
        }
        {   // W: (cos, sin) at columns 64 half64 + g + 8 i
            float ws, wc;
            gm_anchor(ph.x, ph.y, 0u, an.z, GM_WSCALE, cW, ws, wc);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t hi, lo;
                gm_split(wc, ws, hi, lo);
                *reinterpret_cast<uint32_t*>(base + 2 * GT_TILE_BYTES + i * 128) = hi;
                *reinterpret_cast<uint32_t*>(base + 3 * GT_TILE_BYTES + i * 128) = lo;
                const float nc = fmaf(wc, rt.z, -ws * rt.w), ns = fmaf(ws, rt.z, wc * rt.w);
                wc = nc; ws = ns;
            }
        }
        if (s && (s % GT_FLUSH) == 0) {                                 // the previous chunk: committed a stage ago
            gt_wait(&bar_done[(s - 1) & 1u], ((s - 1) >> 1) & 1u);
            drain((s / GT_FLUSH - 1u) & 1u);
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");          // the next stage's records (visible after the barrier)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy stores -> the MMAs' async-proxy reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const unsigned sb = buf0_u32 + b * GT_STAGE_BYTES;
#pragma unroll
            for (unsigned u = 0; u < 2; u++) {
                const unsigned t0 = sb + u * 4u * GT_TILE_BYTES;
                const unsigned long long a_hi = gt_desc(t0), a_lo = gt_desc(t0 + GT_TILE_BYTES);
                const unsigned long long w_hi = gt_desc(t0 + 2 * GT_TILE_BYTES), w_lo = gt_desc(t0 + 3 * GT_TILE_BYTES);
                const unsigned d = tmem_d + ((s / GT_FLUSH) & 1u) * 128u;                // this chunk's accumulator
                gt_mma(d, a_lo, w_hi, idesc, ((s % GT_FLUSH) | u) ? 1u : 0u);            // a chunk starts from zero
                gt_mma(d, a_hi, w_lo, idesc, 1u);
                gt_mma(d, a_hi, w_hi, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(gt_smem_u32(&bar_done[b])) : "memory");
        }
    }
    if (n_stages) {                                                      // the last chunk
        const unsigned last = n_stages - 1;
        gt_wait(&bar_done[last & 1u], (last >> 1) & 1u);
        drain((last / GT_FLUSH) & 1u);
    }
    // ---- write-out: this thread's row (block 32 (w & 3) + lane of the tile), columns 64 (w >> 2) ... + 63
    {
        const BufferDesc bd = p.bufdesc[p.first_buf + v];
        const unsigned long long t_row = (tile * GM_M + 32u * (w & 3u) + lane) * (unsigned long long)GM_N + 64u * (w >> 2);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const unsigned long long t = t_row + 4u * q;
            const float un = vs.y * (1.0f / GM_WSCALE);
            const float o[4] = {sums[4 * q] * un, sums[4 * q + 1] * un, sums[4 * q + 2] * un, sums[4 * q + 3] * un};
            if (t >= p.lo && t + 4 <= p.hi) {
                *reinterpret_cast<float4*>(bd.data + (t & bd.mask)) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (t + u >= p.lo && t + u < p.hi) bd.data[(t + u) & bd.mask] = o[u];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(GT_TMEM_COLS) : "memory");
}

}  // namespace frb
