// interp.cu — K2/K3: fused elementwise + Delay pass over a block of time (sm_100a).
//
// This is the INTERPRETER tier (the immediate path after a graph edit); a hot stage is compiled by the stage JIT
// (jit.cc) into a straight-line kernel built from the same device helpers (interp_device.inc).
// One launch evaluates one stage program (schedule.hpp) for every sample of a time block; blockIdx.y selects one
// of the stage's independent strands.  Each thread owns eight consecutive samples (absolute time aligned to 8, so
// undelayed plane reads/writes are 128-bit and coalesced); the program's registers live in shared memory as
// float4 columns (conflict-free), so intermediates of the fused nodes never touch HBM.  Delay is an indexed read
// at t - floor(d) from the external-input history or from a ring buffer in HBM written by an earlier stage.
//
// Arithmetic is the reference's, bit for bit (reference src/render/reference.rs:197-262): IEEE f32 with no
// FMA contraction (__fmul_rn/__fadd_rn/__fdiv_rn never contract), fmodf for `%`, fminf for f32::min.
#include "interp.cuh"

namespace frb {

// Each thread owns INTERP_VW float4 (= 8 consecutive samples): one decoded instruction is applied to both, which
// halves the interpretive overhead per byte moved (the kernel is instruction-issue bound, not HBM bound, otherwise).
__global__ void __launch_bounds__(INTERP_THREADS)
interp_kernel(InterpParams p) {
    extern __shared__ float4 s_regs[];
    const unsigned tid = threadIdx.x;
    const unsigned nthr = blockDim.x;
    constexpr int VW = INTERP_VW;
    // the program is staged in shared memory behind the register columns when it fits (broadcast LDS instead of
    // a dependent global load per interpreted instruction)
    const unsigned s_lo = __ldg(p.program + blockIdx.y), s_hi = __ldg(p.program + blockIdx.y + 1);   // this block's strand
    const uint4* prog = reinterpret_cast<const uint4*>(p.program + p.prog_base) + s_lo;
    if (p.prog_in_smem) {
        uint4* s_prog = reinterpret_cast<uint4*>(s_regs + (size_t)p.n_regs * VW * nthr);
        for (unsigned i = tid; i < s_hi - s_lo; i += nthr) s_prog[i] = __ldg(prog + i);
        __syncthreads();
        prog = s_prog;
    }

    for (unsigned long long g = (unsigned long long)blockIdx.x * nthr + tid; g < p.n_groups;
         g += (unsigned long long)gridDim.x * nthr) {
        const unsigned long long t0g = p.t_begin + (unsigned long long)(4 * VW) * g;   // absolute time of element 0; multiple of 8
#define REG(r, w) s_regs[((r) * VW + (w)) * nthr + tid]
        uint4 ins = prog[0];
        for (unsigned pc = 0;; pc++) {
            const uint4 nxt = prog[pc + 1];                       // prefetch (programs end with two I_ENDs)
            const unsigned op = ins.x & 0xFFu, flags = (ins.x >> 8) & 0xFFu, dst = ins.x >> 16;
            if (op == I_END) break;
            float4 a[VW], b[VW];
            if (op != I_LDIN && op != I_LDBUF && op != I_TAP_IN && op != I_TAP_BUF) {
                if (flags & IF_A_IMM) {
                    const float v = __uint_as_float(ins.y);
#pragma unroll
                    for (int w = 0; w < VW; w++) a[w] = make_float4(v, v, v, v);
                } else {
#pragma unroll
                    for (int w = 0; w < VW; w++) a[w] = REG(ins.y, w);
                }
            }
            if (op <= I_MIN || op == I_DLY_TI) {
                if (flags & IF_B_IMM) {
                    const float v = __uint_as_float(ins.z);
#pragma unroll
                    for (int w = 0; w < VW; w++) b[w] = make_float4(v, v, v, v);
                } else {
#pragma unroll
                    for (int w = 0; w < VW; w++) b[w] = REG(ins.z, w);
                }
            }
            switch (op) {
                case I_ADD:
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = make_float4(__fadd_rn(a[w].x, b[w].x), __fadd_rn(a[w].y, b[w].y), __fadd_rn(a[w].z, b[w].z), __fadd_rn(a[w].w, b[w].w));
                    break;
                case I_MUL:
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = make_float4(__fmul_rn(a[w].x, b[w].x), __fmul_rn(a[w].y, b[w].y), __fmul_rn(a[w].z, b[w].z), __fmul_rn(a[w].w, b[w].w));
                    break;
                case I_DIV:
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = make_float4(__fdiv_rn(a[w].x, b[w].x), __fdiv_rn(a[w].y, b[w].y), __fdiv_rn(a[w].z, b[w].z), __fdiv_rn(a[w].w, b[w].w));
                    break;
                case I_MOD:
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = make_float4(op_mod(a[w].x, b[w].x), op_mod(a[w].y, b[w].y), op_mod(a[w].z, b[w].z), op_mod(a[w].w, b[w].w));
                    break;
                case I_MIN:
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = f4min(a[w], b[w], p.sparkle_min);
                    break;
                case I_MOV:
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = a[w];
                    break;
                case I_LDIN: {
                    const InputDesc in = input_desc(p, ins.w);
#pragma unroll
                    for (int w = 0; w < VW; w++) {
                        const unsigned long long t = t0g + 4 * w;
                        float4 r;
                        if (t >= in.base_time && t + 4 <= in.end_time) {
                            r = *reinterpret_cast<const float4*>(in.data + (t - in.base_time));   // base_time % 4 == 0
                        } else {
                            r = make_float4(load_input(in, t), load_input(in, t + 1), load_input(in, t + 2), load_input(in, t + 3));
                        }
                        REG(dst, w) = r;
                    }
                    break;
                }
                case I_LDBUF: {
                    const BufferDesc bd = p.buffers[ins.w];
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = *reinterpret_cast<const float4*>(bd.data + ((t0g + 4 * w) & bd.mask));
                    break;
                }
                case I_TAP_IN: {
                    const InputDesc in = input_desc(p, ins.w);
                    const unsigned long long shift = ((unsigned long long)ins.z << 32) | ins.y;
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = f4tap_in(in, t0g + 4 * w, shift);
                    break;
                }
                case I_TAP_BUF: {
                    const BufferDesc bd = p.buffers[ins.w];
                    const unsigned long long shift = ((unsigned long long)ins.z << 32) | ins.y;
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = f4tap_buf(bd, t0g + 4 * w, shift);
                    break;
                }
                case I_GATE: {
                    const unsigned long long thr = ((unsigned long long)ins.w << 32) | ins.z;
#pragma unroll
                    for (int w = 0; w < VW; w++) REG(dst, w) = f4gate(a[w], t0g + 4 * w, thr);
                    break;
                }
                case I_STBUF: {
                    const BufferDesc bd = p.buffers[ins.w];
#pragma unroll
                    for (int w = 0; w < VW; w++) *reinterpret_cast<float4*>(bd.data + ((t0g + 4 * w) & bd.mask)) = a[w];
                    break;
                }
                case I_STOUT: {
                    // out[slot][t - t0], only for t0 <= t < t1
                    float* row = p.out + (unsigned long long)ins.w * p.out_stride;
#pragma unroll
                    for (int w = 0; w < VW; w++) {
                        const unsigned long long t = t0g + 4 * w;
                        if (t >= p.t0 && t + 4 <= p.t1 && p.out_vec_ok) {
                            *reinterpret_cast<float4*>(row + (t - p.t0)) = a[w];
                        } else {
                            const float v[4] = {a[w].x, a[w].y, a[w].z, a[w].w};
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                if (t + i >= p.t0 && t + i < p.t1) row[t + i - p.t0] = v[i];
                        }
                    }
                    break;
                }
                case I_DLY_IN: case I_DLY_BUF: case I_DLY_TI: {
                    const InputDesc in = (op == I_DLY_IN) ? input_desc(p, ins.w) : InputDesc{nullptr, 0, 0};
                    const BufferDesc bd = (op == I_DLY_BUF) ? p.buffers[ins.w] : BufferDesc{nullptr, 0};
#pragma unroll
                    for (int w = 0; w < VW; w++) {
                        const float d[4] = {a[w].x, a[w].y, a[w].z, a[w].w};
                        const float sv[4] = {b[w].x, b[w].y, b[w].z, b[w].w};
                        float r[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            unsigned long long o;
                            const bool live = delay_origin(d[i], t0g + 4 * w + i, p.sparkle_delay, &o);
                            float v = 0.0f;
                            if (live) v = (op == I_DLY_IN) ? load_input(in, o) : (op == I_DLY_BUF) ? bd.data[o & bd.mask] : sv[i];
                            r[i] = v;
                        }
                        REG(dst, w) = make_float4(r[0], r[1], r[2], r[3]);
                    }
                    break;
                }
                default: break;
            }
            ins = nxt;
        }
#undef REG
    }
}

// fold: ((x0 + x1) + x2) + ... over consecutive planes of one extension instance (the "voices summed to one slot" mix),
// in the chain's own order.  HBM-bound: reads count x 4 B per sample.
// volatile asm: keeps the loads of a batch ahead of the adds (the compiler otherwise sinks each load next to its use to
// save registers, which leaves only a few loads in flight per thread)
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// One thread = ONE sample, walking the planes in order.  The sum is a left fold (bit-exact with the Sum2 chain it stands
// for), so the only parallelism is across samples, and a block is at most 64 Ki - 1 Mi samples: ~440 - 7,000 samples per SM.
// What hides DRAM latency is therefore the depth of the per-thread load pipeline, and scalar lanes afford the deepest one
// for the registers: FB loads of batch k+1 are issued before the FB ordered adds of batch k (2 x FB x 4 B in flight per
// thread; a float4-per-thread version at FB = 16 has the same bytes in flight per thread but a quarter of the threads).
constexpr int FOLD_THREADS = 256;
constexpr int FOLD_FB = 32;

__global__ void __launch_bounds__(FOLD_THREADS)
fold_kernel(const BufferDesc* __restrict__ bufs, unsigned first, unsigned count, unsigned out_buf,
            unsigned long long lo, unsigned long long hi) {
    const BufferDesc b0 = bufs[first];
    const BufferDesc ob = bufs[out_buf];
    const unsigned long long stride = (unsigned long long)(bufs[first + 1].data - b0.data);
    const unsigned long long n = hi - lo;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long t = lo + g;
        const float* base = b0.data + (t & b0.mask);
        float acc = ld_stream_f1(base);
        unsigned i = 1;
        constexpr int FB = FOLD_FB;
        float cur[FB];
        const unsigned n_batches = (count - 1) / FB;
        if (n_batches) {
#pragma unroll
            for (int q = 0; q < FB; q++) cur[q] = ld_stream_f1(base + (unsigned long long)(i + q) * stride);
            for (unsigned bi = 0; bi < n_batches; bi++) {
                float nxt[FB];
                const unsigned j = i + FB;
                if (bi + 1 < n_batches) {
#pragma unroll
                    for (int q = 0; q < FB; q++) nxt[q] = ld_stream_f1(base + (unsigned long long)(j + q) * stride);
                }
#pragma unroll
                for (int q = 0; q < FB; q++) acc = __fadd_rn(acc, cur[q]);
                if (bi + 1 < n_batches) {
#pragma unroll
                    for (int q = 0; q < FB; q++) cur[q] = nxt[q];
                }
                i = j;
            }
        }
        for (; i < count; i++) acc = __fadd_rn(acc, ld_stream_f1(base + (unsigned long long)i * stride));
        ob.data[t & ob.mask] = acc;
    }
}

cudaError_t launch_fold(const BufferDesc* d_bufdesc, unsigned first, unsigned count, unsigned out_buf,
                        unsigned long long lo, unsigned long long hi, int sm_count, cudaStream_t stream) {
    if (hi <= lo) return cudaSuccess;
    unsigned long long blocks = (hi - lo + FOLD_THREADS - 1) / FOLD_THREADS;
    if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
    fold_kernel<<<(unsigned)blocks, FOLD_THREADS, 0, stream>>>(d_bufdesc, first, count, out_buf, lo, hi);
    return cudaGetLastError();
}

// per device, called when a renderer is created on it
cudaError_t interp_init_device() {
    return cudaFuncSetAttribute(interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
}

cudaError_t launch_interp(const InterpParams& p_in, unsigned n_regs, int sm_count, cudaStream_t stream) {
    if (p_in.n_groups == 0) return cudaSuccess;
    InterpParams p = p_in;
    const unsigned threads = INTERP_THREADS;
    p.n_regs = n_regs ? n_regs : 1;
    size_t smem = (size_t)p.n_regs * INTERP_VW * threads * sizeof(float4);
    const size_t prog_bytes = (size_t)p.n_instr * sizeof(uint4);
    p.prog_in_smem = (prog_bytes <= 48 * 1024 && smem + prog_bytes <= 200 * 1024) ? 1u : 0u;
    if (p.prog_in_smem) smem += prog_bytes;
    unsigned long long blocks = (p.n_groups + threads - 1) / threads;
    // persistent-style grid: a multiple of the SM count, enough CTAs per SM to cover latency
    unsigned long long per_sm = 227ull * 1024ull / (smem + 1024);
    if (per_sm > 12) per_sm = 12;
    if (per_sm < 1) per_sm = 1;
    unsigned long long cap = (unsigned long long)sm_count * per_sm / p.n_strands;   // rounded down: never a second wave of a few CTAs
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    interp_kernel<<<dim3((unsigned)blocks, p.n_strands), threads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace frb
