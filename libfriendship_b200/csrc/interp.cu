// interp.cu — K2/K3: fused elementwise + Delay pass over a block of time (sm_100a).
//
// One launch evaluates one stage program (schedule.hpp) for every sample of a time block.  Each thread owns
// four consecutive samples (absolute time aligned to 4, so undelayed plane reads/writes are 128-bit and
// coalesced); the program's registers live in shared memory as float4 columns regs[r][thread] (conflict-free),
// so intermediates of the fused nodes never touch HBM.  Delay is an indexed read at t - floor(d) from the
// external-input history or from a ring buffer in HBM written by an earlier stage.
//
// Arithmetic is the reference's, bit for bit (reference src/render/reference.rs:197-262): IEEE f32 with no
// FMA contraction (__fmul_rn/__fadd_rn/__fdiv_rn never contract), fmodf for `%`, fminf for f32::min.
#include "interp.cuh"

namespace frb {

__device__ __forceinline__ float load_input(const InputDesc& in, unsigned long long t) {
    // value of external input at absolute time t; 0 outside the stored history (reference.rs:90-96)
    return (t >= in.base_time && t < in.end_time) ? in.data[t - in.base_time] : 0.0f;
}

// Delay clamps, reference.rs:200-212.  Returns false when the output is 0 without reading the source.
__device__ __forceinline__ bool delay_origin(float d, unsigned long long t, bool sparkle, unsigned long long* origin) {
    if (d >= 18446744073709551616.0f) return false;               // >= 2^64: indexing negative time
    unsigned long long di;
    if (!(d >= 0.0f)) {                                            // negative or NaN
        if (sparkle) return false;                                 // sparkle.rs:525-542 returns 0.0
        di = 0ull;                                                 // reference.rs:206-207 (NaN: saturating cast -> 0)
    } else {
        di = __float2ull_rz(d);                                    // truncation toward zero (:208-210)
    }
    if (t < di) return false;                                      // checked_sub -> None -> 0 (:212)
    *origin = t - di;
    return true;
}

__device__ __forceinline__ float op_mod(float a, float b) {
    float r = fmodf(a, b);                                         // Rust `%` on f32
    return (r < 0.0f) ? __fadd_rn(r, b) : r;                       // reference.rs:255-261
}

__global__ void __launch_bounds__(INTERP_THREADS)
interp_kernel(InterpParams p) {
    extern __shared__ float4 s_regs[];
    const unsigned tid = threadIdx.x;
    const unsigned nthr = blockDim.x;
    // the program is staged in shared memory behind the register columns when it fits (broadcast LDS instead of
    // a dependent global load per interpreted instruction)
    const uint4* prog = reinterpret_cast<const uint4*>(p.program);
    if (p.prog_in_smem) {
        uint4* s_prog = reinterpret_cast<uint4*>(s_regs + (size_t)p.n_regs * nthr);
        for (unsigned i = tid; i < p.n_instr; i += nthr) s_prog[i] = __ldg(prog + i);
        __syncthreads();
        prog = s_prog;
    }

    for (unsigned long long g = (unsigned long long)blockIdx.x * nthr + tid; g < p.n_groups;
         g += (unsigned long long)gridDim.x * nthr) {
        const unsigned long long t = p.t_begin + 4ull * g;        // absolute time of element 0; multiple of 4
#define REG(r) s_regs[(r) * nthr + tid]
        uint4 ins = prog[0];
        for (unsigned pc = 0;; pc++) {
            const uint4 nxt = prog[pc + 1];                       // prefetch (programs end with two I_ENDs)
            const unsigned op = ins.x & 0xFFu, flags = (ins.x >> 8) & 0xFFu, dst = ins.x >> 16;
            if (op == I_END) break;
            float4 a, b;
            if (op != I_LDIN && op != I_LDBUF && op != I_FOLD) {
                if (flags & IF_A_IMM) { float v = __uint_as_float(ins.y); a = make_float4(v, v, v, v); }
                else a = REG(ins.y);
            }
            switch (op) {
                case I_ADD: case I_MUL: case I_DIV: case I_MOD: case I_MIN: {
                    if (flags & IF_B_IMM) { float v = __uint_as_float(ins.z); b = make_float4(v, v, v, v); }
                    else b = REG(ins.z);
                    float4 r;
                    if (op == I_ADD) r = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
                    else if (op == I_MUL) r = make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
                    else if (op == I_DIV) r = make_float4(__fdiv_rn(a.x, b.x), __fdiv_rn(a.y, b.y), __fdiv_rn(a.z, b.z), __fdiv_rn(a.w, b.w));
                    else if (op == I_MOD) r = make_float4(op_mod(a.x, b.x), op_mod(a.y, b.y), op_mod(a.z, b.z), op_mod(a.w, b.w));
                    else r = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w));
                    REG(dst) = r;
                    break;
                }
                case I_MOV: REG(dst) = a; break;
                case I_LDIN: {
                    const InputDesc in = p.inputs[ins.w];
                    float4 r;
                    if (t >= in.base_time && t + 4 <= in.end_time) {
                        r = *reinterpret_cast<const float4*>(in.data + (t - in.base_time));   // base_time % 4 == 0
                    } else {
                        r = make_float4(load_input(in, t), load_input(in, t + 1), load_input(in, t + 2), load_input(in, t + 3));
                    }
                    REG(dst) = r;
                    break;
                }
                case I_LDBUF: {
                    const BufferDesc bd = p.buffers[ins.w];
                    REG(dst) = *reinterpret_cast<const float4*>(bd.data + (t & bd.mask));
                    break;
                }
                case I_FOLD: {
                    // left fold of ins.z consecutive planes starting at buffer ins.y: the Sum2 chain's own order
                    // the lanes of one extension instance are one allocation: plane i = base + i * stride
                    const BufferDesc b0 = p.buffers[ins.y];
                    const unsigned long long stride = (unsigned long long)(p.buffers[ins.y + 1].data - b0.data);
                    const float* base = b0.data + (t & b0.mask);
                    float4 acc = *reinterpret_cast<const float4*>(base);
                    unsigned i = 1;
                    for (; i + 16 <= ins.z; i += 16) {             // 16 independent plane loads in flight, then the ordered adds
                        float4 v[16];
#pragma unroll
                        for (int q = 0; q < 16; q++) v[q] = *reinterpret_cast<const float4*>(base + (unsigned long long)(i + q) * stride);
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            acc.x = __fadd_rn(acc.x, v[q].x); acc.y = __fadd_rn(acc.y, v[q].y);
                            acc.z = __fadd_rn(acc.z, v[q].z); acc.w = __fadd_rn(acc.w, v[q].w);
                        }
                    }
                    for (; i < ins.z; i++) {
                        const float4 v = *reinterpret_cast<const float4*>(base + (unsigned long long)i * stride);
                        acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
                        acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
                    }
                    REG(dst) = acc;
                    break;
                }
                case I_STBUF: {
                    const BufferDesc bd = p.buffers[ins.w];
                    *reinterpret_cast<float4*>(bd.data + (t & bd.mask)) = a;
                    break;
                }
                case I_STOUT: {
                    // out[slot][t - t0], only for t0 <= t < t1
                    float* row = p.out + (unsigned long long)ins.w * p.out_stride;
                    const float v[4] = {a.x, a.y, a.z, a.w};
                    if (t >= p.t0 && t + 4 <= p.t1 && p.out_vec_ok) {
                        *reinterpret_cast<float4*>(row + (t - p.t0)) = a;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (t + i >= p.t0 && t + i < p.t1) row[t + i - p.t0] = v[i];
                    }
                    break;
                }
                case I_DLY_IN: {
                    const InputDesc in = p.inputs[ins.w];
                    const float d[4] = {a.x, a.y, a.z, a.w};
                    float r[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        unsigned long long o;
                        r[i] = delay_origin(d[i], t + i, p.sparkle_delay, &o) ? load_input(in, o) : 0.0f;
                    }
                    REG(dst) = make_float4(r[0], r[1], r[2], r[3]);
                    break;
                }
                case I_DLY_BUF: {
                    const BufferDesc bd = p.buffers[ins.w];
                    const float d[4] = {a.x, a.y, a.z, a.w};
                    float r[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        unsigned long long o;
                        r[i] = delay_origin(d[i], t + i, p.sparkle_delay, &o) ? bd.data[o & bd.mask] : 0.0f;
                    }
                    REG(dst) = make_float4(r[0], r[1], r[2], r[3]);
                    break;
                }
                case I_DLY_TI: {
                    if (flags & IF_B_IMM) { float v = __uint_as_float(ins.z); b = make_float4(v, v, v, v); }
                    else b = REG(ins.z);
                    const float d[4] = {a.x, a.y, a.z, a.w};
                    const float s[4] = {b.x, b.y, b.z, b.w};
                    float r[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        unsigned long long o;
                        r[i] = delay_origin(d[i], t + i, p.sparkle_delay, &o) ? s[i] : 0.0f;
                    }
                    REG(dst) = make_float4(r[0], r[1], r[2], r[3]);
                    break;
                }
                default: break;
            }
            ins = nxt;
        }
#undef REG
    }
}

cudaError_t launch_interp(const InterpParams& p_in, unsigned n_regs, int sm_count, cudaStream_t stream) {
    if (p_in.n_groups == 0) return cudaSuccess;
    InterpParams p = p_in;
    const unsigned threads = INTERP_THREADS;
    p.n_regs = n_regs ? n_regs : 1;
    size_t smem = (size_t)p.n_regs * threads * sizeof(float4);
    const size_t prog_bytes = (size_t)p.n_instr * sizeof(uint4);
    p.prog_in_smem = (prog_bytes <= 48 * 1024 && smem + prog_bytes <= 200 * 1024) ? 1u : 0u;
    if (p.prog_in_smem) smem += prog_bytes;
    unsigned long long blocks = (p.n_groups + threads - 1) / threads;
    // persistent-style grid: a multiple of the SM count, enough CTAs per SM to cover latency
    unsigned long long per_sm = 227ull * 1024ull / (smem + 1024);
    if (per_sm > 12) per_sm = 12;
    if (per_sm < 1) per_sm = 1;
    unsigned long long cap = (unsigned long long)sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    interp_kernel<<<(unsigned)blocks, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace frb
