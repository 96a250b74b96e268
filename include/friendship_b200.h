/* friendship_b200.h — C ABI of the B200-native renderer for libfriendship's effect tree.
 *
 * This header is the drop-in boundary for the reference's `Renderer` trait
 * (reference: src/render/renderer.rs:6-17) and the `GraphWatcher` trait it inherits
 * (reference: src/routing/graphwatcher.rs:4-9).  A Rust `impl Renderer for B200Renderer`
 * (see INTEGRATION.md and rust/) binds exactly these entry points; tests/ and bench.py bind
 * them through Python ctypes.  Plain pointers and sizes only; nothing throws across the ABI.
 *
 * Vocabulary follows the reference:
 *   - node handle: u32, 0 == "toplevel" (the graph's own inputs/outputs), reference
 *     src/routing/nullable_int.rs:27-31, src/routing/routegraph.rs:330-343.
 *   - edge: {from, to, from_slot, to_slot}, reference src/routing/routegraph.rs:20-25,38-44.
 *   - primitive effects: reference src/routing/effect.rs:86-112 (enum PrimitiveEffect).
 *   - F32Constant's value is f32::from_bits(from_slot), reference src/render/reference.rs:217-220.
 */
#ifndef FRIENDSHIP_B200_H
#define FRIENDSHIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (the reference's methods return () and panic; we return codes) ---- */
#define FRB_OK                 0
#define FRB_E_BAD_HANDLE      -1   /* unknown node / effect key (reference would panic: reference.rs:131,145,186) */
#define FRB_E_INPUT_TOO_LONG  -2   /* input row longer than n_times (reference.rs:71 assert) */
#define FRB_E_INPUT_GAP       -3   /* slot fed now but skipped earlier: len != idx (reference.rs:69 assert_eq) */
#define FRB_E_BAD_SLOT        -4   /* from_slot != 0 on a single-output primitive (reference.rs:199,223,... assert) */
#define FRB_E_CUDA            -5
#define FRB_E_INVALID         -6   /* malformed arguments */
#define FRB_E_UNSUPPORTED     -7   /* graph shape the device schedule cannot express */
#define FRB_E_EXISTS          -8   /* handle / key already defined */
#define FRB_E_NO_DEVICE       -9   /* no CUDA device: there is no CPU fallback */

/* ---- node kinds ---- */
/* the seven primitives of reference src/routing/effect.rs:86-112 */
#define FRB_KIND_DELAY        0u
#define FRB_KIND_F32CONSTANT  1u
#define FRB_KIND_SUM2         2u
#define FRB_KIND_MULTIPLY     3u
#define FRB_KIND_DIVIDE       4u
#define FRB_KIND_MODULO       5u
#define FRB_KIND_MINIMUM      6u
/* a nested effect (reference EffectData::RouteGraph, effect.rs:79-83); `key` names a definition */
#define FRB_KIND_EFFECT       16u
/* extension nodes named by the north star; NOT in the reference */
#define FRB_KIND_OSCBANK      32u  /* 0 inputs, n_voices outputs; `key` names a bank definition */
#define FRB_KIND_DIRECTFORM   33u  /* n_lanes inputs, n_lanes outputs; per-lane biquad (Direct Form I) */
#define FRB_KIND_FBDELAY      34u  /* n_lanes inputs, n_lanes outputs; y[n] = x[n] + g*y[n-D] per lane */

/* Edge, verbatim from reference src/routing/routegraph.rs:38-44 + :20-25 (handle 0 = toplevel). */
typedef struct frb_edge {
    uint32_t from;
    uint32_t to;
    uint32_t from_slot;
    uint32_t to_slot;
} frb_edge;

/* One node of a nested effect definition (what RefRenderer::make_node walks, reference.rs:98-113). */
typedef struct frb_node {
    uint32_t handle;   /* non-zero */
    uint32_t kind;     /* FRB_KIND_* */
    uint64_t key;      /* definition key for EFFECT / OSCBANK / DIRECTFORM / FBDELAY, else 0 */
} frb_node;

typedef struct frb_config {
    int32_t  device;          /* CUDA device ordinal */
    uint32_t flags;           /* FRB_FLAG_* */
    uint32_t osc_anchor;      /* oscillator re-anchor interval in samples (0 = default: 128; a multiple of 16, <= 256).
                                 Banks whose voices have at most one partial each are always re-anchored every 8 */
    uint32_t n_devices;       /* 0 or 1: one B200.  N > 1: devices device .. device + N - 1 behind this one handle — the
                                 oscillator-bank voices are sharded round-robin over them (voice v on device v mod N), each
                                 device renders the sub-graph of its voices and stores its mix block straight into a slab in
                                 the first device's HBM over NVLink, where the rows are summed in device order (K5).  Every
                                 output slot must then be linear in the bank lanes (sums, gains, delays, linear filters) or
                                 use no lane at all; otherwise frb_fill_buffer returns FRB_E_UNSUPPORTED.  Output and
                                 device-resident inputs live on the first device.  The reference's caller owns ONE renderer
                                 (dispatch.rs:99-106): this keeps it that way. */
} frb_config;
#define FRB_FLAG_SPARKLE_DELAY 1u  /* negative / NaN delay amounts yield 0.0 (reference sparkle.rs:525-542)
                                      instead of clamping to delay 0 (reference.rs:205-210, the default) */
#define FRB_FLAG_SPARKLE_MIN  32u  /* Minimum as the reference's JIT renderer computes it, select(a ULT b, a, b): a NaN in either
                                      operand yields a (reference sparkle.rs:492-498), instead of f32::min = minNum, which
                                      yields the operand that is not NaN (reference.rs:242-248, the default) */
#define FRB_FLAG_NO_TENSOR_OSC 64u /* oscillator banks never take the matrix-product (tensor-core) kernels.  Those work in tiles of
                                    * 16,384 samples of a voice: the right unit for long renders of big banks (cfg4: 7.6x the
                                    * resonator kernel), the wrong one for a big bank streamed in short real-time calls, where every
                                    * call computes whole tiles and keeps a slice.  Which kernel renders a sample is a property of
                                    * the bank and this flag only — never of how a render is cut into calls. */
#define FRB_FLAG_NO_JIT        2u  /* always interpret stage programs; never compile them (see frb_jit_cubin_size) */
#define FRB_FLAG_JIT_EAGER     4u  /* compile a stage program the first time it runs (default: once it is hot) */
#define FRB_FLAG_NO_CHAIN_FUSION 8u /* run DirectForm -> FbDelay chains as two kernels (16 B per lane-sample) even where the
                                      fused kernel (8 B) applies; same bits either way (tests compare the two) */
#define FRB_FLAG_NO_EXCITER_FUSION 16u /* keep one-partial oscillator banks on their own kernel and rings even where a fused
                                      chain could evaluate them itself; same bits either way (tests compare the two) */

/* Oscillator bank definition (extension).  Voice v owns partials [voice_offsets[v], voice_offsets[v+1]).
 *   out_v(t) = sum_p amp_p * min(t/attack_p, 1) * exp(-t/tau_p) * sin(2*pi*freq_p*t/sample_rate + phase_p)
 * attack_p <= 0 means no attack ramp; tau_p <= 0 or +inf means no decay.  t is the absolute sample index.
 * At most 65,535 voices per bank (FRB_E_UNSUPPORTED above; any number of banks). */
typedef struct frb_oscbank_desc {
    uint32_t        n_voices;
    uint32_t        reserved;
    uint64_t        n_partials;
    double          sample_rate;
    const uint64_t* voice_offsets;   /* n_voices + 1 entries, non-decreasing, last == n_partials */
    const double*   freq_hz;         /* n_partials */
    const float*    amp;             /* n_partials */
    const float*    phase;           /* n_partials, radians */
    const float*    attack;          /* n_partials, samples */
    const float*    tau;             /* n_partials, samples */
} frb_oscbank_desc;

/* Direct Form I biquad bank (extension): per lane
 *   y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2],   x, y == 0 for n < 0. */
typedef struct frb_directform_desc {
    uint32_t     n_lanes;
    uint32_t     reserved;
    const float* b0; const float* b1; const float* b2; const float* a1; const float* a2;  /* n_lanes each */
} frb_directform_desc;

/* Feedback delay bank (extension): per lane  y[n] = x[n] + g * y[n - D],  D >= 1, y == 0 for n < 0.
 * At most 65,535 lanes per bank (FRB_E_UNSUPPORTED above). */
typedef struct frb_fbdelay_desc {
    uint32_t        n_lanes;
    uint32_t        reserved;
    const uint32_t* delay;   /* n_lanes, samples, >= 1 */
    const float*    gain;    /* n_lanes */
} frb_fbdelay_desc;

typedef struct frb_renderer frb_renderer;

/* ---- lifetime: replaces `SparkleRenderer::default()` moved into Dispatch::new (reference dispatch.rs:99-106) ---- */
frb_renderer* frb_create(const frb_config* cfg);          /* NULL on failure; see frb_last_error(NULL) */
void          frb_destroy(frb_renderer* r);
const char*   frb_last_error(const frb_renderer* r);      /* r may be NULL for creation errors */

/* ---- definitions consumed by on_add_node ---- */
/* Nested effect body = EffectData::RouteGraph (reference effect.rs:79-83); children must be defined first.
 * The definition is deep-copied into each node that instantiates it (reference.rs:98-113). */
int frb_define_effect(frb_renderer* r, uint64_t key, const frb_node* nodes, uint32_t n_nodes,
                      const frb_edge* edges, uint32_t n_edges);
int frb_define_oscbank(frb_renderer* r, uint64_t key, const frb_oscbank_desc* desc);
int frb_define_directform(frb_renderer* r, uint64_t key, const frb_directform_desc* desc);
int frb_define_fbdelay(frb_renderer* r, uint64_t key, const frb_fbdelay_desc* desc);

/* ---- GraphWatcher (reference src/routing/graphwatcher.rs:4-9; RefRenderer impl reference.rs:116-137) ---- */
int frb_add_node(frb_renderer* r, uint32_t handle, uint32_t kind, uint64_t key);   /* on_add_node */
int frb_del_node(frb_renderer* r, uint32_t handle);                                /* on_del_node */
int frb_add_edge(frb_renderer* r, frb_edge e);                                     /* on_add_edge */
int frb_del_edge(frb_renderer* r, frb_edge e);                                     /* on_del_edge */

/* ---- Renderer::fill_buffer (reference src/render/renderer.rs:6-17; RefRenderer reference.rs:46-86) ----
 * out: host, row-major [n_slots x n_times] f32 (what Dispatch allocates, dispatch.rs:149).
 * inputs: jagged rows (Jagged2<f32>): row r feeds external-input slot r and occupies
 *         in_data[in_row_offsets[r] .. in_row_offsets[r+1]); in_row_offsets has n_in_rows+1 entries.
 * idx: absolute index of the first sample; idx != previous end is a seek (renderer.rs:12-15).
 * Synchronous: on return `out` is filled. */
int frb_fill_buffer(frb_renderer* r, float* out, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                    const float* in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows);

/* Same contract with `out` and `in_data` in DEVICE memory on the renderer's device (in_row_offsets stays on
 * the host).  The call returns after enqueueing on the renderer's stream; frb_sync() waits.  Used to time the
 * device-resident path and to hand the mixed block to a collective without a host round trip. */
int frb_fill_buffer_device(frb_renderer* r, float* d_out, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                           const float* d_in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows);
int frb_sync(frb_renderer* r);
void* frb_stream(frb_renderer* r);   /* the cudaStream_t the renderer launches on */

/* ---- N4: streaming render — the step after the path ----
 * The reference hands every rendered buffer to Client::audio_rendered (src/dispatch.rs:151, src/client/client.rs:8-15);
 * a long offline render (BASELINE.json configs[4]: 11.52 M samples) must move its output off the device without
 * stalling the render.  frb_render_stream renders [idx, idx + n_total) as consecutive fill_buffer calls of `block`
 * samples (the last one shorter) — same results, bit for bit, as those calls — pipelined: while block k renders, block
 * k-1 travels device->host into one of two pinned staging buffers on a copy stream and the host thread runs `sink` on
 * it; the external inputs of block k+1 are fetched through `source` into pinned memory meanwhile.
 *   source(user, rows, n_in_rows, n_times, idx): fill rows[r * n_times + i] = external input slot r at time idx + i;
 *           may be NULL when n_in_rows == 0.  Non-zero return aborts the render (FRB_E_INVALID).
 *   sink(user, block, n_slots, n_times, idx): block is row-major [n_slots x n_times], valid during the call
 *           (= audio_rendered(buffer, idx)).  Non-zero return aborts the render. */
typedef int (*frb_source_fn)(void* user, float* rows, uint32_t n_in_rows, uint64_t n_times, uint64_t idx);
typedef int (*frb_sink_fn)(void* user, const float* block, uint32_t n_slots, uint64_t n_times, uint64_t idx);
int frb_render_stream(frb_renderer* r, uint32_t n_slots, uint64_t idx, uint64_t n_total, uint64_t block,
                      uint32_t n_in_rows, frb_source_fn source, frb_sink_fn sink, void* user);

/* ---- K5: cross-GPU mix without a separate collective ----
 * One process per GPU.  Rank 0 owns a slab [world x n_slots x n_times] in its HBM; every rank renders its shard of
 * the voices with frb_fill_buffer_device pointing at ITS row of that slab (the stage kernel's output stores go
 * straight over NVLink through a CUDA-IPC mapping: the mix epilogue and the transfer are one kernel), and after a
 * barrier rank 0 sums the rows in rank order (deterministic, unlike a tree reduce) with frb_sum_rows.
 * frb_ipc_export / frb_ipc_open wrap cudaIpcGetMemHandle / cudaIpcOpenMemHandle (64-byte opaque handle). */
/* plain cudaMalloc / cudaFree on the renderer's device: an IPC handle names a whole allocation, so the slab must not
 * be a sub-block of a pooling allocator */
int frb_device_alloc(frb_renderer* r, uint64_t bytes, void** d_ptr_out);
int frb_device_free(frb_renderer* r, void* d_ptr);
int frb_ipc_export(frb_renderer* r, const void* d_ptr, unsigned char handle[64]);
int frb_ipc_open(frb_renderer* r, const unsigned char handle[64], void** d_ptr_out);
int frb_ipc_close(frb_renderer* r, void* d_ptr);
/* d_out[i] = ((rows[0][i] + rows[1][i]) + ...) + rows[n_rows-1][i], rows contiguous with stride row_stride floats */
int frb_sum_rows(frb_renderer* r, float* d_out, const float* d_rows, uint32_t n_rows, uint64_t row_stride, uint64_t n);

/* ---- introspection for parity tests of routing order / buffer indexing / delay-line offsets ---- */
/* Builds (if dirty) the device schedule for `n_slots` outputs and copies it out as u32 words:
 * see libfriendship_b200/csrc/schedule.hpp for the record layout.  Returns the number of words the
 * schedule needs (may exceed cap; nothing is written beyond cap), or a negative status. */
int64_t frb_dump_schedule(frb_renderer* r, uint32_t n_slots, uint32_t* words, uint64_t cap);

/* Voice sharding (frb_config::n_devices): how the outputs of the graph depend on the oscillator-bank lanes —
 * 0: linearly (sums, gains, delays, linear filters: the voices may be rendered on different devices and added),
 * 1: not at all, 2: otherwise (n_devices > 1 refuses such a graph) — and the schedule device `rank` of `world` would run:
 * the graph restricted to the voices it owns (v mod world == rank), which are lanes 0, 1, ... of its compact bank. */
int frb_lane_use(frb_renderer* r, uint32_t n_slots);
int64_t frb_dump_schedule_shard(frb_renderer* r, uint32_t n_slots, uint32_t rank, uint32_t world, uint32_t* words, uint64_t cap);

/* The stage JIT (the B200 counterpart of the reference's LLVM JIT, src/render/sparkle.rs): CUDA source generated for
 * stage `stage` of the schedule for `n_slots` outputs, and the size of the sm_100a cubin NVRTC builds from it
 * (negative status on failure; needs no GPU).  frb_jit_source writes at most cap bytes (NUL-terminated) and returns
 * the length needed. */
int64_t frb_jit_source(frb_renderer* r, uint32_t n_slots, uint32_t stage, char* out, uint64_t cap);
int64_t frb_jit_cubin_size(frb_renderer* r, uint32_t n_slots, uint32_t stage);
/* Statements the compiler sees for that stage.  Runs of like instruction groups in a stage program (the terms of a
 * Sum2 chain, the voices of a mix) are compiled as loops whose trip counts and operands come from a device table, so
 * this is a measure of the program's STRUCTURE, not of its length: a 9,000-node chain is 11.  NVRTC's time grows faster
 * than linearly in one straight body (100: 1 s, 200: 2 s, 400: 8 s), so a long APERIODIC program is cut into chunk functions
 * of 64 statements (400: 2 s, 1,000: 6 s, 2,000: 16 s); a stage above FRB_JIT_MAX_CODE is never compiled (interpreted for
 * good), and one above FRB_JIT_MAX_SYNC_CODE is compiled beside the render loop — the interpreter serving meanwhile —
 * rather than on the render thread (unless FRB_FLAG_JIT_EAGER asks for that). */
#define FRB_JIT_MAX_SYNC_CODE 128
#define FRB_JIT_MAX_CODE 4096
int64_t frb_jit_code_instructions(frb_renderer* r, uint32_t n_slots, uint32_t stage);

/* Counters since creation: kernel launches, bytes H2D, bytes D2H. */
typedef struct frb_stats {
    uint64_t kernel_launches;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t schedule_builds;
    uint64_t osc_launches;
    uint64_t interp_launches;
    uint64_t scan_launches;
    uint64_t jit_launches;      /* stage launches that ran a JIT-compiled kernel instead of the interpreter */
    uint64_t chain_launches;    /* fused DirectForm -> FbDelay launches (counted in scan_launches too) */
    uint64_t osc_tensor_launches; /* of osc_launches: the matrix-product oscillator kernels (tensor cores: K1T tcgen05 / K1G mma.sync) */
} frb_stats;
int frb_get_stats(const frb_renderer* r, frb_stats* out);

/* Time (ms, CUDA events on the renderer's stream) the most recent fill spent in each kernel family. */
typedef struct frb_timing {
    float osc_ms;
    float interp_ms;
    float scan_ms;
    float total_ms;
} frb_timing;
int frb_set_profiling(frb_renderer* r, int enabled);
int frb_get_timing(const frb_renderer* r, frb_timing* out);

const char* frb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FRIENDSHIP_B200_H */
