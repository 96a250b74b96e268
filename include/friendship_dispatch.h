/* friendship_dispatch.h — C ABI of the host-side control path that SURROUNDS the hot path:
 * the reference's `Dispatch` (src/dispatch.rs:18-162) with its `RouteGraph` validation (src/routing/routegraph.rs),
 * effect loading (`Effect::from_id`, src/routing/effect.rs:135-220) and `ResMan` (src/resman.rs), restated in C++
 * because the image has no Rust toolchain.  The renderer underneath is the B200 renderer of friendship_b200.h.
 *
 * One call per OSC message of src/dispatch.rs:31-86.  EffectIds cross the ABI in the reference's own JSON wire
 * format (serde): {"name":"Delay","sha256":null,"urls":["primitive:///Delay"]}.
 */
#ifndef FRIENDSHIP_DISPATCH_H
#define FRIENDSHIP_DISPATCH_H

#include "friendship_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* dispatch.rs:89-95 Error::{RouteGraphError, EffectError}; values continue after the FRB_E_* renderer codes */
#define FRD_E_WOULD_CYCLE             -101   /* routegraph.rs:48 */
#define FRD_E_NODE_IN_USE             -102   /* routegraph.rs:50 */
#define FRD_E_NODE_EXISTS             -103   /* routegraph.rs:52 */
#define FRD_E_SLOT_ALREADY_CONNECTED  -104   /* routegraph.rs:54 */
#define FRD_E_NO_SUCH_NODE            -105   /* routegraph.rs:56 */
#define FRD_E_NO_SUCH_SLOT            -106   /* routegraph.rs:58 */
#define FRD_E_NO_MATCHING_EFFECT      -107   /* effect.rs:20 */
#define FRD_E_BAD_MESSAGE             -108   /* malformed JSON / arguments */

/* Client callbacks (reference src/client/client.rs:8-15); any may be NULL (the trait's methods default to no-ops). */
typedef struct frd_client {
    void* user;
    /* audio_rendered(buffer, idx): buffer is row-major [n_slots x n_times], valid during the call */
    void (*audio_rendered)(void* user, const float* buffer, uint32_t n_slots, uint64_t n_times, uint64_t idx);
    /* node_meta(handle, meta) / node_id(handle, id): JSON in the reference's serde format */
    void (*node_meta)(void* user, uint32_t handle, const char* meta_json);
    void (*node_id)(void* user, uint32_t handle, const char* id_json);
} frd_client;

typedef struct frd_dispatch frd_dispatch;

/* Dispatch::new(renderer, client) (dispatch.rs:99-106).  cfg as for frb_create (device -1: planning only). */
frd_dispatch* frd_create(const frb_config* cfg, const frd_client* client);
void          frd_destroy(frd_dispatch* d);
const char*   frd_last_error(const frd_dispatch* d);
frb_renderer* frd_renderer(frd_dispatch* d);

/* OscRouteGraph (dispatch.rs:47-62, handled at :114-146) */
int frd_add_node(frd_dispatch* d, uint32_t handle, const char* effect_id_json);
int frd_add_edge(frd_dispatch* d, frb_edge e);
int frd_del_node(frd_dispatch* d, uint32_t handle);
int frd_del_edge(frd_dispatch* d, frb_edge e);
int frd_query_meta(frd_dispatch* d, uint32_t handle);
int frd_query_id(frd_dispatch* d, uint32_t handle);
/* OscRenderer::RenderRange(range, num_slots, inputs) (dispatch.rs:66-76, :147-153) */
int frd_render_range(frd_dispatch* d, uint64_t start, uint64_t end, uint32_t n_slots,
                     const float* in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows);
/* N4: RenderRange over [start, end) delivered as consecutive audio_rendered(buffer, idx) calls of
 * `block` samples (no inputs) — what a client sending back-to-back RenderRange messages gets, bit for bit, but
 * pipelined through frb_render_stream: block k renders while block k-1 is copied to pinned host memory and handed to
 * the client.  This is how an offline render (BASELINE.json configs[4]) leaves the device. */
int frd_render_stream(frd_dispatch* d, uint64_t start, uint64_t end, uint32_t n_slots, uint64_t block);
/* OscResMan::AddDir (dispatch.rs:80-86, :155-159) */
int frd_add_dir(frd_dispatch* d, const char* path);

/* N4 output sink: a RIFF/WAVE writer (IEEE float32, slot s = channel s) usable as a Client.  The reference leaves
 * file output to the client (README.md:22-26); this is that client.  frd_wav_audio_rendered has the signature of
 * frd_client.audio_rendered with `user` = the frd_wav*; blocks must arrive in order (idx is not used to seek). */
typedef struct frd_wav frd_wav;
frd_wav* frd_wav_open(const char* path, uint32_t n_channels, uint32_t sample_rate);   /* NULL on failure */
int      frd_wav_write(frd_wav* w, const float* buffer, uint32_t n_slots, uint64_t n_times);
void     frd_wav_audio_rendered(void* user, const float* buffer, uint32_t n_slots, uint64_t n_times, uint64_t idx);
int      frd_wav_close(frd_wav* w);          /* patches the header, frees w; FRB_OK iff every write succeeded */
const char* frd_wav_error(const frd_wav* w);

/* helpers: SHA-256 of a file (what tests/load_effect.rs:84-88 computes) and the graph as an AdjList JSON */
int frd_sha256_file(const char* path, uint8_t out[32]);
/* writes at most cap bytes (NUL-terminated); returns the length needed */
int64_t frd_adjlist_json(frd_dispatch* d, char* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
