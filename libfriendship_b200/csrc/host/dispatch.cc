// dispatch.cc — Dispatch (reference src/dispatch.rs:18-162) over the B200 renderer; C ABI in friendship_dispatch.h.
#include <cstring>
#include <map>
#include <string>

#include "../../../include/friendship_dispatch.h"
#include "routing.hpp"
#include "wav.hpp"

using namespace frb::host;

struct frd_dispatch {
    RouteGraph routegraph;            // dispatch.rs:20
    frb_renderer* renderer = nullptr; // dispatch.rs:21
    ResMan resman;                    // dispatch.rs:24
    frd_client client{};              // dispatch.rs:27
    std::string err;
    std::map<const Effect*, uint64_t> effect_keys;   // nested definitions already handed to the renderer
    uint64_t next_key = 1;
    std::vector<NodeData> keep_alive;
};

static int rg_code(RgError e) {
    switch (e) {
        case RgError::None: return FRB_OK;
        case RgError::WouldCycle: return FRD_E_WOULD_CYCLE;
        case RgError::NodeInUse: return FRD_E_NODE_IN_USE;
        case RgError::NodeExists: return FRD_E_NODE_EXISTS;
        case RgError::SlotAlreadyConnected: return FRD_E_SLOT_ALREADY_CONNECTED;
        case RgError::NoSuchNode: return FRD_E_NO_SUCH_NODE;
        case RgError::NoSuchSlot: return FRD_E_NO_SUCH_SLOT;
        case RgError::NoMatchingEffect: return FRD_E_NO_MATCHING_EFFECT;
    }
    return FRD_E_BAD_MESSAGE;
}
static const char* rg_name(RgError e) {
    switch (e) {
        case RgError::WouldCycle: return "WouldCycle"; case RgError::NodeInUse: return "NodeInUse";
        case RgError::NodeExists: return "NodeExists"; case RgError::SlotAlreadyConnected: return "SlotAlreadyConnected";
        case RgError::NoSuchNode: return "NoSuchNode"; case RgError::NoSuchSlot: return "NoSuchSlot";
        case RgError::NoMatchingEffect: return "NoMatchingEffect"; default: return "ok";
    }
}
static uint32_t prim_kind(Primitive p) {
    switch (p) {
        case Primitive::Delay: return FRB_KIND_DELAY; case Primitive::F32Constant: return FRB_KIND_F32CONSTANT;
        case Primitive::Sum2: return FRB_KIND_SUM2; case Primitive::Multiply: return FRB_KIND_MULTIPLY;
        case Primitive::Divide: return FRB_KIND_DIVIDE; case Primitive::Modulo: return FRB_KIND_MODULO;
        case Primitive::Minimum: return FRB_KIND_MINIMUM;
    }
    return FRB_KIND_F32CONSTANT;
}
static int renderer_rc(frd_dispatch* d, int rc) {
    if (rc != FRB_OK) d->err = frb_last_error(d->renderer);
    return rc;
}

// What RefRenderer::make_node does with the Rc<Effect> it receives (reference.rs:98-113): nested graphs are
// handed to the renderer as definitions, children first; returns (kind, key) for the node itself.
static int kind_of(frd_dispatch* d, const NodeData& data, uint32_t* kind, uint64_t* key) {
    if (data->primitive) { *kind = prim_kind(*data->primitive); *key = 0; return FRB_OK; }
    auto it = d->effect_keys.find(data.get());
    if (it != d->effect_keys.end()) { *kind = FRB_KIND_EFFECT; *key = it->second; return FRB_OK; }
    std::vector<frb_node> nodes;
    for (auto& kv : data->graph->nodes()) {
        if (!kv.second.data) continue;
        uint32_t k; uint64_t ky;
        int rc = kind_of(d, kv.second.data, &k, &ky);
        if (rc != FRB_OK) return rc;
        nodes.push_back(frb_node{kv.first, k, ky});
    }
    std::vector<frb_edge> edges;
    for (auto& e : data->graph->edges()) edges.push_back(frb_edge{e.from, e.to, e.from_slot, e.to_slot});
    uint64_t nk = d->next_key++;
    int rc = frb_define_effect(d->renderer, nk, nodes.data(), (uint32_t)nodes.size(), edges.data(), (uint32_t)edges.size());
    if (rc != FRB_OK) return renderer_rc(d, rc);
    d->effect_keys[data.get()] = nk;
    d->keep_alive.push_back(data);
    *kind = FRB_KIND_EFFECT; *key = nk;
    return FRB_OK;
}

extern "C" {

frd_dispatch* frd_create(const frb_config* cfg, const frd_client* client) {
    frb_renderer* r = frb_create(cfg);
    if (!r) return nullptr;
    auto* d = new frd_dispatch();
    d->renderer = r;
    if (client) d->client = *client;
    return d;
}
void frd_destroy(frd_dispatch* d) {
    if (!d) return;
    frb_destroy(d->renderer);
    delete d;
}
const char* frd_last_error(const frd_dispatch* d) { return d ? d->err.c_str() : frb_last_error(nullptr); }
frb_renderer* frd_renderer(frd_dispatch* d) { return d ? d->renderer : nullptr; }

int frd_add_node(frd_dispatch* d, uint32_t handle, const char* effect_id_json) {      // dispatch.rs:115-119
    if (!d || !effect_id_json) return FRD_E_BAD_MESSAGE;
    EffectId id;
    try { id = EffectId::from_json(Json::parse(effect_id_json)); }
    catch (const std::exception& e) { d->err = std::string("bad EffectId: ") + e.what(); return FRD_E_BAD_MESSAGE; }
    RgError err = RgError::None;
    NodeData data = Effect::from_id(id, d->resman, &err);
    if (!data) { d->err = "EffectError(NoMatchingEffect(" + id.name + "))"; return rg_code(RgError::NoMatchingEffect); }
    err = d->routegraph.add_node(handle, data);
    if (err != RgError::None) { d->err = std::string("RouteGraphError(") + rg_name(err) + ")"; return rg_code(err); }
    uint32_t kind; uint64_t key;
    int rc = kind_of(d, data, &kind, &key);
    if (rc == FRB_OK) rc = renderer_rc(d, frb_add_node(d->renderer, handle, kind, key));    // on_add_node, dispatch.rs:202-204
    if (rc != FRB_OK) d->routegraph.del_node(handle);      // the renderer refused: the two graphs stay in step
    return rc;
}
int frd_add_edge(frd_dispatch* d, frb_edge e) {                                          // dispatch.rs:120-123
    if (!d) return FRD_E_BAD_MESSAGE;
    RgError err = d->routegraph.add_edge(Edge{e.from, e.to, e.from_slot, e.to_slot});
    if (err != RgError::None) { d->err = std::string("RouteGraphError(") + rg_name(err) + ")"; return rg_code(err); }
    const int rc = renderer_rc(d, frb_add_edge(d->renderer, e));
    if (rc != FRB_OK) d->routegraph.del_edge(Edge{e.from, e.to, e.from_slot, e.to_slot});   // refused (e.g. to_slot out of the renderer's range)
    return rc;
}
int frd_del_node(frd_dispatch* d, uint32_t handle) {                                     // dispatch.rs:124-127
    if (!d) return FRD_E_BAD_MESSAGE;
    RgError err = d->routegraph.del_node(handle);
    if (err != RgError::None) { d->err = std::string("RouteGraphError(") + rg_name(err) + ")"; return rg_code(err); }
    return renderer_rc(d, frb_del_node(d->renderer, handle));
}
int frd_del_edge(frd_dispatch* d, frb_edge e) {                                          // dispatch.rs:128-131
    if (!d) return FRD_E_BAD_MESSAGE;
    d->routegraph.del_edge(Edge{e.from, e.to, e.from_slot, e.to_slot});
    // the reference forwards unconditionally and RefRenderer panics when the target node is unknown
    // (reference.rs:131); RouteGraph::del_edge itself is silent, so an unknown target is reported, not fatal
    return renderer_rc(d, frb_del_edge(d->renderer, e));
}
int frd_query_meta(frd_dispatch* d, uint32_t handle) {                                   // dispatch.rs:132-138
    if (!d) return FRD_E_BAD_MESSAGE;
    NodeData data = d->routegraph.get_data(handle);
    if (!data) return FRB_OK;                       // "QueryMeta: no such effect" is only a warning
    if (d->client.node_meta) d->client.node_meta(d->client.user, handle, data->meta.to_json().dump().c_str());
    return FRB_OK;
}
int frd_query_id(frd_dispatch* d, uint32_t handle) {                                     // dispatch.rs:139-145
    if (!d) return FRD_E_BAD_MESSAGE;
    NodeData data = d->routegraph.get_data(handle);
    if (!data) return FRB_OK;
    if (d->client.node_id) d->client.node_id(d->client.user, handle, data->meta.id.to_json().dump().c_str());
    return FRB_OK;
}
int frd_render_range(frd_dispatch* d, uint64_t start, uint64_t end, uint32_t n_slots,
                     const float* in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows) {   // dispatch.rs:147-153
    if (!d || end < start) return FRD_E_BAD_MESSAGE;
    const uint64_t n_times = end - start;
    std::vector<float> buff((size_t)n_slots * n_times, 0.0f);        // ArrayBase::zeros (dispatch.rs:149)
    int rc = frb_fill_buffer(d->renderer, buff.data(), n_slots, n_times, start, in_data, in_row_offsets, n_in_rows);
    if (rc != FRB_OK) return renderer_rc(d, rc);
    if (d->client.audio_rendered) d->client.audio_rendered(d->client.user, buff.data(), n_slots, n_times, start);
    return FRB_OK;
}
static int stream_sink(void* user, const float* block, uint32_t n_slots, uint64_t n_times, uint64_t idx) {
    auto* d = static_cast<frd_dispatch*>(user);
    if (d->client.audio_rendered) d->client.audio_rendered(d->client.user, block, n_slots, n_times, idx);   // dispatch.rs:151
    return 0;
}
int frd_render_stream(frd_dispatch* d, uint64_t start, uint64_t end, uint32_t n_slots, uint64_t block) {
    if (!d || end < start || block == 0) return FRD_E_BAD_MESSAGE;
    return renderer_rc(d, frb_render_stream(d->renderer, n_slots, start, end - start, block, 0, nullptr, stream_sink, d));
}
int frd_add_dir(frd_dispatch* d, const char* path) {                                      // dispatch.rs:155-159
    if (!d || !path) return FRD_E_BAD_MESSAGE;
    d->resman.add_dir(path);
    return FRB_OK;
}
}  // extern "C"

struct frd_wav {
    WavWriter w;
    std::string err;
    bool failed = false;
};

extern "C" {

frd_wav* frd_wav_open(const char* path, uint32_t n_channels, uint32_t sample_rate) {
    if (!path) return nullptr;
    auto* w = new frd_wav();
    if (!w->w.open(path, n_channels, sample_rate, &w->err)) { delete w; return nullptr; }
    return w;
}
int frd_wav_write(frd_wav* w, const float* buffer, uint32_t n_slots, uint64_t n_times) {
    if (!w || (!buffer && n_times)) return FRD_E_BAD_MESSAGE;
    if (!w->w.write(buffer, n_slots, n_times, &w->err)) { w->failed = true; return FRD_E_BAD_MESSAGE; }
    return FRB_OK;
}
void frd_wav_audio_rendered(void* user, const float* buffer, uint32_t n_slots, uint64_t n_times, uint64_t) {
    frd_wav_write(static_cast<frd_wav*>(user), buffer, n_slots, n_times);
}
int frd_wav_close(frd_wav* w) {
    if (!w) return FRD_E_BAD_MESSAGE;
    const bool ok = w->w.close() && !w->failed;
    delete w;
    return ok ? FRB_OK : FRD_E_BAD_MESSAGE;
}
const char* frd_wav_error(const frd_wav* w) { return w ? w->err.c_str() : "wav: cannot open"; }

int frd_sha256_file(const char* path, uint8_t out[32]) {
    std::string bytes;
    if (!path || !ResMan::read_file(path, &bytes)) return FRD_E_BAD_MESSAGE;
    Sha h = Sha256::digest(bytes);
    std::memcpy(out, h.data(), 32);
    return FRB_OK;
}
int64_t frd_adjlist_json(frd_dispatch* d, char* out, uint64_t cap) {
    if (!d) return FRD_E_BAD_MESSAGE;
    std::string s = d->routegraph.to_adjlist().to_json().dump();
    if (out && cap) {
        size_t n = std::min<size_t>(s.size(), cap - 1);
        std::memcpy(out, s.data(), n);
        out[n] = 0;
    }
    return (int64_t)s.size();
}

}  // extern "C"
