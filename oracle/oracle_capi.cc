// oracle/oracle_capi.cc — TEST INFRASTRUCTURE ONLY (see ref_renderer.hpp header).
// C shim so tests/ and bench.py's cpu_baseline / --impl reference legs can drive the CPU restatement of
// RefRenderer through ctypes with the same argument shapes as include/friendship_b200.h (orc_* mirrors frb_*).
#include "ref_renderer.hpp"
#include "../include/friendship_b200.h"

#include <chrono>
#include <thread>

using namespace oracle;

struct orc_renderer {
    RefRenderer r;
    std::string err;
};

template <typename F>
static int guarded(orc_renderer* h, F f) {
    try { f(); return 0; }
    catch (const Panic& p) { h->err = p.what(); return p.code; }
    catch (const std::exception& e) { h->err = e.what(); return -6; }
}

extern "C" {

orc_renderer* orc_create(void) { return new orc_renderer(); }
void orc_destroy(orc_renderer* h) { delete h; }
const char* orc_last_error(const orc_renderer* h) { return h ? h->err.c_str() : ""; }
/* mode 0 = fp64 extension oracle, 1 = reference-style f32 */
int orc_set_ext_mode(orc_renderer* h, int mode) { h->r.ext_mode = mode ? ExtMode::F32_REFSTYLE : ExtMode::FP64; return 0; }

int orc_set_sparkle_delay(orc_renderer* h, int on) { h->r.sparkle_delay = on != 0; return 0; }
int orc_set_sparkle_min(orc_renderer* h, int on) { h->r.sparkle_min = on != 0; return 0; }

int orc_define_effect(orc_renderer* h, uint64_t key, const frb_node* nodes, uint32_t n_nodes,
                      const frb_edge* edges, uint32_t n_edges) {
    return guarded(h, [&] {
        std::vector<std::pair<uint32_t, std::pair<uint32_t, uint64_t>>> nds;
        for (uint32_t i = 0; i < n_nodes; i++) nds.push_back({nodes[i].handle, {nodes[i].kind, nodes[i].key}});
        std::vector<Edge> es;
        for (uint32_t i = 0; i < n_edges; i++) es.push_back({edges[i].from, edges[i].to, edges[i].from_slot, edges[i].to_slot});
        h->r.define_effect(key, nds, es);
    });
}
int orc_define_oscbank(orc_renderer* h, uint64_t key, const frb_oscbank_desc* d) {
    return guarded(h, [&] {
        auto b = std::make_shared<OscBankDef>();
        b->sample_rate = d->sample_rate;
        b->voice_offsets.assign(d->voice_offsets, d->voice_offsets + d->n_voices + 1);
        b->freq_hz.assign(d->freq_hz, d->freq_hz + d->n_partials);
        b->amp.assign(d->amp, d->amp + d->n_partials);
        b->phase.assign(d->phase, d->phase + d->n_partials);
        b->attack.assign(d->attack, d->attack + d->n_partials);
        b->tau.assign(d->tau, d->tau + d->n_partials);
        h->r.osc_defs[key] = b;
    });
}
int orc_define_directform(orc_renderer* h, uint64_t key, const frb_directform_desc* d) {
    return guarded(h, [&] {
        auto b = std::make_shared<DirectFormDef>();
        b->b0.assign(d->b0, d->b0 + d->n_lanes); b->b1.assign(d->b1, d->b1 + d->n_lanes);
        b->b2.assign(d->b2, d->b2 + d->n_lanes); b->a1.assign(d->a1, d->a1 + d->n_lanes);
        b->a2.assign(d->a2, d->a2 + d->n_lanes);
        h->r.df_defs[key] = b;
    });
}
int orc_define_fbdelay(orc_renderer* h, uint64_t key, const frb_fbdelay_desc* d) {
    return guarded(h, [&] {
        auto b = std::make_shared<FbDelayDef>();
        b->delay.assign(d->delay, d->delay + d->n_lanes);
        b->gain.assign(d->gain, d->gain + d->n_lanes);
        for (auto D : b->delay) if (D < 1) throw Panic(-6, "fbdelay: delay must be >= 1");
        h->r.fb_defs[key] = b;
    });
}
int orc_add_node(orc_renderer* h, uint32_t handle, uint32_t kind, uint64_t key) {
    return guarded(h, [&] { h->r.on_add_node(handle, kind, key); });
}
int orc_del_node(orc_renderer* h, uint32_t handle) { return guarded(h, [&] { h->r.on_del_node(handle); }); }
int orc_add_edge(orc_renderer* h, frb_edge e) {
    return guarded(h, [&] { h->r.on_add_edge({e.from, e.to, e.from_slot, e.to_slot}); });
}
int orc_del_edge(orc_renderer* h, frb_edge e) {
    return guarded(h, [&] { h->r.on_del_edge({e.from, e.to, e.from_slot, e.to_slot}); });
}
int orc_fill_buffer(orc_renderer* h, float* out, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                    const float* in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows) {
    return guarded(h, [&] {
        std::vector<std::vector<float>> rows(n_in_rows);
        for (uint32_t r = 0; r < n_in_rows; r++)
            rows[r].assign(in_data + in_row_offsets[r], in_data + in_row_offsets[r + 1]);
        h->r.fill_buffer(out, n_slots, n_times, idx, rows);
    });
}

/* CPU-baseline helper: evaluate output slots [0, n_slots) of the current graph over [idx, idx+n_times) with
 * `n_threads` host threads splitting the time range (the pull evaluator is a pure function of time, so the
 * split is exact for feed-forward graphs; recurrence nodes keep a per-thread memo and recompute from 0).
 * Input history must already be present (call orc_fill_buffer with n_times = 0 rows first if needed).
 * Returns seconds of wall time in *seconds_out. */
int orc_fill_buffer_mt(orc_renderer* h, float* out, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                       uint32_t n_threads, double* seconds_out) {
    return guarded(h, [&] {
        h->r.nodes.set_ext_mode(h->r.ext_mode);
        h->r.nodes.clear_memo();
        auto get_input = [h](uint64_t t, uint32_t s) -> float {
            return s < h->r.inputs.size() ? h->r.inputs[s].at(t) : 0.0f;
        };
        if (n_threads < 1) n_threads = 1;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        std::vector<std::string> errs(n_threads);
        for (uint32_t w = 0; w < n_threads; w++) {
            th.emplace_back([&, w] {
                uint64_t lo = n_times * w / n_threads, hi = n_times * (w + 1) / n_threads;
                try {
                    for (uint32_t slot = 0; slot < n_slots; slot++)
                        for (uint64_t t = lo; t < hi; t++)
                            out[(uint64_t)slot * n_times + t] = h->r.nodes.get_output(idx + t, slot, get_input);
                } catch (const std::exception& e) { errs[w] = e.what(); }
            });
        }
        for (auto& t : th) t.join();
        auto t1 = std::chrono::steady_clock::now();
        for (auto& e : errs) if (!e.empty()) throw Panic(-6, e);
        if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    });
}

}  // extern "C"
