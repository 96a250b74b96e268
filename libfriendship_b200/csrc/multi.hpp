// multi.hpp — one renderer object, several B200s (see multi.cu).
//
// The reference's caller owns ONE renderer value inside `Dispatch` (reference src/dispatch.rs:99-106, render arm
// :147-153) and a Rust `Dispatch<B200Renderer, C>` keeps doing so: frb_config::n_devices > 1 puts the voice sharding and
// the exchange behind the same C ABI — one host thread calls frb_fill_buffer, the library drives the devices.
#pragma once
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "renderer.hpp"

namespace frb {

class MultiRenderer {
public:
    explicit MultiRenderer(const frb_config& cfg);
    ~MultiRenderer();
    MultiRenderer(const MultiRenderer&) = delete;

    // same surface as Renderer (capi.cu calls whichever the handle holds)
    void define_effect(uint64_t key, const frb_node* nodes, uint32_t n_nodes, const frb_edge* edges, uint32_t n_edges);
    void define_oscbank(uint64_t key, const frb_oscbank_desc* d);
    void define_directform(uint64_t key, const frb_directform_desc* d);
    void define_fbdelay(uint64_t key, const frb_fbdelay_desc* d);
    void add_node(uint32_t handle, uint32_t kind, uint64_t key);
    void del_node(uint32_t handle);
    void add_edge(const frb_edge& e);
    void del_edge(const frb_edge& e);
    void fill(float* out, bool out_on_device, uint32_t n_slots, uint64_t n_times, uint64_t idx,
              const float* in_data, bool in_on_device, const uint64_t* in_row_offsets, uint32_t n_in_rows);
    void render_stream(uint32_t n_slots, uint64_t idx, uint64_t n_total, uint64_t block, uint32_t n_in_rows,
                       frb_source_fn source, frb_sink_fn sink, void* user);
    void sync();
    void sum_rows(float* d_out, const float* d_rows, uint32_t n_rows, uint64_t row_stride, uint64_t n) { kids_[0]->sum_rows(d_out, d_rows, n_rows, row_stride, n); }
    const Schedule& schedule(uint32_t n_slots) { return plan_->schedule(n_slots); }   // the whole graph's (unsharded) schedule
    Schedule schedule_for_shard(uint32_t n_slots, uint32_t rank, uint32_t world) const { return plan_->schedule_for_shard(n_slots, rank, world); }
    cudaStream_t stream() const { return kids_[0]->stream(); }
    void use_device() const { kids_[0]->use_device(); }
    void set_profiling(bool on) { for (auto& k : kids_) k->profiling = on; }
    frb_stats get_stats() const;
    frb_timing get_timing() const;

    std::string last_error;

private:
    void run_all(const std::function<void(size_t)>& fn);      // fn(i) on worker i (device i current), all in parallel; rethrows
    void worker(size_t i);
    void plan(uint32_t n_slots);

    std::vector<std::unique_ptr<Renderer>> kids_;             // one per device; kids_[0] owns the exchange slab
    std::unique_ptr<Renderer> plan_;                          // host-only mirror of the whole graph: lane analysis, dumps
    std::map<uint64_t, uint32_t> bank_voices_;
    bool graph_dirty_ = true;
    uint32_t plan_slots_ = ~0u;
    LaneUse mode_ = LANES_UNUSED;

    float* d_slab_ = nullptr;                                 // device 0: [n_devices x n_slots x n_times] + the summed block
    size_t slab_cap_ = 0;
    std::vector<cudaEvent_t> done_;                           // per device: its block is in the slab

    // one worker thread per device
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_go_, cv_done_;
    const std::function<void(size_t)>* job_ = nullptr;
    uint64_t job_seq_ = 0;
    size_t pending_ = 0;
    bool quit_ = false;
    std::vector<std::unique_ptr<Error>> errors_;
};

}  // namespace frb
