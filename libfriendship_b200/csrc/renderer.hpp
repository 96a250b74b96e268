// renderer.hpp — B200Renderer: the host object behind the C ABI (include/friendship_b200.h).
// Mirrors the reference's RefRenderer state (reference src/render/reference.rs:20-29): graph mirror, full
// external-input history, playhead — with the history and every sample buffer resident in HBM.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "flatten.hpp"
#include "graph.hpp"
#include "interp.cuh"
#include "jit.hpp"
#include "osc.cuh"
#include "scan.cuh"
#include "schedule.hpp"

namespace frb {

struct InputSlot {
    // values for absolute times [base, end); zeros before base; reads at >= end are 0 (reference.rs:90-96)
    uint64_t base = 0;       // multiple of 4
    uint64_t end = 0;        // == the reference's slot vector length
    float* d_data = nullptr;
    uint64_t cap = 0;        // floats allocated
};

class Renderer {
public:
    explicit Renderer(const frb_config& cfg);
    ~Renderer();
    Renderer(const Renderer&) = delete;

    // definitions
    void define_effect(uint64_t key, const frb_node* nodes, uint32_t n_nodes, const frb_edge* edges, uint32_t n_edges);
    void define_oscbank(uint64_t key, const frb_oscbank_desc* d);
    void define_directform(uint64_t key, const frb_directform_desc* d);
    void define_fbdelay(uint64_t key, const frb_fbdelay_desc* d);
    // GraphWatcher
    void add_node(uint32_t handle, uint32_t kind, uint64_t key);
    void del_node(uint32_t handle);
    void add_edge(const frb_edge& e);
    void del_edge(const frb_edge& e);
    // Renderer::fill_buffer
    void fill(float* out, bool out_on_device, uint32_t n_slots, uint64_t n_times, uint64_t idx,
              const float* in_data, bool in_on_device, const uint64_t* in_row_offsets, uint32_t n_in_rows);
    // N4: pipelined block render with pinned double-buffered staging both ways (include/friendship_b200.h)
    void render_stream(uint32_t n_slots, uint64_t idx, uint64_t n_total, uint64_t block, uint32_t n_in_rows,
                       frb_source_fn source, frb_sink_fn sink, void* user);
    void sync();
    // Voice sharding (multi.cu): this renderer is rank `rank` of `world` devices.  Bank definitions keep only the voices
    // the rank owns (fixed for the renderer's life); `flatten_sharded` says whether the next schedules evaluate the
    // graph restricted to those voices (linear mixes) or whole (graphs that use no bank lane: rank 0 alone renders).
    void set_shard(uint32_t rank, uint32_t world) { shard_rank_ = rank; shard_world_ = world ? world : 1; dirty_ = true; }
    void set_flatten_sharded(bool on) { if (on != flatten_sharded_) { flatten_sharded_ = on; dirty_ = true; } }
    void invalidate() { cache_valid_ = false; head_ = ~0ull; }        // the next call is a seek whatever its idx
    int device() const { return device_; }
    void sum_rows(float* d_out, const float* d_rows, uint32_t n_rows, uint64_t row_stride, uint64_t n);

    const Schedule& schedule(uint32_t n_slots);   // (re)builds if needed
    // what rank `rank` of `world` devices would flatten this graph to (host only: nothing is uploaded; for tests and tools)
    Schedule schedule_for_shard(uint32_t n_slots, uint32_t rank, uint32_t world) const;
    cudaStream_t stream() const { return stream_; }
    void use_device() const;                      // makes the renderer's device current (throws when planning only)

    std::string last_error;
    std::string last_jit_error;
    frb_stats stats{};
    frb_timing timing{};
    bool profiling = false;

private:
    GraphNode make_node(uint32_t kind, uint64_t key) const;
    void require_device() const;
    void ensure_schedule(uint32_t n_slots);
    void upload_schedule();
    void ensure_rings(uint64_t t_end);
    void ingest_inputs(uint32_t n_slots, uint64_t n_times, uint64_t idx, const float* in_data, bool in_on_device,
                       const uint64_t* offs, uint32_t n_rows);
    void validate_inputs(uint32_t n_slots, uint64_t n_times, uint64_t idx, const uint64_t* offs, uint32_t n_rows) const;
    void materialise_slot(size_t r);
    void grow_slot(InputSlot& s, uint64_t need_end);
    void run_range(uint64_t lo, uint64_t hi, float* d_out, uint64_t t0, uint64_t t1, uint64_t out_stride);
    void poll_stage_jit(size_t sg, uint64_t n_groups);
    void free_device_schedule();
    FlattenEnv flatten_env(uint32_t rank, uint32_t world, bool compact_banks) const;

    frb_config cfg_;
    int device_ = -1;
    bool host_only_ = false;
    int sm_count_ = 148;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};

    Graph graph_;
    std::map<uint64_t, std::shared_ptr<Graph>> effect_defs_;
    std::map<uint64_t, std::shared_ptr<OscBankDev>> osc_defs_;
    std::map<uint64_t, std::shared_ptr<DirectFormDev>> df_defs_;
    std::map<uint64_t, std::shared_ptr<FbDelayDev>> fb_defs_;
    // planning-only handles keep just the shape of extension definitions: (kind, key) -> (lanes, max delay)
    std::map<std::pair<uint32_t, uint64_t>, std::pair<uint32_t, uint64_t>> meta_lanes_;

    bool dirty_ = true;
    uint32_t shard_rank_ = 0, shard_world_ = 1;
    bool flatten_sharded_ = true;
    Schedule sched_;
    uint32_t sched_slots_ = ~0u;

    // device-resident schedule
    std::vector<uint32_t*> d_programs_;         // per stage
    // state: 0 untried, 1 ready, 2 failed, 3 compiling in the background (the stage keeps being interpreted meanwhile)
    // code_instrs: jit_code_instructions of the stage, computed when the stage first qualifies (~0 = not yet)
    // job: the compile running beside the render loop (jit.hpp; dropped, never waited for, when the schedule goes away)
    struct StageJit { JitKernel* k = nullptr; int state = 0; uint64_t uses = 0; uint64_t code_instrs = ~0ull;
                      std::shared_ptr<JitJob> job; std::vector<uint32_t> table; unsigned groups_per_thread = 1; };
    static constexpr uint64_t JIT_MAX_CODE = FRB_JIT_MAX_CODE;        // above: interpreted for good (NVRTC needs minutes)
    static constexpr uint64_t JIT_MAX_SYNC_CODE = FRB_JIT_MAX_SYNC_CODE;   // above: compiled only beside the render loop
    std::vector<StageJit> stage_jit_;
    std::vector<uint32_t*> d_ext_in_bufs_;      // per extension instance: input ring ids per lane
    // Fused DirectForm -> FbDelay chains (scan.cu dfcomb_kernel): chain_of_[fb instance] = df instance whose lanes feed it
    // one to one and nothing else; chained_[df instance] = 1 (not launched, its output rings are not allocated)
    std::vector<int32_t> chain_of_;
    std::vector<uint8_t> chained_;
    std::vector<std::shared_ptr<ChainStateDev>> chain_state_;   // per fb instance
    // Exciter fusion: a fused chain whose biquad lanes each read their own voice of a one-partial oscillator bank, when
    // nothing else reads that bank, evaluates the oscillator inside the chain kernel (osc_one.cuh): exc_of_[fb instance] =
    // the bank's instance, d_exc_voice_[fb instance] = voice per lane; exc_fused_[bank instance] = 1 (not launched, no rings)
    std::vector<int32_t> exc_of_;
    std::vector<uint8_t> exc_fused_;
    std::vector<uint32_t*> d_exc_voice_;
    std::vector<BufferDesc> h_bufdesc_;
    struct RingGroup { uint32_t first, count; float* data; uint64_t cap; };   // consecutive buffers in one allocation
    std::vector<RingGroup> ring_groups_;
    BufferDesc* d_bufdesc_ = nullptr;
    size_t d_bufdesc_cap_ = 0;
    bool bufdesc_dirty_ = true;
    uint64_t chunk_ = 1ull << 16;
    bool align_split_ = true;                                      // run_range: unaligned heads rendered on their own

    // external-input history
    std::vector<InputSlot> inputs_;
    uint64_t n_slot_vectors_ = 0;                                  // reference's self.inputs.len()
    std::vector<std::pair<uint64_t, uint64_t>> epochs_;            // (slot upper bound, base) for never-fed slots
    uint64_t head_ = 0;                                            // reference.rs:26-28
    InputDesc* d_indesc_ = nullptr;
    size_t d_indesc_cap_ = 0;
    uint32_t n_indesc_ = 0;                                        // entries of d_indesc_ in use
    static constexpr uint64_t kMinInputSlotCap = 1u << 16;         // FlattenEnv::input_slot_cap is at least this
    uint64_t sched_input_cap_ = 0;                                 // ... and this is what the current schedule was flattened with
    float* d_in_stage_ = nullptr;
    size_t d_in_stage_cap_ = 0;
    void* d_ingest_ = nullptr;                  // row descriptors of a batched ingest (renderer.cu ingest_rows_kernel)
    size_t d_ingest_cap_ = 0;

    // output staging
    float* d_out_ = nullptr;
    size_t d_out_cap_ = 0;
    std::vector<InputDesc> h_indesc_;                              // what d_indesc_ currently holds
    // Pinned staging for short calls through host buffers (a pageable cudaMemcpyAsync synchronises the stream before it
    // copies: three of them per call made three waits out of one): the descriptor table (two buffers, an event each),
    // the input rows and the output of calls up to kPinBytes.
    static constexpr size_t kPinBytes = 256 * 1024;
    InputDesc* h_pin_indesc_[2] = {nullptr, nullptr};
    size_t h_pin_indesc_cap_ = 0;
    cudaEvent_t ev_indesc_[2] = {nullptr, nullptr};
    unsigned pin_indesc_next_ = 0;
    float* h_pin_in_ = nullptr;
    float* h_pin_out_ = nullptr;

    // streaming render (N4): two blocks in flight
    struct StreamStage { float* d_out = nullptr; float* h_out = nullptr; float* h_in = nullptr;
                         size_t out_cap = 0, in_cap = 0; cudaEvent_t rendered = nullptr, copied = nullptr; };
    StreamStage sstage_[2];
    cudaStream_t copy_stream_ = nullptr;

    // cached rings are valid for a fill that starts exactly at cache_head_
    bool cache_valid_ = false;
    uint64_t cache_head_ = 0;
};

}  // namespace frb
