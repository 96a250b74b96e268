// multi.cu — N B200s behind one renderer handle (frb_config::n_devices).
//
// The reference's caller owns one renderer inside `Dispatch` and calls fill_buffer from one thread (reference
// src/dispatch.rs:99-106, :147-153).  Here that one call fans out: the oscillator-bank voices are sharded round-robin
// over the devices (voice v on device v mod N, BASELINE.json north_star), every device renders the graph restricted to
// its voices — the other voices are zero signals and fold away at flatten time, so a device's schedule is exactly the
// sub-graph of its voices — and the exchange is K5: each device's stage kernel stores its mix block
// straight into its row of a slab in device 0's HBM over NVLink (peer access, one process: no IPC handles), device 0's
// stream waits for the other devices' events and sums the rows in device order (deterministic left fold).  No
// collective library on this path.
//
// What may be sharded is checked, not assumed: the sum of the devices' outputs equals the whole graph's output only
// where the outputs are linear in the bank lanes (lane_use_of_outputs, flatten.cc).  A graph that uses no bank lane
// renders on device 0 alone, bit-exact as ever; a graph that puts a lane through Minimum / Modulo / a product of lanes
// is refused with FRB_E_UNSUPPORTED (use a one-device renderer for it).
#include "multi.hpp"

#include <algorithm>
#include <cstring>

namespace frb {

#define CU(expr)                                                                                    \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw Error{FRB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)};            \
    } while (0)

MultiRenderer::MultiRenderer(const frb_config& cfg) {
    const uint32_t n = cfg.n_devices;
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0) throw Error{FRB_E_NO_DEVICE, std::string("no CUDA device available: ") + cudaGetErrorString(e)};
    if (cfg.device < 0 || (uint64_t)cfg.device + n > (uint64_t)have)
        throw Error{FRB_E_NO_DEVICE, "n_devices = " + std::to_string(n) + " from device " + std::to_string(cfg.device) + ": only " + std::to_string(have) + " CUDA devices"};
    const int dev0 = cfg.device;
    for (uint32_t i = 1; i < n; i++) {                       // every device stores into device 0's slab and may read its input rows
        int ok = 0;
        CU(cudaDeviceCanAccessPeer(&ok, dev0 + (int)i, dev0));
        if (!ok) throw Error{FRB_E_UNSUPPORTED, "device " + std::to_string(dev0 + i) + " has no peer access to device " + std::to_string(dev0)};
        CU(cudaSetDevice(dev0 + (int)i));
        e = cudaDeviceEnablePeerAccess(dev0, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
        else if (e != cudaSuccess) throw Error{FRB_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)};
    }
    frb_config pc = cfg;
    pc.device = -1;
    pc.n_devices = 0;
    plan_ = std::make_unique<Renderer>(pc);
    for (uint32_t i = 0; i < n; i++) {
        frb_config kc = cfg;
        kc.device = dev0 + (int)i;
        kc.n_devices = 0;
        kids_.push_back(std::make_unique<Renderer>(kc));
        kids_.back()->set_shard(i, n);
    }
    done_.assign(n, nullptr);
    for (uint32_t i = 0; i < n; i++) {
        CU(cudaSetDevice(dev0 + (int)i));
        CU(cudaEventCreateWithFlags(&done_[i], cudaEventDisableTiming));
    }
    errors_.resize(n);
    for (uint32_t i = 0; i < n; i++) threads_.emplace_back([this, i] { worker(i); });
}

MultiRenderer::~MultiRenderer() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        quit_ = true;
    }
    cv_go_.notify_all();
    for (auto& t : threads_) t.join();
    for (size_t i = 0; i < kids_.size(); i++) {
        cudaSetDevice(kids_[i]->device());
        cudaStreamSynchronize(kids_[i]->stream());
        if (done_[i]) cudaEventDestroy(done_[i]);
    }
    if (d_slab_) { cudaSetDevice(kids_[0]->device()); cudaFree(d_slab_); }
}

void MultiRenderer::worker(size_t i) {
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(size_t)>* job;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_go_.wait(lk, [&] { return quit_ || job_seq_ != seen; });
            if (quit_) return;
            seen = job_seq_;
            job = job_;
        }
        try {
            cudaSetDevice(kids_[i]->device());
            (*job)(i);
        } catch (const Error& e) {
            errors_[i] = std::make_unique<Error>(e);
        } catch (const std::exception& e) {
            errors_[i] = std::make_unique<Error>(Error{FRB_E_INVALID, e.what()});
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            pending_--;
        }
        cv_done_.notify_all();
    }
}

void MultiRenderer::run_all(const std::function<void(size_t)>& fn) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &fn;
        pending_ = kids_.size();
        job_seq_++;
    }
    cv_go_.notify_all();
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
    std::unique_ptr<Error> first;
    for (auto& e : errors_) {
        if (e && !first) first = std::move(e);
        e.reset();
    }
    if (first) {
        for (auto& k : kids_) k->invalidate();               // some devices advanced, some did not: the next call starts over
        throw *first;
    }
}

// ---- definitions and graph edits go to every device and to the planning mirror ----
void MultiRenderer::define_effect(uint64_t key, const frb_node* nodes, uint32_t n_nodes, const frb_edge* edges, uint32_t n_edges) {
    plan_->define_effect(key, nodes, n_nodes, edges, n_edges);
    for (auto& k : kids_) k->define_effect(key, nodes, n_nodes, edges, n_edges);
}
void MultiRenderer::define_oscbank(uint64_t key, const frb_oscbank_desc* d) {
    plan_->define_oscbank(key, d);                           // the whole bank's shape (lanes); a new lane count re-plans
    auto it = bank_voices_.find(key);
    if (it == bank_voices_.end() || it->second != d->n_voices) graph_dirty_ = true;
    bank_voices_[key] = d->n_voices;
    run_all([&](size_t i) { kids_[i]->define_oscbank(key, d); });   // each device: the voices it owns, in parallel
}
void MultiRenderer::define_directform(uint64_t key, const frb_directform_desc* d) {
    plan_->define_directform(key, d);
    run_all([&](size_t i) { kids_[i]->define_directform(key, d); });
    graph_dirty_ = true;
}
void MultiRenderer::define_fbdelay(uint64_t key, const frb_fbdelay_desc* d) {
    plan_->define_fbdelay(key, d);
    run_all([&](size_t i) { kids_[i]->define_fbdelay(key, d); });
    graph_dirty_ = true;
}
void MultiRenderer::add_node(uint32_t handle, uint32_t kind, uint64_t key) {
    plan_->add_node(handle, kind, key);
    for (auto& k : kids_) k->add_node(handle, kind, key);
    graph_dirty_ = true;
}
void MultiRenderer::del_node(uint32_t handle) {
    plan_->del_node(handle);
    for (auto& k : kids_) k->del_node(handle);
    graph_dirty_ = true;
}
void MultiRenderer::add_edge(const frb_edge& e) {
    plan_->add_edge(e);
    for (auto& k : kids_) k->add_edge(e);
    graph_dirty_ = true;
}
void MultiRenderer::del_edge(const frb_edge& e) {
    plan_->del_edge(e);
    for (auto& k : kids_) k->del_edge(e);
    graph_dirty_ = true;
}

void MultiRenderer::plan(uint32_t n_slots) {
    if (!graph_dirty_ && plan_slots_ == n_slots) return;
    const Schedule& s = plan_->schedule(n_slots);            // throws on a malformed graph, before any device is touched
    mode_ = lane_use_of_outputs(s);
    if (mode_ == LANES_OTHER)
        throw Error{FRB_E_UNSUPPORTED,
                    "n_devices > 1 shards the oscillator-bank voices over the devices and adds the devices' outputs: every output "
                    "slot must be linear in the bank lanes (sums, gains, delays, linear filters), or use no lane at all"};
    for (auto& k : kids_) k->set_flatten_sharded(mode_ == LANES_LINEAR);
    plan_slots_ = n_slots;
    graph_dirty_ = false;
}

void MultiRenderer::fill(float* out, bool out_on_device, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                         const float* in_data, bool in_on_device, const uint64_t* offs, uint32_t n_rows) {
    plan(n_slots);
    const size_t n_dev = kids_.size();
    const uint64_t n_out = (uint64_t)n_slots * n_times;
    if (n_out == 0) {
        run_all([&](size_t i) { kids_[i]->fill(nullptr, true, n_slots, n_times, idx, in_data, in_on_device, offs, n_rows); });
        return;
    }
    const uint64_t row = (n_out + 3) & ~3ull;                // rows stay 16-byte aligned
    const size_t need = (size_t)row * (n_dev + 1);
    if (slab_cap_ < need) {
        sync();                                              // leaves the LAST device current
        CU(cudaSetDevice(kids_[0]->device()));               // the slab lives in the first device's HBM
        if (d_slab_) CU(cudaFree(d_slab_));
        d_slab_ = nullptr; slab_cap_ = 0;
        CU(cudaMalloc(&d_slab_, need * sizeof(float)));
        slab_cap_ = need;
    }
    // each device renders the sub-graph of its voices; its stage kernel's output stores ARE the transfer (row i of the slab).
    // A graph that uses no bank lane has nothing to shard: every device renders all of it (so that every device's input
    // history and playhead stay those of a one-device renderer) and device 0's block is the result, bit-exact as ever.
    const bool shard = mode_ == LANES_LINEAR;
    float* d_mix = out_on_device ? out : d_slab_ + n_dev * row;
    run_all([&](size_t i) {
        float* dst = (!shard && i == 0) ? d_mix : d_slab_ + i * row;
        kids_[i]->fill(dst, true, n_slots, n_times, idx, in_data, in_on_device, offs, n_rows);
        CU(cudaEventRecord(done_[i], kids_[i]->stream()));
    });
    CU(cudaSetDevice(kids_[0]->device()));
    cudaStream_t s0 = kids_[0]->stream();
    for (size_t i = 1; i < n_dev; i++) CU(cudaStreamWaitEvent(s0, done_[i], 0));
    if (shard) kids_[0]->sum_rows(d_mix, d_slab_, (uint32_t)n_dev, row, n_out);
    if (!out_on_device) {
        CU(cudaMemcpyAsync(out, d_mix, n_out * sizeof(float), cudaMemcpyDeviceToHost, s0));
        CU(cudaStreamSynchronize(s0));
    }
    // the slab rows are rewritten by the next call: the other devices' next kernels must not start before the sum read
    // them.  One event on device 0's stream, awaited by every other stream (device-side: the host does not wait).
    CU(cudaEventRecord(done_[0], s0));
    for (size_t i = 1; i < n_dev; i++) {
        CU(cudaSetDevice(kids_[i]->device()));
        CU(cudaStreamWaitEvent(kids_[i]->stream(), done_[0], 0));
    }
    CU(cudaSetDevice(kids_[0]->device()));
}

// N4 on several devices: consecutive fills through host staging (the blocks' render, exchange and copy do not overlap here;
// the one-device renderer pipelines them)
void MultiRenderer::render_stream(uint32_t n_slots, uint64_t idx, uint64_t n_total, uint64_t block, uint32_t n_in_rows,
                                  frb_source_fn source, frb_sink_fn sink, void* user) {
    if (!sink) throw Error{FRB_E_INVALID, "render_stream: sink is NULL"};
    if (n_in_rows && !source) throw Error{FRB_E_INVALID, "render_stream: inputs without a source"};
    if (block == 0) throw Error{FRB_E_INVALID, "render_stream: block must be positive"};
    if (idx + n_total < idx) throw Error{FRB_E_INVALID, "idx + n_total overflows"};
    block = std::min(block, std::max<uint64_t>(n_total, 1));
    std::vector<float> outb((size_t)n_slots * block), inb((size_t)n_in_rows * block);
    std::vector<uint64_t> offs(n_in_rows + 1, 0);
    const uint64_t n_blocks = n_total ? (n_total + block - 1) / block : 1;
    uint64_t t = idx;
    for (uint64_t k = 0; k < n_blocks; k++) {
        const uint64_t n = std::min(block, idx + n_total - t);
        if (n_in_rows) {
            if (source(user, inb.data(), n_in_rows, n, t) != 0) throw Error{FRB_E_INVALID, "render_stream: source failed"};
            for (uint32_t r = 0; r <= n_in_rows; r++) offs[r] = (uint64_t)r * n;
        }
        fill(outb.data(), false, n_slots, n, t, n_in_rows ? inb.data() : nullptr, false, offs.data(), n_in_rows);
        if (sink(user, outb.data(), n_slots, n, t) != 0) throw Error{FRB_E_INVALID, "render_stream: sink failed"};
        t += n;
    }
}

void MultiRenderer::sync() {
    for (auto& k : kids_) k->sync();
}

frb_stats MultiRenderer::get_stats() const {
    frb_stats s{};
    for (auto& k : kids_) {
        const frb_stats& a = k->stats;
        s.kernel_launches += a.kernel_launches; s.h2d_bytes += a.h2d_bytes; s.d2h_bytes += a.d2h_bytes;
        s.schedule_builds += a.schedule_builds; s.osc_launches += a.osc_launches; s.interp_launches += a.interp_launches;
        s.scan_launches += a.scan_launches; s.jit_launches += a.jit_launches; s.chain_launches += a.chain_launches;
        s.osc_tensor_launches += a.osc_tensor_launches;
    }
    return s;
}

frb_timing MultiRenderer::get_timing() const {
    frb_timing t{};
    for (auto& k : kids_) {
        t.osc_ms = std::max(t.osc_ms, k->timing.osc_ms); t.interp_ms = std::max(t.interp_ms, k->timing.interp_ms);
        t.scan_ms = std::max(t.scan_ms, k->timing.scan_ms); t.total_ms = std::max(t.total_ms, k->timing.total_ms);
    }
    return t;
}

}  // namespace frb
