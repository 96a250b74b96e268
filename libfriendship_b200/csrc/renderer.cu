// renderer.cu — B200Renderer host logic: graph mirror, external-input history in HBM, schedule upload,
// block-wise rendering.  Follows the contract of Renderer::fill_buffer (reference src/render/renderer.rs:6-17)
// and the bookkeeping of RefRenderer::fill_buffer (reference src/render/reference.rs:46-86).
#include "renderer.hpp"
#include "osc_one.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace frb {

#define CU(expr)                                                                                    \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw Error{FRB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)};            \
    } while (0)

static uint64_t pow2_ceil(uint64_t x) {
    uint64_t p = 8;
    while (p < x) p <<= 1;
    return p;
}

// pads data[start .. start+count) with data[last_pos], or 0 when has_last == 0 (reference.rs:72-73)
__global__ void pad_kernel(float* data, unsigned long long start, unsigned long long count,
                           unsigned long long last_pos, int has_last) {
    const float v = has_last ? data[last_pos] : 0.0f;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x)
        data[start + i] = v;
}

// Ingest of many input rows in one launch (row r: n floats from src to dst): blockIdx.y = row.  A cudaMemcpyAsync per row
// left the copy engine / launch path idle between rows: 64 rows x 16 MB moved at 3.3 TB/s (read + write), this kernel at
// the copy peak.  128-bit accesses where both ends are 16-byte aligned, four of them in flight per thread.
struct IngestRow { float* dst; const float* src; unsigned long long n; };
__global__ void __launch_bounds__(256) ingest_rows_kernel(const IngestRow* __restrict__ rows) {
    const IngestRow r = rows[blockIdx.y];
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthr = (unsigned long long)gridDim.x * blockDim.x;
    if ((((uintptr_t)r.dst | (uintptr_t)r.src) & 15u) == 0) {
        const unsigned long long n4 = r.n / 4;
        const float4* s4 = reinterpret_cast<const float4*>(r.src);
        float4* d4 = reinterpret_cast<float4*>(r.dst);
        unsigned long long i = tid;
        for (; i + 3 * nthr < n4; i += 4 * nthr) {
            const float4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + nthr), c = __ldcs(s4 + i + 2 * nthr), d = __ldcs(s4 + i + 3 * nthr);
            d4[i] = a; d4[i + nthr] = b; d4[i + 2 * nthr] = c; d4[i + 3 * nthr] = d;
        }
        for (; i < n4; i += nthr) d4[i] = __ldcs(s4 + i);
        for (unsigned long long j = 4 * n4 + tid; j < r.n; j += nthr) r.dst[j] = r.src[j];
    } else {
        for (unsigned long long j = tid; j < r.n; j += nthr) r.dst[j] = r.src[j];
    }
}

// K5: rank-ordered sum of the per-rank mix rows (deterministic left fold, like every other sum on the path)
__global__ void sum_rows_kernel(float* __restrict__ out, const float* __restrict__ rows, unsigned n_rows,
                                unsigned long long stride, unsigned long long n) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        float acc = rows[i];
        for (unsigned r = 1; r < n_rows; r++) acc = __fadd_rn(acc, rows[r * stride + i]);
        out[i] = acc;
    }
}

void Renderer::sum_rows(float* d_out, const float* d_rows, uint32_t n_rows, uint64_t row_stride, uint64_t n) {
    require_device();
    CU(cudaSetDevice(device_));
    if (!n || !n_rows) return;
    unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count_ * 8);
    sum_rows_kernel<<<blocks, 256, 0, stream_>>>(d_out, d_rows, n_rows, row_stride, n);
    CU(cudaGetLastError());
    stats.kernel_launches++;
}

Renderer::Renderer(const frb_config& cfg) : cfg_(cfg) {
    if (cfg.device < 0) {
        // planning-only instance: graph mirror + schedule dumps; rendering is refused (there is no CPU fallback)
        host_only_ = true;
        return;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error{FRB_E_NO_DEVICE, std::string("no CUDA device available: ") + cudaGetErrorString(e)};
    if (cfg.device >= n) throw Error{FRB_E_NO_DEVICE, "CUDA device ordinal out of range"};
    device_ = cfg.device;
    CU(cudaSetDevice(device_));
    CU(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_));
    CU(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto& ev : ev_) CU(cudaEventCreate(&ev));
    if (const char* e = getenv("FRB_NO_ALIGN_SPLIT")) align_split_ = !(e[0] == '1');   // measurement knob (tools/align_probe.py)
    CU(interp_init_device());
    CU(osc_init_device());
}

Renderer::~Renderer() {
    if (host_only_) return;
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    free_device_schedule();
    jit_wait_idle();                      // no compile thread outlives the renderers (a sub-second wait at most)
    for (auto& s : inputs_) if (s.d_data) cudaFree(s.d_data);
    if (d_indesc_) cudaFree(d_indesc_);
    if (d_in_stage_) cudaFree(d_in_stage_);
    if (d_ingest_) cudaFree(d_ingest_);
    for (int i = 0; i < 2; i++) { if (h_pin_indesc_[i]) cudaFreeHost(h_pin_indesc_[i]); if (ev_indesc_[i]) cudaEventDestroy(ev_indesc_[i]); }
    if (h_pin_in_) cudaFreeHost(h_pin_in_);
    if (h_pin_out_) cudaFreeHost(h_pin_out_);
    if (d_out_) cudaFree(d_out_);
    if (d_bufdesc_) cudaFree(d_bufdesc_);
    for (auto& st : sstage_) {
        if (st.d_out) cudaFree(st.d_out);
        if (st.h_out) cudaFreeHost(st.h_out);
        if (st.h_in) cudaFreeHost(st.h_in);
        if (st.rendered) cudaEventDestroy(st.rendered);
        if (st.copied) cudaEventDestroy(st.copied);
    }
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    for (auto& ev : ev_) if (ev) cudaEventDestroy(ev);
    if (stream_) cudaStreamDestroy(stream_);
}

void Renderer::require_device() const {
    if (host_only_) throw Error{FRB_E_NO_DEVICE, "renderer was created without a CUDA device (planning only); there is no CPU fallback"};
}

void Renderer::use_device() const {
    require_device();
    CU(cudaSetDevice(device_));
}

// ------------------------------------------------------------------------------------------------ definitions

GraphNode Renderer::make_node(uint32_t kind, uint64_t key) const {
    GraphNode n;
    n.kind = kind;
    n.key = key;
    switch (kind) {
        case FRB_KIND_DELAY: case FRB_KIND_F32CONSTANT: case FRB_KIND_SUM2: case FRB_KIND_MULTIPLY:
        case FRB_KIND_DIVIDE: case FRB_KIND_MODULO: case FRB_KIND_MINIMUM:
            break;
        case FRB_KIND_EFFECT: {
            auto it = effect_defs_.find(key);
            if (it == effect_defs_.end()) throw Error{FRB_E_BAD_HANDLE, "unknown effect definition key"};
            n.body = it->second->deep_copy();   // reference.rs:98-113: private copy per node
            break;
        }
        case FRB_KIND_OSCBANK:
            if (!osc_defs_.count(key) && !meta_lanes_.count({kind, key})) throw Error{FRB_E_BAD_HANDLE, "unknown oscbank key"};
            break;
        case FRB_KIND_DIRECTFORM:
            if (!df_defs_.count(key) && !meta_lanes_.count({kind, key})) throw Error{FRB_E_BAD_HANDLE, "unknown directform key"};
            break;
        case FRB_KIND_FBDELAY:
            if (!fb_defs_.count(key) && !meta_lanes_.count({kind, key})) throw Error{FRB_E_BAD_HANDLE, "unknown fbdelay key"};
            break;
        default: throw Error{FRB_E_INVALID, "unknown node kind"};
    }
    return n;
}

void Renderer::define_effect(uint64_t key, const frb_node* nodes, uint32_t n_nodes, const frb_edge* edges, uint32_t n_edges) {
    auto g = std::make_shared<Graph>();
    for (uint32_t i = 0; i < n_nodes; i++) {
        if (nodes[i].handle == 0) throw Error{FRB_E_INVALID, "node handle 0 is the toplevel"};
        g->nodes[nodes[i].handle] = make_node(nodes[i].kind, nodes[i].key);
    }
    for (uint32_t i = 0; i < n_edges; i++)
        if (!g->add_edge(edges[i])) throw Error{FRB_E_BAD_HANDLE, "effect definition: edge to unknown node"};
    effect_defs_[key] = g;
}

// the bank and comb kernels put voices / lanes on grid.y: refuse what a launch could not hold at definition time,
// not in the middle of a render
static constexpr uint32_t kMaxGridY = 65535;

void Renderer::define_oscbank(uint64_t key, const frb_oscbank_desc* d) {
    if (d->n_voices > kMaxGridY) throw Error{FRB_E_UNSUPPORTED, "oscbank: at most 65,535 voices per bank (use several banks)"};
    if (host_only_) {                                        // planning only: shape
        auto& m = meta_lanes_[{FRB_KIND_OSCBANK, key}];
        if (m.first != d->n_voices) dirty_ = true;
        m = {d->n_voices, 0};
        return;
    }
    require_device();
    CU(cudaSetDevice(device_));
    std::string err;
    std::shared_ptr<OscBankDev> old;
    if (auto it = osc_defs_.find(key); it != osc_defs_.end()) old = it->second;
    const uint32_t old_voices = old ? osc_info(*old).n_voices : 0;
    auto b = osc_create(d, stream_, &err, old, shard_rank_, shard_world_, (cfg_.flags & FRB_FLAG_NO_TENSOR_OSC) == 0);
    if (!b) {
        if (old && !osc_usable(*old)) { osc_defs_.erase(key); dirty_ = true; }   // its allocations went into the failed attempt
        throw Error{FRB_E_INVALID, err};
    }
    stats.h2d_bytes += osc_info(*b).n_partials * (sizeof(double) + 4 * sizeof(float)) + ((uint64_t)osc_info(*b).n_voices + 1) * sizeof(uint64_t);
    stats.kernel_launches += 3;   // rank, fill, setup
    if (old) {
        // re-definition = new parameters for the same node (the per-render input of a synthesis graph)
        if (old_voices != osc_info(*b).n_voices) dirty_ = true;   // the lane count is part of the schedule
        if (osc_one_source(*old, nullptr) != osc_one_source(*b, nullptr)) dirty_ = true;   // ... and so is whether a chain can evaluate it
        cache_valid_ = false;                               // rings hold the old bank's samples
    }
    osc_defs_[key] = b;
}
void Renderer::define_directform(uint64_t key, const frb_directform_desc* d) {
    if (host_only_) { meta_lanes_[{FRB_KIND_DIRECTFORM, key}] = {d->n_lanes, 0}; return; }
    require_device();
    CU(cudaSetDevice(device_));
    std::string err;
    auto b = directform_create(d, stream_, &err);
    if (!b) throw Error{FRB_E_INVALID, err};
    if (df_defs_.count(key)) { CU(cudaStreamSynchronize(stream_)); dirty_ = true; }   // lane count / fusability may change
    df_defs_[key] = b;
}
void Renderer::define_fbdelay(uint64_t key, const frb_fbdelay_desc* d) {
    if (d->n_lanes > kMaxGridY) throw Error{FRB_E_UNSUPPORTED, "fbdelay: at most 65,535 lanes per bank (use several banks)"};
    if (host_only_) {
        uint64_t mx = 0;
        if (d->n_lanes && (!d->delay || !d->gain)) throw Error{FRB_E_INVALID, "fbdelay: null array"};
        for (uint32_t i = 0; i < d->n_lanes; i++) {
            if (d->delay[i] < 1) throw Error{FRB_E_INVALID, "fbdelay: delay must be >= 1"};
            mx = std::max<uint64_t>(mx, d->delay[i]);
        }
        meta_lanes_[{FRB_KIND_FBDELAY, key}] = {d->n_lanes, mx};
        return;
    }
    require_device();
    CU(cudaSetDevice(device_));
    std::string err;
    auto b = fbdelay_create(d, stream_, &err);
    if (!b) throw Error{FRB_E_INVALID, err};
    if (fb_defs_.count(key)) { CU(cudaStreamSynchronize(stream_)); dirty_ = true; }   // lookbacks / fusability may change
    fb_defs_[key] = b;
}

// ------------------------------------------------------------------------------------------------ GraphWatcher

void Renderer::add_node(uint32_t handle, uint32_t kind, uint64_t key) {   // reference.rs:117-120
    if (handle == 0) throw Error{FRB_E_INVALID, "node handle 0 is the toplevel"};
    graph_.nodes[handle] = make_node(kind, key);                          // HashMap::insert replaces
    dirty_ = true;
}
void Renderer::del_node(uint32_t handle) {                                 // reference.rs:121-123
    graph_.nodes.erase(handle);
    dirty_ = true;
}
void Renderer::add_edge(const frb_edge& e) {                               // reference.rs:124-126
    // the reference grows `inbound` to to_slot + 1 entries (reference.rs:150); RouteGraph validation keeps slots small
    // before the renderer is told, a raw C caller may not
    if (e.to_slot > (1u << 24)) throw Error{FRB_E_INVALID, "to_slot out of range"};
    if (!graph_.add_edge(e)) throw Error{FRB_E_BAD_HANDLE, "add_edge: target node does not exist (reference.rs:145)"};
    dirty_ = true;
}
void Renderer::del_edge(const frb_edge& e) {                               // reference.rs:127-136
    if (!graph_.del_edge(e)) throw Error{FRB_E_BAD_HANDLE, "Attempt to delete edge, but it was never created! (reference.rs:131)"};
    dirty_ = true;
}

// ------------------------------------------------------------------------------------------------ schedule

void Renderer::free_device_schedule() {
    for (auto p : d_programs_) if (p) cudaFree(p);
    d_programs_.clear();
    // a compile still running beside the loop is simply dropped: its thread owns the job, the cubin goes to the cache
    for (auto& j : stage_jit_) jit_free(j.k);
    stage_jit_.clear();
    for (auto& g : ring_groups_) if (g.data) cudaFree(g.data);
    ring_groups_.clear();
    h_bufdesc_.clear();
    for (auto p : d_ext_in_bufs_) if (p) cudaFree(p);
    d_ext_in_bufs_.clear();
    for (auto p : d_exc_voice_) if (p) cudaFree(p);
    d_exc_voice_.clear();
}

FlattenEnv Renderer::flatten_env(uint32_t rank, uint32_t world, bool compact_banks) const {
    FlattenEnv env;
    env.ext_lanes = [this, rank, world, compact_banks](uint32_t kind, uint64_t key) -> int64_t {
        int64_t n = -1;
        auto mit = meta_lanes_.find({kind, key});
        if (mit != meta_lanes_.end()) n = (int64_t)mit->second.first;
        else if (kind == FRB_KIND_OSCBANK) { auto it = osc_defs_.find(key); n = it == osc_defs_.end() ? -1 : (int64_t)osc_info(*it->second).n_voices; }
        else if (kind == FRB_KIND_DIRECTFORM) { auto it = df_defs_.find(key); n = it == df_defs_.end() ? -1 : (int64_t)directform_lanes(*it->second); }
        else if (kind == FRB_KIND_FBDELAY) { auto it = fb_defs_.find(key); n = it == fb_defs_.end() ? -1 : (int64_t)fbdelay_lanes(*it->second); }
        // the whole bank's lane count -> the lanes rank `rank` owns (osc_create keeps voices rank, rank + world, ...)
        if (compact_banks && kind == FRB_KIND_OSCBANK && n >= 0 && world > 1) n = n > (int64_t)rank ? (n - rank + world - 1) / world : 0;
        return n;
    };
    env.ext_max_delay = [this](uint64_t key) -> uint64_t {
        auto mit = meta_lanes_.find({FRB_KIND_FBDELAY, key});
        if (mit != meta_lanes_.end()) return mit->second.second;
        auto it = fb_defs_.find(key);
        return it == fb_defs_.end() ? 0 : fbdelay_max_delay(*it->second);
    };
    env.max_regs = 48;
    env.sparkle_delay = (cfg_.flags & FRB_FLAG_SPARKLE_DELAY) != 0;
    env.shard_rank = rank;
    env.shard_world = world ? world : 1;
    env.input_slot_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, std::max<uint64_t>(kMinInputSlotCap, inputs_.size()));
    return env;
}

Schedule Renderer::schedule_for_shard(uint32_t n_slots, uint32_t rank, uint32_t world) const {
    if (world == 0 || rank >= world) throw Error{FRB_E_INVALID, "shard rank out of range"};
    // a renderer that is itself a shard already holds compact banks; any other (planning handles included) holds whole ones
    return flatten(graph_, n_slots, flatten_env(rank, world, shard_world_ <= 1));
}

const Schedule& Renderer::schedule(uint32_t n_slots) {
    if (dirty_ || sched_slots_ != n_slots) {
        // this renderer's banks already hold only the voices it owns (set_shard), so their lane counts are the compact ones
        const bool sharded = shard_world_ > 1 && flatten_sharded_;
        FlattenEnv env = flatten_env(sharded ? shard_rank_ : 0, sharded ? shard_world_ : 1, false);
        Schedule s = flatten(graph_, n_slots, env);   // throws on malformed graphs; state unchanged then
        sched_input_cap_ = env.input_slot_cap;
        if (!host_only_) {
            CU(cudaSetDevice(device_));
            CU(cudaStreamSynchronize(stream_));
            free_device_schedule();
        }
        sched_ = std::move(s);
        sched_slots_ = ~0u;                  // not valid until the upload below succeeded
        dirty_ = true;
        cache_valid_ = false;
        stats.schedule_builds++;
        if (!host_only_) {
            try {
                upload_schedule();
            } catch (...) {
                cudaStreamSynchronize(stream_);
                free_device_schedule();      // nothing half-uploaded is ever launched: the next call rebuilds
                throw;
            }
        }
        sched_slots_ = n_slots;
        dirty_ = false;
    }
    return sched_;
}

void Renderer::ensure_schedule(uint32_t n_slots) { (void)schedule(n_slots); }

void Renderer::upload_schedule() {
    d_programs_.assign(sched_.stages.size(), nullptr);
    stage_jit_ = std::vector<StageJit>(sched_.stages.size());
    for (size_t i = 0; i < sched_.stages.size(); i++) {
        // device layout: [strand offsets, padded to a multiple of 4 words][instructions]
        const Stage& stg = sched_.stages[i];
        std::vector<uint32_t> words(stg.strand_offsets);
        while (words.size() % 4) words.push_back(0);
        for (const Instr& in : stg.program) { words.push_back(in.w0); words.push_back(in.a); words.push_back(in.b); words.push_back(in.aux); }
        size_t bytes = words.size() * sizeof(uint32_t);
        CU(cudaMalloc(&d_programs_[i], bytes));
        CU(cudaMemcpy(d_programs_[i], words.data(), bytes, cudaMemcpyHostToDevice));
        stats.h2d_bytes += bytes;
    }
    h_bufdesc_.assign(sched_.buffers.size(), BufferDesc{nullptr, 0});
    bufdesc_dirty_ = true;
    d_ext_in_bufs_.assign(sched_.ext.size(), nullptr);
    for (size_t i = 0; i < sched_.ext.size(); i++) {
        auto& x = sched_.ext[i];
        if (x.in_bufs.empty()) continue;
        CU(cudaMalloc(&d_ext_in_bufs_[i], x.in_bufs.size() * sizeof(uint32_t)));
        CU(cudaMemcpyAsync(d_ext_in_bufs_[i], x.in_bufs.data(), x.in_bufs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    }
    CU(cudaStreamSynchronize(stream_));   // host vectors above may go away

    // Chain fusion: a FbDelay instance whose lane l reads lane l of a DirectForm instance of the same stage, when nothing
    // else reads that biquad (no other node, no output slot, no other lane), runs with it in one kernel and the
    // biquad's output rings are never allocated.
    chain_of_.assign(sched_.ext.size(), -1);
    chained_.assign(sched_.ext.size(), 0);
    chain_state_.assign(sched_.ext.size(), nullptr);
    if (!(cfg_.flags & FRB_FLAG_NO_CHAIN_FUSION)) {
        const auto& V = sched_.values;
        std::vector<uint32_t> uses(V.size(), 0);
        for (const Value& x : V) {
            if (x.op == V_DELAY || (x.op >= V_SUM2 && x.op <= V_MIN)) { uses[x.a]++; uses[x.b]++; }
            if (x.op == V_TAP || x.op == V_GATE) uses[x.a]++;
        }
        for (uint32_t o : sched_.outputs) uses[o]++;
        for (const auto& inst : sched_.ext) for (uint32_t in : inst.inputs) uses[in]++;
        for (size_t j = 0; j < sched_.ext.size(); j++) {
            const ExtInstance& fb = sched_.ext[j];
            if (fb.kind != EXT_FBDELAY || fb.n_lanes == 0) continue;
            const Value& v0 = V[fb.inputs[0]];
            if (v0.op != V_EXT) continue;
            const uint32_t i = v0.a;
            const ExtInstance& df = sched_.ext[i];
            if (df.kind != EXT_DIRECTFORM || df.n_lanes != fb.n_lanes || df.stage != fb.stage || chained_[i]) continue;
            bool ok = true;
            for (uint32_t l = 0; l < fb.n_lanes && ok; l++) {
                const Value& v = V[fb.inputs[l]];
                ok = v.op == V_EXT && v.a == i && v.imm == l && uses[fb.inputs[l]] == 1;
            }
            if (!ok || !chain_fusable(*df_defs_.at(df.key), *fb_defs_.at(fb.key))) continue;
            chain_state_[j] = chain_state_create(fb.n_lanes);
            if (!chain_state_[j]) throw Error{FRB_E_CUDA, "out of device memory (chain state)"};
            chain_of_[j] = (int32_t)i;
            chained_[i] = 1;
        }
    }
    exc_of_.assign(sched_.ext.size(), -1);
    exc_fused_.assign(sched_.ext.size(), 0);
    d_exc_voice_.assign(sched_.ext.size(), nullptr);
    if (!(cfg_.flags & (FRB_FLAG_NO_CHAIN_FUSION | FRB_FLAG_NO_EXCITER_FUSION))) {
        const auto& V = sched_.values;
        std::vector<uint32_t> uses(V.size(), 0);
        for (const Value& x : V) {
            if (x.op == V_DELAY || (x.op >= V_SUM2 && x.op <= V_MIN)) { uses[x.a]++; uses[x.b]++; }
            if (x.op == V_TAP || x.op == V_GATE) uses[x.a]++;
        }
        for (uint32_t o : sched_.outputs) uses[o]++;
        for (const auto& inst : sched_.ext) for (uint32_t in : inst.inputs) uses[in]++;
        std::vector<uint32_t> used_voices(sched_.ext.size(), 0);   // per instance: lanes somebody reads
        for (size_t k = 0; k < V.size(); k++) if (V[k].op == V_EXT && uses[k]) used_voices[V[k].a]++;
        for (size_t j = 0; j < sched_.ext.size(); j++) {
            if (chain_of_[j] < 0) continue;
            const ExtInstance& df = sched_.ext[chain_of_[j]];
            const Value& v0 = V[df.inputs[0]];
            if (v0.op != V_EXT) continue;
            const uint32_t o = v0.a;
            const ExtInstance& bank = sched_.ext[o];
            if (bank.kind != EXT_OSCBANK || exc_fused_[o] || used_voices[o] != df.n_lanes) continue;
            const auto bank_def = osc_defs_.find(bank.key);
            if (bank_def == osc_defs_.end() || !osc_one_source(*bank_def->second, nullptr)) continue;
            std::vector<uint32_t> voice(df.n_lanes);
            bool ok = true;
            for (uint32_t l = 0; l < df.n_lanes && ok; l++) {
                const Value& v = V[df.inputs[l]];
                ok = v.op == V_EXT && v.a == o && uses[df.inputs[l]] == 1;
                voice[l] = v.imm;
            }
            if (!ok) continue;
            CU(cudaMalloc(&d_exc_voice_[j], voice.size() * sizeof(uint32_t)));
            CU(cudaMemcpy(d_exc_voice_[j], voice.data(), voice.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            exc_of_[j] = (int32_t)o;
            exc_fused_[o] = 1;
        }
    }
    // Block length: as long as the rings allow (more samples per launch = more parallelism for the stage kernels — the
    // ordered fold over an extension node's lanes is parallel across the samples of a block only: cfg3's mix takes 1.32 ms
    // with 262,144-sample blocks, 1.50 ms with 65,536, 11.5 ms with 4,096), with the ring memory (<= 2 x block x 4 B per
    // allocated ring) kept under ~16 GB of the 180.
    {
        uint64_t nb = 0;
        for (const BufferInfo& bi : sched_.buffers) nb += !(bi.ext != ~0u && (chained_[bi.ext] || exc_fused_[bi.ext]));
        nb = std::max<uint64_t>(nb, 1);
        uint64_t c = 1ull << 22;
        while (c > (1ull << 14) && 2 * c * nb > (1ull << 32)) c >>= 1;
        if (const char* e = getenv("FRB_BLOCK_SAMPLES")) {           // measurement knob (tools/k4_probe.py): a power of two >= 1024
            const uint64_t v = strtoull(e, nullptr, 10);
            if (v >= 1024 && (v & (v - 1)) == 0) c = v;
        }
        chunk_ = c;
    }
}

// Ring buffers: capacity = pow2 >= lookback + block (+ slack); LOOKBACK_FULL rings hold [0, t_end).
// The lanes of one extension instance share one allocation, lane l at base + l * capacity (they have the same
// lookback), so a fold over consecutive lanes needs no per-plane descriptor.
void Renderer::ensure_rings(uint64_t t_end) {
    if (ring_groups_.empty() && !sched_.buffers.empty()) {
        for (size_t i = 0; i < sched_.buffers.size();) {
            size_t n = 1;
            if (sched_.buffers[i].ext != ~0u)
                while (i + n < sched_.buffers.size() && sched_.buffers[i + n].ext == sched_.buffers[i].ext) n++;
            ring_groups_.push_back(RingGroup{(uint32_t)i, (uint32_t)n, nullptr, 0});
            i += n;
        }
    }
    for (auto& g : ring_groups_) {
        const BufferInfo& b = sched_.buffers[g.first];
        if (b.ext != ~0u && (chained_[b.ext] || exc_fused_[b.ext])) continue;   // biquad / exciter of a fused chain: its output stays in registers
        const bool full = b.lookback == LOOKBACK_FULL;
        uint64_t need = full ? pow2_ceil(t_end + 8) : pow2_ceil(b.lookback + chunk_ + 8);
        if (need <= g.cap) continue;
        // lane stride = capacity + 96 floats: consecutive lanes must not sit at the same power-of-two offset (the
        // fold reads the same time index of every lane back to back)
        const uint64_t lstride = need + (g.count > 1 ? 96 : 0);
        float* nd = nullptr;
        CU(cudaMalloc(&nd, (size_t)g.count * lstride * sizeof(float)));
        CU(cudaMemsetAsync(nd, 0, (size_t)g.count * lstride * sizeof(float), stream_));
        if (g.data) {
            if (full) {   // index t & mask == t for t < old capacity: a prefix copy per lane preserves the history
                for (uint32_t l = 0; l < g.count; l++)
                    CU(cudaMemcpyAsync(nd + (size_t)l * lstride, g.data + (size_t)l * (g.cap + 96), g.cap * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
            } else {
                cache_valid_ = false;
            }
            CU(cudaStreamSynchronize(stream_));
            CU(cudaFree(g.data));
        }
        g.data = nd;
        g.cap = need;
        for (uint32_t l = 0; l < g.count; l++) {
            h_bufdesc_[g.first + l].data = nd + (size_t)l * lstride;
            h_bufdesc_[g.first + l].mask = need - 1;
        }
        bufdesc_dirty_ = true;
    }
    if (bufdesc_dirty_ && !h_bufdesc_.empty()) {
        if (d_bufdesc_cap_ < h_bufdesc_.size()) {
            if (d_bufdesc_) { CU(cudaStreamSynchronize(stream_)); CU(cudaFree(d_bufdesc_)); }
            d_bufdesc_cap_ = h_bufdesc_.size() * 2;
            CU(cudaMalloc(&d_bufdesc_, d_bufdesc_cap_ * sizeof(BufferDesc)));
        }
        CU(cudaMemcpyAsync(d_bufdesc_, h_bufdesc_.data(), h_bufdesc_.size() * sizeof(BufferDesc), cudaMemcpyHostToDevice, stream_));
        CU(cudaStreamSynchronize(stream_));
    }
    bufdesc_dirty_ = false;
}

// ------------------------------------------------------------------------------------------------ inputs

void Renderer::materialise_slot(size_t r) {
    while (inputs_.size() <= r) {
        size_t s = inputs_.size();
        uint64_t base = 0;
        for (auto& ep : epochs_) if (s < ep.first) { base = ep.second; break; }
        InputSlot sl;
        sl.base = base & ~3ull;
        sl.end = base;               // `base` zeros so far; [sl.base, base) are explicit zeros once allocated
        inputs_.push_back(sl);
    }
}

void Renderer::grow_slot(InputSlot& s, uint64_t need_end) {
    uint64_t need = need_end - s.base;
    if (need <= s.cap) return;
    uint64_t ncap = std::max<uint64_t>(pow2_ceil(need + 8), 1024);
    float* nd = nullptr;
    CU(cudaMalloc(&nd, ncap * sizeof(float)));
    CU(cudaMemsetAsync(nd, 0, ncap * sizeof(float), stream_));
    if (s.d_data) {
        if (s.end > s.base) CU(cudaMemcpyAsync(nd, s.d_data, (s.end - s.base) * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
        CU(cudaStreamSynchronize(stream_));
        CU(cudaFree(s.d_data));
    }
    s.d_data = nd;
    s.cap = ncap;
}

void Renderer::validate_inputs(uint32_t n_slots, uint64_t n_times, uint64_t idx, const uint64_t* offs, uint32_t n_rows) const {
    const bool seek = idx != head_;
    const uint64_t buff_len = (uint64_t)n_slots * n_times;                // ndarray len(): element count (:60)
    const uint64_t n_vec_after = std::max(n_slot_vectors_, buff_len);
    const size_t n_fed = (size_t)std::min<uint64_t>(n_rows, n_vec_after); // zip(rows, slots) stops at the shorter (:68)
    // everything is checked before any state changes (the reference panics half-way; we refuse the call instead)
    for (size_t r = 0; r < n_fed; r++) {
        if (offs[r + 1] < offs[r]) throw Error{FRB_E_INVALID, "in_row_offsets must be non-decreasing"};
        uint64_t row_len = offs[r + 1] - offs[r];
        // length of the reference's slot vector r when the zip reaches it
        uint64_t len_now = idx;                                           // after a seek (:52-58) or created now (:60-65)
        if (!seek && r < n_slot_vectors_) {
            if (r < inputs_.size()) len_now = inputs_[r].end;
            else for (auto& ep : epochs_) if (r < ep.first) { len_now = ep.second; break; }
        }
        if (len_now != idx) throw Error{FRB_E_INPUT_GAP, "input slot " + std::to_string(r) + " has length " + std::to_string(len_now) + " != idx (reference.rs:69 assert_eq)"};
        if (row_len > n_times) throw Error{FRB_E_INPUT_TOO_LONG, "cannot send inputs ahead of outputs (reference.rs:71 assert)"};
    }
}

// reference.rs:49-75
void Renderer::ingest_inputs(uint32_t n_slots, uint64_t n_times, uint64_t idx, const float* in_data, bool in_on_device,
                             const uint64_t* offs, uint32_t n_rows) {
    const bool seek = idx != head_;
    const uint64_t buff_len = (uint64_t)n_slots * n_times;                // ndarray len(): element count (:60)
    const uint64_t n_vec_after = std::max(n_slot_vectors_, buff_len);
    const size_t n_fed = (size_t)std::min<uint64_t>(n_rows, n_vec_after); // zip(rows, slots) stops at the shorter (:68)

    if (seek) {                                                            // :52-58
        for (auto& s : inputs_) { s.base = idx & ~3ull; s.end = idx; if (s.d_data && s.cap) CU(cudaMemsetAsync(s.d_data, 0, std::min<uint64_t>(s.cap, 8) * sizeof(float), stream_)); }
        epochs_.clear();
        if (n_slot_vectors_ > inputs_.size()) epochs_.emplace_back(n_slot_vectors_, idx);
        cache_valid_ = false;
    }
    if (n_slot_vectors_ < buff_len) {                                      // :60-65
        epochs_.emplace_back(buff_len, idx);
        n_slot_vectors_ = buff_len;
    }
    if (n_fed > 0) {
        const float* d_rows = in_data;
        cudaMemcpyKind rows_kind = cudaMemcpyDeviceToDevice;
        uint64_t total = offs[n_fed] - offs[0];
        if (!in_on_device && total > 0) {
            if (total * sizeof(float) <= kPinBytes) {
                // short call: the rows go through pinned memory (no stream wait) straight into their history slots
                if (!h_pin_in_) CU(cudaMallocHost(&h_pin_in_, kPinBytes));
                CU(cudaStreamSynchronize(stream_));                        // the previous call's copies out of h_pin_in_ (idle stream: free)
                memcpy(h_pin_in_, in_data + offs[0], total * sizeof(float));
                rows_kind = cudaMemcpyHostToDevice;
                d_rows = h_pin_in_ - offs[0];
            } else {
                if (d_in_stage_cap_ < total) {
                    if (d_in_stage_) { CU(cudaStreamSynchronize(stream_)); CU(cudaFree(d_in_stage_)); }
                    d_in_stage_cap_ = pow2_ceil(total);
                    CU(cudaMalloc(&d_in_stage_, d_in_stage_cap_ * sizeof(float)));
                }
                CU(cudaMemcpyAsync(d_in_stage_, in_data + offs[0], total * sizeof(float), cudaMemcpyHostToDevice, stream_));
                d_rows = d_in_stage_ - offs[0];
            }
            stats.h2d_bytes += total * sizeof(float);
        }
        // many long rows: one kernel for all of them instead of a device-to-device copy each
        const bool batched = n_fed >= 4 && n_fed <= 65535 && total >= (1ull << 20);   // rows are grid.y
        std::vector<IngestRow> batch;
        for (size_t r = 0; r < n_fed; r++) {
            materialise_slot(r);
            InputSlot& s = inputs_[r];
            uint64_t row_len = offs[r + 1] - offs[r];
            grow_slot(s, idx + n_times);
            if (row_len && batched) batch.push_back(IngestRow{s.d_data + (idx - s.base), d_rows + offs[r], row_len});
            else if (row_len) CU(cudaMemcpyAsync(s.d_data + (idx - s.base), d_rows + offs[r], row_len * sizeof(float), rows_kind, stream_));
        }
        if (!batch.empty()) {
            const size_t bytes = batch.size() * sizeof(IngestRow);
            if (d_ingest_cap_ < batch.size()) {
                if (d_ingest_) { CU(cudaStreamSynchronize(stream_)); CU(cudaFree(d_ingest_)); }
                d_ingest_cap_ = pow2_ceil(batch.size());
                CU(cudaMalloc(&d_ingest_, d_ingest_cap_ * sizeof(IngestRow)));
            }
            CU(cudaMemcpyAsync(d_ingest_, batch.data(), bytes, cudaMemcpyHostToDevice, stream_));   // pageable: staged before the call returns
            uint64_t longest = 0;
            for (const IngestRow& b : batch) longest = std::max<uint64_t>(longest, b.n);
            const unsigned per_row = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((longest / 4 + 1023) / 1024, (uint64_t)sm_count_ * 8 / batch.size()));   // rounded down: the grid fits in one wave (8 CTAs of 256 threads per SM)
            ingest_rows_kernel<<<dim3(per_row, (unsigned)batch.size()), 256, 0, stream_>>>(static_cast<const IngestRow*>(d_ingest_));
            CU(cudaGetLastError());
            stats.kernel_launches++;
        }
        for (size_t r = 0; r < n_fed; r++) {                               // after the rows are in: pad with the last value (:72-73)
            InputSlot& s = inputs_[r];
            uint64_t row_len = offs[r + 1] - offs[r];
            uint64_t filled_end = idx + row_len;
            if (filled_end < idx + n_times) {
                // Vec::last of the slot after the extend: the row's last value, else the previous content's last
                // value (explicit zeros count), else 0 for an empty vector.
                // Times below s.base are implicit zeros, so a "last" there (or an empty vector) pads with 0.
                int has_last = filled_end > s.base;
                uint64_t last_pos = has_last ? filled_end - 1 - s.base : 0;
                uint64_t count = idx + n_times - filled_end;
                unsigned blocks = (unsigned)std::min<uint64_t>((count + 255) / 256, 1024);
                pad_kernel<<<blocks, 256, 0, stream_>>>(s.d_data, filled_end - s.base, count, last_pos, has_last);
                CU(cudaGetLastError());
                stats.kernel_launches++;
            }
            s.end = idx + n_times;
        }
    }
    // device table of input descriptors for the slots the schedule reads
    // (built in fill() after the schedule is known)
}

// ------------------------------------------------------------------------------------------------ rendering

// Decides whether stage `sg` runs compiled: starts / collects the NVRTC build according to the policy flags.
void Renderer::poll_stage_jit(size_t sg, uint64_t n_groups) {
    const Stage& st = sched_.stages[sg];
    StageJit& sj = stage_jit_[sg];
    sj.uses++;
    const bool asked = (cfg_.flags & FRB_FLAG_JIT_EAGER) != 0;
    const bool eager = asked || n_groups >= (1ull << 15);
    if (sj.state == 0 && !(cfg_.flags & FRB_FLAG_NO_JIT) && (eager || sj.uses >= 4)) {
        // What NVRTC is handed is bounded (jit.cc: runs of like instruction groups are loops, so this is the program's
        // STRUCTURE, not its length; straight-line code costs 2 s at 200 statements, 16 s at 500): a stage above
        // JIT_MAX_CODE stays on the interpreter for good, and one above JIT_MAX_SYNC_CODE is never compiled on the
        // render thread unless the caller asked for that with FRB_FLAG_JIT_EAGER.  A cubin the process has compiled
        // before (same structure: other constants, other repeat counts, another renderer) is loaded right away.
        JitProgram prog = jit_generate(st);
        sj.code_instrs = prog.code_instrs;
        sj.table = std::move(prog.table);
        sj.groups_per_thread = prog.groups_per_thread;
        std::string jerr;
        if (auto hit = jit_cache_lookup(prog.source)) {
            sj.k = jit_load(*hit, sj.table, sj.groups_per_thread, &jerr);
            sj.state = sj.k ? 1 : 2;
            if (!sj.k) last_jit_error = jerr;
        } else if (sj.code_instrs > JIT_MAX_CODE) {
            sj.state = 2;
            last_jit_error = "stage program too long for the JIT; interpreted";
        } else if (eager && (asked || sj.code_instrs <= JIT_MAX_SYNC_CODE)) {
            // a long block (or an explicit request) pays for the NVRTC call right away
            std::string cubin, log;
            if (jit_compile_to_cubin(prog.source, &cubin, &log)) sj.k = jit_load(cubin, sj.table, sj.groups_per_thread, &jerr);
            else jerr = "NVRTC: " + log;
            sj.state = sj.k ? 1 : 2;
            if (!sj.k) last_jit_error = jerr;
        } else {
            // streaming in short blocks (or a long body): compile beside the render loop, never stall a block for it
            sj.job = jit_compile_async(std::move(prog.source));
            sj.state = 3;
        }
        if (sj.state != 3) std::vector<uint32_t>().swap(sj.table);
    }
    if (sj.state == 3 && sj.job->done.load(std::memory_order_acquire) != 0) {
        std::string jerr;
        const bool ok = sj.job->done.load(std::memory_order_acquire) > 0;
        sj.k = ok ? jit_load(sj.job->cubin, sj.table, sj.groups_per_thread, &jerr) : nullptr;
        sj.state = sj.k ? 1 : 2;
        if (!sj.k) last_jit_error = ok ? jerr : "NVRTC: " + sj.job->log;
        sj.job.reset();
        std::vector<uint32_t>().swap(sj.table);
    }
}

void Renderer::run_range(uint64_t lo, uint64_t hi, float* d_out, uint64_t t0, uint64_t t1, uint64_t out_stride) {
    if (hi <= lo) return;
    // with profiling on, every kernel family is bracketed by CUDA events on the renderer's stream
    auto timed = [&](float frb_timing::*field, auto&& launch) {
        if (profiling) CU(cudaEventRecord(ev_[2], stream_));
        launch();
        if (profiling) {
            CU(cudaEventRecord(ev_[3], stream_));
            CU(cudaEventSynchronize(ev_[3]));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
            timing.*field += ms;
        }
    };
    const int out_vec_ok = d_out && (t0 % 4 == 0) && (out_stride % 4 == 0) && ((uintptr_t)d_out % 16 == 0);
    // The recurrence kernels move 128-bit quads only from a block start that is a multiple of 4 (8 when the chain
    // evaluates its exciters); an unaligned start would cost them the fast tiles of the whole chunk (2.2x at
    // 4,097-sample blocks).  The at most 7 samples up to the next multiple of 8 are therefore rendered as a range of
    // their own — the same thing as two consecutive calls — when the rest is long enough to pay for the launches.
    bool has_recurrence = false;
    for (const ExtInstance& x : sched_.ext) has_recurrence |= (x.kind == EXT_DIRECTFORM || x.kind == EXT_FBDELAY);
    constexpr uint64_t kAlignSplitMin = 2048;
    uint64_t c0 = lo;
    while (c0 < hi) {
        uint64_t c1 = std::min(hi, (c0 / chunk_ + 1) * chunk_);
        if (has_recurrence && align_split_ && (c0 & 7) && c1 - c0 >= kAlignSplitMin) c1 = (c0 | 7) + 1;
        for (size_t sg = 0; sg < sched_.stages.size(); sg++) {
            const Stage& st = sched_.stages[sg];
            for (uint32_t xi : st.ext) {
                const ExtInstance& x = sched_.ext[xi];
                uint64_t nl = 0, nt = 0;
                if (x.kind == EXT_OSCBANK && exc_fused_[xi]) {
                    continue;                                   // evaluated inside the chain kernel that reads it
                } else if (x.kind == EXT_OSCBANK) {
                    timed(&frb_timing::osc_ms, [&] { CU(launch_osc(*osc_defs_.at(x.key), d_bufdesc_, x.first_out_buf, c0, c1, cfg_.osc_anchor, sm_count_, stream_, &nl, &nt)); });
                    stats.osc_launches += nl;
                    stats.osc_tensor_launches += nt;
                } else if (x.kind == EXT_DIRECTFORM && chained_[xi]) {
                    continue;                                   // runs inside its comb's launch
                } else if (x.kind == EXT_FBDELAY && chain_of_[xi] >= 0) {
                    const uint32_t di = (uint32_t)chain_of_[xi];
                    const ExtInstance& df = sched_.ext[di];
                    OscOneSrc exc;
                    const bool has_exc = exc_of_[xi] >= 0 && osc_one_source(*osc_defs_.at(sched_.ext[exc_of_[xi]].key), &exc);
                    if (exc_of_[xi] >= 0 && !has_exc) throw Error{FRB_E_INVALID, "oscillator bank changed shape under a fused chain"};
                    timed(&frb_timing::scan_ms, [&] { CU(launch_dfcomb(*df_defs_.at(df.key), *fb_defs_.at(x.key), *chain_state_[xi], d_bufdesc_, d_ext_in_bufs_[di], x.first_out_buf, c0, c1, stream_, &nl, has_exc ? &exc : nullptr, d_exc_voice_[xi])); });
                    stats.scan_launches += nl;
                    stats.chain_launches += nl;
                } else if (x.kind == EXT_DIRECTFORM) {
                    timed(&frb_timing::scan_ms, [&] { CU(launch_directform(*df_defs_.at(x.key), d_bufdesc_, d_ext_in_bufs_[xi], x.first_out_buf, c0, c1, sm_count_, stream_, &nl)); });
                    stats.scan_launches += nl;
                } else {
                    timed(&frb_timing::scan_ms, [&] { CU(launch_fbdelay(*fb_defs_.at(x.key), d_bufdesc_, d_ext_in_bufs_[xi], x.first_out_buf, c0, c1, sm_count_, stream_, &nl)); });
                    stats.scan_launches += nl;
                }
                stats.kernel_launches += nl;
            }
            for (const FoldJob& fj : st.folds) {
                timed(&frb_timing::interp_ms, [&] { CU(launch_fold(d_bufdesc_, fj.first_buf, fj.count, fj.out_buf, c0, c1, sm_count_, stream_)); });
                stats.kernel_launches++;
                stats.interp_launches++;
            }
            if (st.program.size() <= 2) continue;   // only I_END (+ pad)
            InterpParams p;
            p.program = d_programs_[sg];
            p.n_strands = (unsigned)st.strand_offsets.size() - 1;
            p.prog_base = (unsigned)((st.strand_offsets.size() + 3) / 4 * 4);
            p.n_instr = 0;
            for (size_t k = 0; k + 1 < st.strand_offsets.size(); k++)
                p.n_instr = std::max(p.n_instr, st.strand_offsets[k + 1] - st.strand_offsets[k]);   // longest strand
            p.inputs = d_indesc_;
            p.n_inputs = n_indesc_;
            p.buffers = d_bufdesc_;
            p.out = d_out;
            p.out_stride = out_stride;
            p.t_begin = c0 & ~7ull;
            p.n_groups = ((c1 + 7) / 8) - (c0 / 8);
            p.t0 = t0;
            p.t1 = d_out ? t1 : t0;                  // warm-up ranges write no output
            p.out_vec_ok = out_vec_ok;
            p.sparkle_delay = (cfg_.flags & FRB_FLAG_SPARKLE_DELAY) ? 1 : 0;
            p.sparkle_min = (cfg_.flags & FRB_FLAG_SPARKLE_MIN) ? 1 : 0;
            // Tiered like the reference's JIT renderer (sparkle.rs:271-288 compiles lazily at the next render): a stage
            // program is interpreted until it is hot (4th launch, or a block of >= 256 Ki samples), then compiled
            // once by NVRTC into a fused kernel.  A stage that fails to compile stays on the interpreter.
            poll_stage_jit(sg, p.n_groups);
            timed(&frb_timing::interp_ms, [&] {
                StageJit& sj = stage_jit_[sg];
                if (sj.state == 1 && jit_launch(sj.k, p, sm_count_, stream_)) stats.jit_launches++;
                else CU(launch_interp(p, st.n_regs, sm_count_, stream_));
            });
            stats.kernel_launches++;
            stats.interp_launches++;
        }
        c0 = c1;
    }
}

// FRB_TRACE_FILL=1: host-clock time per phase of fill(), summed over calls and printed when the renderer goes away (tuning aid)
namespace {
struct FillTrace {
    bool on = getenv("FRB_TRACE_FILL") != nullptr;
    double ns[8] = {0}; uint64_t calls = 0;
    std::chrono::steady_clock::time_point t;
    void start() { if (on) t = std::chrono::steady_clock::now(); }
    void lap(int i) { if (!on) return; auto n = std::chrono::steady_clock::now(); if (calls > 2000) ns[i] += std::chrono::duration<double, std::nano>(n - t).count(); t = n; }
    ~FillTrace() {
        if (calls > 2000) calls -= 2000;
        if (on && calls) fprintf(stderr, "[frb] fill x %llu: schedule %.2f ingest %.2f table %.2f rings %.2f run %.2f d2h %.2f wait %.2f us per call\n",
                                 (unsigned long long)calls, ns[0] / calls / 1e3, ns[1] / calls / 1e3, ns[2] / calls / 1e3, ns[3] / calls / 1e3,
                                 ns[4] / calls / 1e3, ns[5] / calls / 1e3, ns[6] / calls / 1e3);
    }
} g_fill_trace;

}  // namespace

void Renderer::fill(float* out, bool out_on_device, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                    const float* in_data, bool in_on_device, const uint64_t* offs, uint32_t n_rows) {
    g_fill_trace.start();
    g_fill_trace.calls++;
    require_device();
    CU(cudaSetDevice(device_));
    if (n_rows && !offs) throw Error{FRB_E_INVALID, "in_row_offsets is NULL"};
    if (idx + n_times < idx) throw Error{FRB_E_INVALID, "idx + n_times overflows"};
    // the schedule maps input slots beyond its cap to zero: a call that feeds a slot up there (absurd, but legal) re-flattens
    if (std::max<uint64_t>(inputs_.size(), n_rows) > sched_input_cap_) dirty_ = true;
    ensure_schedule(n_slots);            // may throw on a malformed graph, before any state changes
    g_fill_trace.lap(0);
    if (profiling) { timing = frb_timing{}; CU(cudaEventRecord(ev_[0], stream_)); }
    // From here on state changes (input history, rings, recurrence carries).  A failure half-way (out of device memory,
    // a launch error) leaves the history ahead of the playhead, so the next call is made a seek whatever its idx: the
    // reference's seek rule (renderer.rs:12-15) then resets every slot and the rings are rebuilt — consistent again.
    validate_inputs(n_slots, n_times, idx, offs, n_rows);      // a refused call (reference: assert) changes nothing
    struct FailGuard {
        Renderer* r; bool armed = true;
        ~FailGuard() { if (armed) { r->cache_valid_ = false; r->head_ = ~0ull; } }
    } guard{this};
    ingest_inputs(n_slots, n_times, idx, in_data, in_on_device, offs, n_rows);
    g_fill_trace.lap(1);

    const uint64_t t1 = idx + n_times;
    bool pinned_out = false;
    uint64_t n_out_host = 0;
    if (n_times > 0 && n_slots > 0) {
        // input descriptor table for the slots the schedule reads
        // one entry per slot number the schedule names (all below its input-slot cap); slots no call ever fed are null
        const uint32_t nin = (uint32_t)sched_.n_input_slots;
        n_indesc_ = nin;
        if (nin) {
            std::vector<InputDesc> h(nin);
            for (uint32_t s = 0; s < nin; s++) {
                if (s < inputs_.size() && inputs_[s].d_data) h[s] = InputDesc{inputs_[s].d_data, inputs_[s].base, inputs_[s].end};
                else h[s] = InputDesc{nullptr, 0, 0};
            }
            if (d_indesc_cap_ < (size_t)nin) {
                if (d_indesc_) { CU(cudaStreamSynchronize(stream_)); CU(cudaFree(d_indesc_)); }
                d_indesc_cap_ = std::max<size_t>((size_t)nin * 2, 16);
                CU(cudaMalloc(&d_indesc_, d_indesc_cap_ * sizeof(InputDesc)));
                h_indesc_.clear();
            }
            // the table is re-uploaded only when it changed — when streaming, every call (the history's end moves) —
            // from one of two pinned buffers: asynchronous, no wait on the stream
            if (h_indesc_.size() != h.size() || memcmp(h_indesc_.data(), h.data(), h.size() * sizeof(InputDesc)) != 0) {
                h_indesc_ = h;
                if (h_pin_indesc_cap_ < h.size()) {
                    CU(cudaStreamSynchronize(stream_));
                    for (int i = 0; i < 2; i++) {
                        if (h_pin_indesc_[i]) CU(cudaFreeHost(h_pin_indesc_[i]));
                        h_pin_indesc_[i] = nullptr;
                        CU(cudaMallocHost(&h_pin_indesc_[i], d_indesc_cap_ * sizeof(InputDesc)));
                        if (!ev_indesc_[i]) CU(cudaEventCreateWithFlags(&ev_indesc_[i], cudaEventDisableTiming));
                    }
                    h_pin_indesc_cap_ = d_indesc_cap_;
                }
                const unsigned b = pin_indesc_next_++ & 1u;
                CU(cudaEventSynchronize(ev_indesc_[b]));                   // the copy that last read this buffer (two uploads ago)
                memcpy(h_pin_indesc_[b], h.data(), h.size() * sizeof(InputDesc));
                CU(cudaMemcpyAsync(d_indesc_, h_pin_indesc_[b], h.size() * sizeof(InputDesc), cudaMemcpyHostToDevice, stream_));
                CU(cudaEventRecord(ev_indesc_[b], stream_));
            }
        }
        g_fill_trace.lap(2);
        ensure_rings(t1);
        g_fill_trace.lap(3);

        float* d_out = out;
        const uint64_t n_out = (uint64_t)n_slots * n_times;
        if (!out_on_device) {
            if (d_out_cap_ < n_out) {
                if (d_out_) { CU(cudaStreamSynchronize(stream_)); CU(cudaFree(d_out_)); }
                d_out_cap_ = n_out;
                CU(cudaMalloc(&d_out_, d_out_cap_ * sizeof(float)));
            }
            d_out = d_out_;
        }
        if (!(cache_valid_ && cache_head_ == idx)) {
            // seek, graph edit or first call: rebuild the rings the block depends on (SURVEY.md A.3 trap T1:
            // constants and input history before idx are live, only the external inputs were zeroed by a seek)
            uint64_t start = sched_.from_zero ? 0 : (idx > sched_.max_lookback ? idx - sched_.max_lookback : 0);
            run_range(start & ~7ull, idx, nullptr, idx, idx, 0);
        }
        run_range(idx, t1, d_out, idx, t1, n_times);
        g_fill_trace.lap(4);
        cache_valid_ = true;
        cache_head_ = t1;
        if (!out_on_device) {
            pinned_out = n_out * sizeof(float) <= kPinBytes;
            if (pinned_out && !h_pin_out_) CU(cudaMallocHost(&h_pin_out_, kPinBytes));
            CU(cudaMemcpyAsync(pinned_out ? h_pin_out_ : out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, stream_));
            stats.d2h_bytes += n_out * sizeof(float);
            n_out_host = n_out;
        }
    }
    head_ = t1;                           // reference.rs:84
    if (profiling) {
        CU(cudaEventRecord(ev_[1], stream_));
        CU(cudaEventSynchronize(ev_[1]));
        CU(cudaEventElapsedTime(&timing.total_ms, ev_[0], ev_[1]));
    }
    g_fill_trace.lap(5);
    if (!out_on_device) {
        CU(cudaStreamSynchronize(stream_));
        if (pinned_out) memcpy(out, h_pin_out_, n_out_host * sizeof(float));
    }
    g_fill_trace.lap(6);
    guard.armed = false;
}

// N4 (include/friendship_b200.h): consecutive fill_buffer calls of `block` samples, two blocks in flight.
void Renderer::render_stream(uint32_t n_slots, uint64_t idx, uint64_t n_total, uint64_t block, uint32_t n_in_rows,
                             frb_source_fn source, frb_sink_fn sink, void* user) {
    require_device();
    CU(cudaSetDevice(device_));
    if (!sink) throw Error{FRB_E_INVALID, "render_stream: sink is NULL"};
    if (n_in_rows && !source) throw Error{FRB_E_INVALID, "render_stream: inputs without a source"};
    if (block == 0) throw Error{FRB_E_INVALID, "render_stream: block must be positive"};
    if (idx + n_total < idx) throw Error{FRB_E_INVALID, "idx + n_total overflows"};
    block = std::min(block, std::max<uint64_t>(n_total, 1));
    if (!copy_stream_) CU(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    const size_t out_n = (size_t)n_slots * block, in_n = (size_t)n_in_rows * block;
    for (auto& st : sstage_) {
        if (!st.rendered) { CU(cudaEventCreateWithFlags(&st.rendered, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming)); }
        if (st.out_cap < out_n) {
            CU(cudaStreamSynchronize(stream_)); CU(cudaStreamSynchronize(copy_stream_));
            if (st.d_out) CU(cudaFree(st.d_out));
            if (st.h_out) CU(cudaFreeHost(st.h_out));
            st.d_out = st.h_out = nullptr; st.out_cap = 0;
            CU(cudaMalloc(&st.d_out, std::max<size_t>(out_n, 1) * sizeof(float)));
            CU(cudaMallocHost(&st.h_out, std::max<size_t>(out_n, 1) * sizeof(float)));
            st.out_cap = out_n;
        }
        if (st.in_cap < in_n) {
            CU(cudaStreamSynchronize(stream_));
            if (st.h_in) CU(cudaFreeHost(st.h_in));
            st.h_in = nullptr; st.in_cap = 0;
            CU(cudaMallocHost(&st.h_in, in_n * sizeof(float)));
            st.in_cap = in_n;
        }
    }
    std::vector<uint64_t> offs(n_in_rows + 1);
    struct Pending { bool live = false; uint64_t t = 0, n = 0; } pend[2];
    auto deliver = [&](int b) {                               // wait for block b's copy, hand it to the sink
        if (!pend[b].live) return;
        pend[b].live = false;
        CU(cudaEventSynchronize(sstage_[b].copied));
        if (sink(user, sstage_[b].h_out, n_slots, pend[b].n, pend[b].t) != 0) throw Error{FRB_E_INVALID, "render_stream: sink failed"};
    };
    try {
        // a zero-length render is one zero-length fill_buffer (bookkeeping only) and one empty buffer for the sink
        const uint64_t n_blocks = n_total ? (n_total + block - 1) / block : 1;
        uint64_t t = idx;
        for (uint64_t k = 0; k < n_blocks; k++) {
            const int b = (int)(k & 1);
            StreamStage& st = sstage_[b];
            const uint64_t n = std::min(block, idx + n_total - t);
            if (n_in_rows) {
                // st.h_in was last read by the host->device copy of block k-2, which finished before that block was delivered
                if (source(user, st.h_in, n_in_rows, n, t) != 0) throw Error{FRB_E_INVALID, "render_stream: source failed"};
                for (uint32_t r = 0; r <= n_in_rows; r++) offs[r] = (uint64_t)r * n;
            }
            // st.d_out was last read by the device->host copy of block k-2 (delivered already: the event has fired)
            fill(st.d_out, true, n_slots, n, t, n_in_rows ? st.h_in : nullptr, false, offs.data(), n_in_rows);
            CU(cudaEventRecord(st.rendered, stream_));
            CU(cudaStreamWaitEvent(copy_stream_, st.rendered, 0));
            if ((size_t)n_slots * n)
                CU(cudaMemcpyAsync(st.h_out, st.d_out, (size_t)n_slots * n * sizeof(float), cudaMemcpyDeviceToHost, copy_stream_));
            stats.d2h_bytes += (size_t)n_slots * n * sizeof(float);
            CU(cudaEventRecord(st.copied, copy_stream_));
            pend[b] = Pending{true, t, n};
            deliver(b ^ 1);                                   // block k-1: the GPU is busy with block k meanwhile
            t += n;
        }
        deliver((int)((n_blocks - 1) & 1));                   // only the last block is still pending
    } catch (...) {
        cudaStreamSynchronize(stream_);                       // nothing of an aborted render stays in flight
        cudaStreamSynchronize(copy_stream_);
        throw;
    }
}

void Renderer::sync() {
    if (host_only_) return;
    CU(cudaSetDevice(device_));
    CU(cudaStreamSynchronize(stream_));
}

}  // namespace frb
