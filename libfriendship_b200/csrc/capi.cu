// capi.cu — extern "C" surface of libfriendship_b200.so (include/friendship_b200.h).
// Each entry point cites the reference interface it replaces in the header.  Nothing throws across the ABI:
// frb::Error becomes a status code + frb_last_error().
#include <new>
#include <string>

#include <algorithm>
#include <cstring>

#include "jit.hpp"
#include "multi.hpp"
#include "renderer.hpp"

using frb::Error;
using frb::MultiRenderer;
using frb::Renderer;

// one B200 (frb_config::n_devices <= 1) or several behind the same handle (multi.cu)
struct frb_renderer {
    std::unique_ptr<Renderer> one;
    std::unique_ptr<MultiRenderer> many;
    explicit frb_renderer(const frb_config& c) {
        if (c.n_devices > 1 && c.device >= 0) many = std::make_unique<MultiRenderer>(c);
        else one = std::make_unique<Renderer>(c);
    }
    std::string& last_error() { return one ? one->last_error : many->last_error; }
    template <typename F>
    auto with(F f) { return one ? f(*one) : f(*many); }
};

static thread_local std::string g_create_error;

template <typename F>
static int guarded(frb_renderer* r, F f) {
    if (!r) return FRB_E_INVALID;
    try {
        f();
        return FRB_OK;
    } catch (const Error& e) {
        r->last_error() = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {
        r->last_error() = "out of host memory";
        return FRB_E_INVALID;
    } catch (const std::exception& e) {
        r->last_error() = e.what();
        return FRB_E_INVALID;
    }
}

extern "C" {

frb_renderer* frb_create(const frb_config* cfg) {
    frb_config c{};
    if (cfg) c = *cfg;
    try {
        return new frb_renderer(c);
    } catch (const Error& e) {
        g_create_error = e.msg;
    } catch (const std::exception& e) {
        g_create_error = e.what();
    }
    return nullptr;
}

void frb_destroy(frb_renderer* r) { delete r; }

const char* frb_last_error(const frb_renderer* r) { return r ? const_cast<frb_renderer*>(r)->last_error().c_str() : g_create_error.c_str(); }

int frb_define_effect(frb_renderer* r, uint64_t key, const frb_node* nodes, uint32_t n_nodes, const frb_edge* edges, uint32_t n_edges) {
    return guarded(r, [&] {
        if ((n_nodes && !nodes) || (n_edges && !edges)) throw Error{FRB_E_INVALID, "null array"};
        r->with([&](auto& impl) { impl.define_effect(key, nodes, n_nodes, edges, n_edges); });
    });
}
int frb_define_oscbank(frb_renderer* r, uint64_t key, const frb_oscbank_desc* d) {
    return guarded(r, [&] { if (!d) throw Error{FRB_E_INVALID, "null desc"}; r->with([&](auto& impl) { impl.define_oscbank(key, d); }); });
}
int frb_define_directform(frb_renderer* r, uint64_t key, const frb_directform_desc* d) {
    return guarded(r, [&] { if (!d) throw Error{FRB_E_INVALID, "null desc"}; r->with([&](auto& impl) { impl.define_directform(key, d); }); });
}
int frb_define_fbdelay(frb_renderer* r, uint64_t key, const frb_fbdelay_desc* d) {
    return guarded(r, [&] { if (!d) throw Error{FRB_E_INVALID, "null desc"}; r->with([&](auto& impl) { impl.define_fbdelay(key, d); }); });
}

int frb_add_node(frb_renderer* r, uint32_t handle, uint32_t kind, uint64_t key) {
    return guarded(r, [&] { r->with([&](auto& impl) { impl.add_node(handle, kind, key); }); });
}
int frb_del_node(frb_renderer* r, uint32_t handle) { return guarded(r, [&] { r->with([&](auto& impl) { impl.del_node(handle); }); }); }
int frb_add_edge(frb_renderer* r, frb_edge e) { return guarded(r, [&] { r->with([&](auto& impl) { impl.add_edge(e); }); }); }
int frb_del_edge(frb_renderer* r, frb_edge e) { return guarded(r, [&] { r->with([&](auto& impl) { impl.del_edge(e); }); }); }

int frb_fill_buffer(frb_renderer* r, float* out, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                    const float* in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows) {
    return guarded(r, [&] {
        if (!out && (uint64_t)n_slots * n_times) throw Error{FRB_E_INVALID, "out is NULL"};
        r->with([&](auto& impl) { impl.fill(out, false, n_slots, n_times, idx, in_data, false, in_row_offsets, n_in_rows); });
    });
}
int frb_fill_buffer_device(frb_renderer* r, float* d_out, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                           const float* d_in_data, const uint64_t* in_row_offsets, uint32_t n_in_rows) {
    return guarded(r, [&] {
        if (!d_out && (uint64_t)n_slots * n_times) throw Error{FRB_E_INVALID, "d_out is NULL"};
        r->with([&](auto& impl) { impl.fill(d_out, true, n_slots, n_times, idx, d_in_data, true, in_row_offsets, n_in_rows); });
    });
}
int frb_render_stream(frb_renderer* r, uint32_t n_slots, uint64_t idx, uint64_t n_total, uint64_t block,
                      uint32_t n_in_rows, frb_source_fn source, frb_sink_fn sink, void* user) {
    return guarded(r, [&] { r->with([&](auto& impl) { impl.render_stream(n_slots, idx, n_total, block, n_in_rows, source, sink, user); }); });
}
int frb_sync(frb_renderer* r) { return guarded(r, [&] { r->with([&](auto& impl) { impl.sync(); }); }); }
void* frb_stream(frb_renderer* r) { return r ? (void*)r->with([](auto& impl) { return impl.stream(); }) : nullptr; }

int frb_device_alloc(frb_renderer* r, uint64_t bytes, void** d_ptr_out) {
    return guarded(r, [&] {
        if (!d_ptr_out) throw Error{FRB_E_INVALID, "d_ptr_out is NULL"};
        r->with([&](auto& impl) { impl.use_device(); });
        cudaError_t e = cudaMalloc(d_ptr_out, bytes);
        if (e == cudaSuccess) e = cudaMemset(*d_ptr_out, 0, bytes);
        if (e != cudaSuccess) throw Error{FRB_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)};
    });
}
int frb_device_free(frb_renderer* r, void* d_ptr) {
    return guarded(r, [&] {
        r->with([&](auto& impl) { impl.use_device(); });
        cudaError_t e = cudaFree(d_ptr);
        if (e != cudaSuccess) throw Error{FRB_E_CUDA, std::string("cudaFree: ") + cudaGetErrorString(e)};
    });
}
int frb_ipc_export(frb_renderer* r, const void* d_ptr, unsigned char handle[64]) {
    return guarded(r, [&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        if (!d_ptr || !handle) throw Error{FRB_E_INVALID, "null argument"};
        r->with([&](auto& impl) { impl.use_device(); });
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr));
        if (e != cudaSuccess) throw Error{FRB_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)};
        memcpy(handle, &h, 64);
    });
}
int frb_ipc_open(frb_renderer* r, const unsigned char handle[64], void** d_ptr_out) {
    return guarded(r, [&] {
        if (!handle || !d_ptr_out) throw Error{FRB_E_INVALID, "null argument"};
        r->with([&](auto& impl) { impl.use_device(); });
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, 64);
        cudaError_t e = cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) throw Error{FRB_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)};
    });
}
int frb_ipc_close(frb_renderer* r, void* d_ptr) {
    return guarded(r, [&] {
        r->with([&](auto& impl) { impl.use_device(); });
        cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
        if (e != cudaSuccess) throw Error{FRB_E_CUDA, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e)};
    });
}
int frb_sum_rows(frb_renderer* r, float* d_out, const float* d_rows, uint32_t n_rows, uint64_t row_stride, uint64_t n) {
    return guarded(r, [&] { r->with([&](auto& impl) { impl.sum_rows(d_out, d_rows, n_rows, row_stride, n); }); });
}

int64_t frb_dump_schedule(frb_renderer* r, uint32_t n_slots, uint32_t* words, uint64_t cap) {
    int64_t n = 0;
    int rc = guarded(r, [&] {
        std::vector<uint32_t> w = r->with([&](auto& impl) -> const frb::Schedule& { return impl.schedule(n_slots); }).dump();
        n = (int64_t)w.size();
        if (words) for (uint64_t i = 0; i < cap && i < w.size(); i++) words[i] = w[i];
    });
    return rc == FRB_OK ? n : (int64_t)rc;
}

int64_t frb_dump_schedule_shard(frb_renderer* r, uint32_t n_slots, uint32_t rank, uint32_t world, uint32_t* words, uint64_t cap) {
    int64_t n = 0;
    int rc = guarded(r, [&] {
        std::vector<uint32_t> w = r->with([&](auto& impl) { return impl.schedule_for_shard(n_slots, rank, world); }).dump();
        n = (int64_t)w.size();
        if (words) for (uint64_t i = 0; i < cap && i < w.size(); i++) words[i] = w[i];
    });
    return rc == FRB_OK ? n : (int64_t)rc;
}
int frb_lane_use(frb_renderer* r, uint32_t n_slots) {
    int use = 0;
    int rc = guarded(r, [&] {
        use = (int)frb::lane_use_of_outputs(r->with([&](auto& impl) -> const frb::Schedule& { return impl.schedule(n_slots); }));
    });
    return rc == FRB_OK ? use : rc;
}

int64_t frb_jit_source(frb_renderer* r, uint32_t n_slots, uint32_t stage, char* out, uint64_t cap) {
    int64_t n = 0;
    int rc = guarded(r, [&] {
        const frb::Schedule& s = r->with([&](auto& impl) -> const frb::Schedule& { return impl.schedule(n_slots); });
        if (stage >= s.stages.size()) throw Error{FRB_E_INVALID, "no such stage"};
        std::string src = frb::jit_generate_source(s.stages[stage]);
        n = (int64_t)src.size();
        if (out && cap) {
            size_t m = std::min<size_t>(src.size(), cap - 1);
            memcpy(out, src.data(), m);
            out[m] = 0;
        }
    });
    return rc == FRB_OK ? n : (int64_t)rc;
}
int64_t frb_jit_cubin_size(frb_renderer* r, uint32_t n_slots, uint32_t stage) {
    int64_t n = 0;
    int rc = guarded(r, [&] {
        const frb::Schedule& s = r->with([&](auto& impl) -> const frb::Schedule& { return impl.schedule(n_slots); });
        if (stage >= s.stages.size()) throw Error{FRB_E_INVALID, "no such stage"};
        std::string cubin, log;
        if (!frb::jit_compile_to_cubin(frb::jit_generate(s.stages[stage]).source, &cubin, &log))
            throw Error{FRB_E_UNSUPPORTED, "NVRTC: " + log};
        n = (int64_t)cubin.size();
    });
    return rc == FRB_OK ? n : (int64_t)rc;
}

int64_t frb_jit_code_instructions(frb_renderer* r, uint32_t n_slots, uint32_t stage) {
    int64_t n = 0;
    int rc = guarded(r, [&] {
        const frb::Schedule& s = r->with([&](auto& impl) -> const frb::Schedule& { return impl.schedule(n_slots); });
        if (stage >= s.stages.size()) throw Error{FRB_E_INVALID, "no such stage"};
        n = (int64_t)frb::jit_code_instructions(s.stages[stage]);
    });
    return rc == FRB_OK ? n : (int64_t)rc;
}

int frb_get_stats(const frb_renderer* r, frb_stats* out) {
    if (!r || !out) return FRB_E_INVALID;
    *out = r->one ? r->one->stats : r->many->get_stats();
    return FRB_OK;
}
int frb_set_profiling(frb_renderer* r, int enabled) {
    if (!r) return FRB_E_INVALID;
    if (r->one) r->one->profiling = enabled != 0;
    else r->many->set_profiling(enabled != 0);
    return FRB_OK;
}
int frb_get_timing(const frb_renderer* r, frb_timing* out) {
    if (!r || !out) return FRB_E_INVALID;
    *out = r->one ? r->one->timing : r->many->get_timing();
    return FRB_OK;
}

const char* frb_version(void) { return "friendship_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
