"""ctypes declarations for the C ABI in include/friendship_b200.h.

The same struct layouts and call shapes are reused by the test infrastructure's CPU restatement of the reference
renderer (its `orc_*` mirror of this ABI loads this file by path); this module never loads or references it.
"""
import ctypes as C

import numpy as np

# status codes
FRB_OK = 0
FRB_E_BAD_HANDLE = -1
FRB_E_INPUT_TOO_LONG = -2
FRB_E_INPUT_GAP = -3
FRB_E_BAD_SLOT = -4
FRB_E_CUDA = -5
FRB_E_INVALID = -6
FRB_E_UNSUPPORTED = -7
FRB_E_EXISTS = -8
FRB_E_NO_DEVICE = -9

# node kinds (reference src/routing/effect.rs:86-112 + extensions)
KIND_DELAY = 0
KIND_F32CONSTANT = 1
KIND_SUM2 = 2
KIND_MULTIPLY = 3
KIND_DIVIDE = 4
KIND_MODULO = 5
KIND_MINIMUM = 6
KIND_EFFECT = 16
KIND_OSCBANK = 32
KIND_DIRECTFORM = 33
KIND_FBDELAY = 34

FLAG_SPARKLE_DELAY = 1
FLAG_NO_JIT = 2
FLAG_JIT_EAGER = 4
FLAG_NO_CHAIN_FUSION = 8
FLAG_NO_EXCITER_FUSION = 16
FLAG_SPARKLE_MIN = 32
FLAG_NO_TENSOR_OSC = 64


class frb_edge(C.Structure):
    _fields_ = [("from_", C.c_uint32), ("to", C.c_uint32), ("from_slot", C.c_uint32), ("to_slot", C.c_uint32)]


class frb_node(C.Structure):
    _fields_ = [("handle", C.c_uint32), ("kind", C.c_uint32), ("key", C.c_uint64)]


class frb_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("osc_anchor", C.c_uint32), ("n_devices", C.c_uint32)]


class frb_oscbank_desc(C.Structure):
    _fields_ = [("n_voices", C.c_uint32), ("reserved", C.c_uint32), ("n_partials", C.c_uint64),
                ("sample_rate", C.c_double),
                ("voice_offsets", C.POINTER(C.c_uint64)), ("freq_hz", C.POINTER(C.c_double)),
                ("amp", C.POINTER(C.c_float)), ("phase", C.POINTER(C.c_float)),
                ("attack", C.POINTER(C.c_float)), ("tau", C.POINTER(C.c_float))]


class frb_directform_desc(C.Structure):
    _fields_ = [("n_lanes", C.c_uint32), ("reserved", C.c_uint32),
                ("b0", C.POINTER(C.c_float)), ("b1", C.POINTER(C.c_float)), ("b2", C.POINTER(C.c_float)),
                ("a1", C.POINTER(C.c_float)), ("a2", C.POINTER(C.c_float))]


class frb_fbdelay_desc(C.Structure):
    _fields_ = [("n_lanes", C.c_uint32), ("reserved", C.c_uint32),
                ("delay", C.POINTER(C.c_uint32)), ("gain", C.POINTER(C.c_float))]


class frb_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kernel_launches", "h2d_bytes", "d2h_bytes", "schedule_builds",
                                           "osc_launches", "interp_launches", "scan_launches", "jit_launches",
                                           "chain_launches", "osc_tensor_launches")]


class frb_timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("osc_ms", "interp_ms", "scan_ms", "total_ms")]


# every symbol include/friendship_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = [
    "frb_create", "frb_destroy", "frb_last_error", "frb_define_effect", "frb_define_oscbank",
    "frb_define_directform", "frb_define_fbdelay", "frb_add_node", "frb_del_node", "frb_add_edge", "frb_del_edge",
    "frb_fill_buffer", "frb_fill_buffer_device", "frb_sync", "frb_stream", "frb_dump_schedule", "frb_get_stats",
    "frb_set_profiling", "frb_get_timing", "frb_version", "frb_jit_source", "frb_jit_cubin_size", "frb_jit_code_instructions", "frb_device_alloc", "frb_device_free", "frb_ipc_export", "frb_ipc_open", "frb_ipc_close", "frb_sum_rows",
    "frb_render_stream", "frb_lane_use", "frb_dump_schedule_shard",
]


def declare(lib, prefix):
    """Attach argtypes/restypes for the `<prefix>_*` renderer entry points present in `lib`."""
    vp = C.c_void_p
    sig = {
        "define_effect": ([vp, C.c_uint64, C.POINTER(frb_node), C.c_uint32, C.POINTER(frb_edge), C.c_uint32], C.c_int),
        "define_oscbank": ([vp, C.c_uint64, C.POINTER(frb_oscbank_desc)], C.c_int),
        "define_directform": ([vp, C.c_uint64, C.POINTER(frb_directform_desc)], C.c_int),
        "define_fbdelay": ([vp, C.c_uint64, C.POINTER(frb_fbdelay_desc)], C.c_int),
        "add_node": ([vp, C.c_uint32, C.c_uint32, C.c_uint64], C.c_int),
        "del_node": ([vp, C.c_uint32], C.c_int),
        "add_edge": ([vp, frb_edge], C.c_int),
        "del_edge": ([vp, frb_edge], C.c_int),
        "fill_buffer": ([vp, vp, C.c_uint32, C.c_uint64, C.c_uint64, vp, C.POINTER(C.c_uint64), C.c_uint32], C.c_int),
        "fill_buffer_device": ([vp, vp, C.c_uint32, C.c_uint64, C.c_uint64, vp, C.POINTER(C.c_uint64), C.c_uint32], C.c_int),
        "sync": ([vp], C.c_int),
        "stream": ([vp], vp),
        "destroy": ([vp], None),
        "last_error": ([vp], C.c_char_p),
        "dump_schedule": ([vp, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64], C.c_int64),
        "dump_schedule_shard": ([vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64], C.c_int64),
        "lane_use": ([vp, C.c_uint32], C.c_int),
        "get_stats": ([vp, C.POINTER(frb_stats)], C.c_int),
        "set_profiling": ([vp, C.c_int], C.c_int),
        "get_timing": ([vp, C.POINTER(frb_timing)], C.c_int),
        "jit_source": ([vp, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64], C.c_int64),
        "jit_cubin_size": ([vp, C.c_uint32, C.c_uint32], C.c_int64),
        "jit_code_instructions": ([vp, C.c_uint32, C.c_uint32], C.c_int64),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, f"{prefix}_{name}", None)
        if fn is not None:
            fn.argtypes = args
            fn.restype = res


class RendererError(RuntimeError):
    """The reference's renderer methods return () and panic on broken invariants; the C ABI returns a status."""

    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _fptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class CRendererBase:
    """Python mirror of the reference's `Renderer` + `GraphWatcher` traits over a C ABI handle.

    Method names and argument meaning follow reference src/render/renderer.rs:6-17 and
    src/routing/graphwatcher.rs:4-9; an `Edge` is the 4-tuple (from, to, from_slot, to_slot) with handle 0 =
    toplevel (reference src/routing/routegraph.rs:38-44, :330-343).
    """

    _lib = None
    _prefix = None

    def __init__(self, handle):
        if not handle:
            raise RendererError(FRB_E_NO_DEVICE, "renderer creation failed")
        self._h = C.c_void_p(handle)
        self._keep = []

    def _fn(self, name):
        return getattr(self._lib, f"{self._prefix}_{name}")

    def _check(self, rc):
        if rc != 0:
            msg = self._fn("last_error")(self._h)
            raise RendererError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self._fn("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- definitions ----
    def define_effect(self, key, nodes, edges):
        """nodes: [(handle, kind, key)], edges: [(from, to, from_slot, to_slot)] — an EffectData::RouteGraph body."""
        n = (frb_node * max(len(nodes), 1))(*[frb_node(h, k, ky) for (h, k, ky) in nodes])
        e = (frb_edge * max(len(edges), 1))(*[frb_edge(*ed) for ed in edges])
        self._check(self._fn("define_effect")(self._h, key, n, len(nodes), e, len(edges)))

    def define_oscbank(self, key, sample_rate, voice_offsets, freq_hz, amp, phase, attack, tau):
        vo = np.ascontiguousarray(voice_offsets, dtype=np.uint64)
        f = np.ascontiguousarray(freq_hz, dtype=np.float64)
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (amp, phase, attack, tau)]
        d = frb_oscbank_desc(len(vo) - 1, 0, len(f), float(sample_rate), _fptr(vo, C.c_uint64), _fptr(f, C.c_double),
                             *[_fptr(a, C.c_float) for a in arrs])
        self._check(self._fn("define_oscbank")(self._h, key, C.byref(d)))

    def define_directform(self, key, b0, b1, b2, a1, a2):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (b0, b1, b2, a1, a2)]
        d = frb_directform_desc(len(arrs[0]), 0, *[_fptr(a, C.c_float) for a in arrs])
        self._check(self._fn("define_directform")(self._h, key, C.byref(d)))

    def define_fbdelay(self, key, delay, gain):
        dl = np.ascontiguousarray(delay, dtype=np.uint32)
        g = np.ascontiguousarray(gain, dtype=np.float32)
        d = frb_fbdelay_desc(len(dl), 0, _fptr(dl, C.c_uint32), _fptr(g, C.c_float))
        self._check(self._fn("define_fbdelay")(self._h, key, C.byref(d)))

    # ---- GraphWatcher ----
    def on_add_node(self, handle, kind, key=0):
        self._check(self._fn("add_node")(self._h, handle, kind, key))

    def on_del_node(self, handle):
        self._check(self._fn("del_node")(self._h, handle))

    def on_add_edge(self, edge):
        self._check(self._fn("add_edge")(self._h, frb_edge(*edge)))

    def on_del_edge(self, edge):
        self._check(self._fn("del_edge")(self._h, frb_edge(*edge)))

    # ---- Renderer ----
    @staticmethod
    def _jagged(inputs):
        rows = [np.ascontiguousarray(r, dtype=np.float32).ravel() for r in (inputs or [])]
        offs = np.zeros(len(rows) + 1, dtype=np.uint64)
        if rows:
            offs[1:] = np.cumsum([len(r) for r in rows])
        data = np.concatenate(rows) if rows and offs[-1] else np.zeros(1, dtype=np.float32)
        return data, offs, len(rows)

    def fill_buffer(self, n_slots, n_times, idx, inputs=None, out=None):
        """Returns buff[n_slots, n_times] for samples idx..idx+n_times after feeding `inputs` (jagged rows)."""
        data, offs, n_rows = self._jagged(inputs)
        if out is None:
            out = np.zeros((n_slots, n_times), dtype=np.float32)   # Dispatch allocates zeros (dispatch.rs:149)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (n_slots, n_times)
        self._check(self._fn("fill_buffer")(self._h, out.ctypes.data, n_slots, n_times, idx, data.ctypes.data,
                                             _fptr(offs, C.c_uint64), n_rows))
        return out
