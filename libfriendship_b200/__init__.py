"""libfriendship_b200 — B200-native renderer for libfriendship's effect tree.

The product is the CUDA shared library `lib/libfriendship_b200.so` (C ABI: include/friendship_b200.h).  This
package is the thin Python host mirror of the reference's `Renderer`/`GraphWatcher` interface used by tests/ and
bench.py.  There is no CPU fallback: importing fails loudly if the library has not been built, and creating a
renderer fails loudly without a CUDA device.
"""
import ctypes as _C
import os as _os

from . import _cabi
from ._cabi import (  # noqa: F401
    KIND_DELAY, KIND_F32CONSTANT, KIND_SUM2, KIND_MULTIPLY, KIND_DIVIDE, KIND_MODULO, KIND_MINIMUM, KIND_EFFECT,
    KIND_OSCBANK, KIND_DIRECTFORM, KIND_FBDELAY, FLAG_SPARKLE_DELAY, FLAG_NO_JIT, FLAG_JIT_EAGER, FLAG_NO_CHAIN_FUSION, FLAG_NO_EXCITER_FUSION, FLAG_SPARKLE_MIN, FLAG_NO_TENSOR_OSC,
    RendererError,
)

LIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lib", "libfriendship_b200.so")
if not _os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first. "
        "libfriendship_b200 has no CPU fallback.")
_lib = _C.CDLL(LIB_PATH)
_cabi.declare(_lib, "frb")
_lib.frb_create.argtypes = [_C.POINTER(_cabi.frb_config)]
_lib.frb_create.restype = _C.c_void_p
_lib.frb_version.restype = _C.c_char_p
_lib.frb_device_alloc.argtypes = [_C.c_void_p, _C.c_uint64, _C.POINTER(_C.c_void_p)]
_lib.frb_device_free.argtypes = [_C.c_void_p, _C.c_void_p]
_lib.frb_ipc_export.argtypes = [_C.c_void_p, _C.c_void_p, _C.POINTER(_C.c_ubyte)]
_lib.frb_ipc_open.argtypes = [_C.c_void_p, _C.POINTER(_C.c_ubyte), _C.POINTER(_C.c_void_p)]
_lib.frb_ipc_close.argtypes = [_C.c_void_p, _C.c_void_p]
SOURCE_FN = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.POINTER(_C.c_float), _C.c_uint32, _C.c_uint64, _C.c_uint64)
SINK_FN = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.POINTER(_C.c_float), _C.c_uint32, _C.c_uint64, _C.c_uint64)
_lib.frb_render_stream.argtypes = [_C.c_void_p, _C.c_uint32, _C.c_uint64, _C.c_uint64, _C.c_uint64, _C.c_uint32,
                                   SOURCE_FN, SINK_FN, _C.c_void_p]
_lib.frb_sum_rows.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_uint32, _C.c_uint64, _C.c_uint64]


def version():
    return _lib.frb_version().decode()


class B200Renderer(_cabi.CRendererBase):
    """Drop-in for the reference's renderer object (`SparkleRenderer::default()` moved into `Dispatch::new`,
    reference src/dispatch.rs:99-106), running on one B200.  device=-1 creates a planning-only instance (graph
    mirror + schedule dumps); it cannot render."""

    _lib = _lib
    _prefix = "frb"

    def __init__(self, device=0, flags=0, osc_anchor=0, n_devices=0):
        """n_devices > 1: devices device .. device + n_devices - 1 behind this one renderer (voices sharded v mod N)."""
        cfg = _cabi.frb_config(device, flags, osc_anchor, n_devices)
        h = _lib.frb_create(_C.byref(cfg))
        if not h:
            msg = _lib.frb_last_error(None)
            raise RendererError(_cabi.FRB_E_NO_DEVICE, msg.decode() if msg else "frb_create failed")
        super().__init__(h)

    # ---- device-resident path (inputs/outputs stay in HBM) ----
    def fill_buffer_device(self, d_out_ptr, n_slots, n_times, idx, d_in_ptr=0, in_row_offsets=None):
        import numpy as np
        offs = np.ascontiguousarray(in_row_offsets if in_row_offsets is not None else [0], dtype=np.uint64)
        n_rows = len(offs) - 1
        self._check(_lib.frb_fill_buffer_device(self._h, d_out_ptr, n_slots, n_times, idx, d_in_ptr,
                                                offs.ctypes.data_as(_C.POINTER(_C.c_uint64)), n_rows))

    # ---- N4: pipelined block render (pinned double-buffered staging both ways) ----
    def render_stream(self, n_slots, idx, n_total, block, sink, n_in_rows=0, source=None):
        """Renders [idx, idx + n_total) as consecutive fill_buffer calls of `block` samples.
        sink(block_array [n_slots x n], idx): the array is a view of pinned staging memory, valid during the call.
        source(idx, n) -> array-like [n_in_rows x n]: external-input rows of the block starting at idx."""
        import numpy as np
        errors = []

        def c_sink(user, ptr, ns, nt, t):
            try:
                n = ns * nt
                arr = np.ctypeslib.as_array(ptr, shape=(n,)).reshape(ns, nt) if n else np.zeros((ns, nt), np.float32)
                sink(arr, t)
                return 0
            except BaseException as e:      # never unwind through the C frames
                errors.append(e)
                return 1

        def c_source(user, ptr, nr, nt, t):
            try:
                rows = np.asarray(source(t, nt), dtype=np.float32).reshape(nr, nt)
                if nr * nt:
                    np.ctypeslib.as_array(ptr, shape=(nr * nt,))[:] = rows.ravel()
                return 0
            except BaseException as e:
                errors.append(e)
                return 1

        cb_sink = SINK_FN(c_sink)
        cb_source = SOURCE_FN(c_source) if n_in_rows else SOURCE_FN()
        rc = _lib.frb_render_stream(self._h, n_slots, idx, n_total, block, n_in_rows, cb_source, cb_sink, None)
        if errors:
            raise errors[0]
        self._check(rc)

    def sync(self):
        self._check(_lib.frb_sync(self._h))

    def stream(self):
        return _lib.frb_stream(self._h)

    def dump_schedule(self, n_slots):
        import numpy as np
        n = _lib.frb_dump_schedule(self._h, n_slots, None, 0)
        if n < 0:
            self._check(int(n))
        w = np.zeros(n, dtype=np.uint32)
        n2 = _lib.frb_dump_schedule(self._h, n_slots, w.ctypes.data_as(_C.POINTER(_C.c_uint32)), n)
        assert n2 == n
        return w

    def dump_schedule_shard(self, n_slots, rank, world):
        """The schedule device `rank` of `world` would run (n_devices = world): the graph restricted to its voices."""
        import numpy as np
        n = _lib.frb_dump_schedule_shard(self._h, n_slots, rank, world, None, 0)
        if n < 0:
            self._check(int(n))
        w = np.zeros(n, dtype=np.uint32)
        _lib.frb_dump_schedule_shard(self._h, n_slots, rank, world, w.ctypes.data_as(_C.POINTER(_C.c_uint32)), n)
        return w

    def lane_use(self, n_slots):
        """0: outputs linear in the oscillator-bank lanes (shardable), 1: no lane used, 2: otherwise."""
        u = _lib.frb_lane_use(self._h, n_slots)
        if u < 0:
            self._check(int(u))
        return int(u)

    def jit_source(self, n_slots, stage):
        """CUDA source the stage JIT generates for `stage` (see csrc/jit.cc)."""
        n = _lib.frb_jit_source(self._h, n_slots, stage, None, 0)
        if n < 0:
            self._check(int(n))
        buf = _C.create_string_buffer(n + 1)
        _lib.frb_jit_source(self._h, n_slots, stage, buf, n + 1)
        return buf.value.decode()

    def jit_cubin_size(self, n_slots, stage):
        """Compiles the stage with NVRTC for sm_100a (no GPU needed) and returns the cubin size."""
        n = _lib.frb_jit_cubin_size(self._h, n_slots, stage)
        if n < 0:
            self._check(int(n))
        return int(n)

    def jit_code_instructions(self, n_slots, stage):
        """Statements the compiled stage would hold (loops over repeated groups count once per unrolled copy, one body
        per distinct strand structure); the renderer compiles a stage only up to FRB_JIT_MAX_CODE = 4096 of them."""
        n = _lib.frb_jit_code_instructions(self._h, n_slots, stage)
        if n < 0:
            self._check(int(n))
        return int(n)

    # ---- K5: CUDA-IPC plumbing for the cross-GPU mix (see include/friendship_b200.h) ----
    def device_alloc(self, nbytes):
        out = _C.c_void_p()
        self._check(_lib.frb_device_alloc(self._h, nbytes, _C.byref(out)))
        return out.value

    def device_free(self, d_ptr):
        self._check(_lib.frb_device_free(self._h, _C.c_void_p(d_ptr)))

    def ipc_export(self, d_ptr):
        h = (_C.c_ubyte * 64)()
        self._check(_lib.frb_ipc_export(self._h, _C.c_void_p(d_ptr), h))
        return bytes(h)

    def ipc_open(self, handle):
        out = _C.c_void_p()
        h = (_C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(_lib.frb_ipc_open(self._h, h, _C.byref(out)))
        return out.value

    def ipc_close(self, d_ptr):
        self._check(_lib.frb_ipc_close(self._h, _C.c_void_p(d_ptr)))

    def sum_rows(self, d_out, d_rows, n_rows, row_stride, n):
        self._check(_lib.frb_sum_rows(self._h, _C.c_void_p(d_out), _C.c_void_p(d_rows), n_rows, row_stride, n))

    def stats(self):
        s = _cabi.frb_stats()
        self._check(_lib.frb_get_stats(self._h, _C.byref(s)))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def set_profiling(self, on):
        self._check(_lib.frb_set_profiling(self._h, 1 if on else 0))

    def timing(self):
        t = _cabi.frb_timing()
        self._check(_lib.frb_get_timing(self._h, _C.byref(t)))
        return {n: getattr(t, n) for n, _ in t._fields_}
