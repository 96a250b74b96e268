"""Python mirror of the reference's `Dispatch` (src/dispatch.rs) over the C ABI of include/friendship_dispatch.h.

`Dispatch(client)` owns the RouteGraph, the ResMan and the B200 renderer exactly as the reference's struct does;
every method is one OSC message of src/dispatch.rs:31-86.  RouteGraph / effect errors surface as DispatchError with
the reference's variant names."""
import ctypes as C
import json

import numpy as np

from . import _cabi, _lib

PRIMITIVES = ("Delay", "F32Constant", "Sum2", "Multiply", "Divide", "Modulo", "Minimum")

ERRORS = {-101: "WouldCycle", -102: "NodeInUse", -103: "NodeExists", -104: "SlotAlreadyConnected", -105: "NoSuchNode",
          -106: "NoSuchSlot", -107: "NoMatchingEffect", -108: "BadMessage"}


class DispatchError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {ERRORS.get(code, 'RendererError')}: {msg}")
        self.code = code
        self.variant = ERRORS.get(code, "RendererError")


class EffectId:
    """reference src/routing/effect.rs:26-39"""

    def __init__(self, name, sha256=None, urls=()):
        self.name, self.sha256, self.urls = name, (bytes(sha256) if sha256 is not None else None), list(urls)

    @staticmethod
    def primitive(name):
        """EffectId::new(name, None, [primitive:///name]) as the reference's tests build them (tests/render_prim.rs:35-67)"""
        assert name in PRIMITIVES
        return EffectId(name, None, [f"primitive:///{name}"])

    def to_json(self):
        return json.dumps({"name": self.name, "sha256": None if self.sha256 is None else list(self.sha256),
                           "urls": self.urls})


_AUDIO = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint64, C.c_uint64)
_NODEJSON = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_char_p)


class _frd_client(C.Structure):
    _fields_ = [("user", C.c_void_p), ("audio_rendered", _AUDIO), ("node_meta", _NODEJSON), ("node_id", _NODEJSON)]


_lib.frd_create.argtypes = [C.POINTER(_cabi.frb_config), C.POINTER(_frd_client)]
_lib.frd_create.restype = C.c_void_p
_lib.frd_destroy.argtypes = [C.c_void_p]
_lib.frd_last_error.argtypes = [C.c_void_p]
_lib.frd_last_error.restype = C.c_char_p
_lib.frd_renderer.argtypes = [C.c_void_p]
_lib.frd_renderer.restype = C.c_void_p
_lib.frd_add_node.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
_lib.frd_add_edge.argtypes = [C.c_void_p, _cabi.frb_edge]
_lib.frd_del_node.argtypes = [C.c_void_p, C.c_uint32]
_lib.frd_del_edge.argtypes = [C.c_void_p, _cabi.frb_edge]
_lib.frd_query_meta.argtypes = [C.c_void_p, C.c_uint32]
_lib.frd_query_id.argtypes = [C.c_void_p, C.c_uint32]
_lib.frd_render_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32]
_lib.frd_render_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
_lib.frd_wav_open.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
_lib.frd_wav_open.restype = C.c_void_p
_lib.frd_wav_write.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
_lib.frd_wav_close.argtypes = [C.c_void_p]
_lib.frd_wav_error.argtypes = [C.c_void_p]
_lib.frd_wav_error.restype = C.c_char_p
_lib.frd_add_dir.argtypes = [C.c_void_p, C.c_char_p]
_lib.frd_sha256_file.argtypes = [C.c_char_p, C.c_char_p]
_lib.frd_adjlist_json.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
_lib.frd_adjlist_json.restype = C.c_int64


def sha256_file(path):
    buf = C.create_string_buffer(32)
    if _lib.frd_sha256_file(str(path).encode(), buf) != 0:
        raise OSError(f"cannot read {path}")
    return buf.raw


class Client:
    """reference src/client/client.rs:8-15: every callback defaults to a no-op."""

    def audio_rendered(self, buffer, idx):
        pass

    def node_meta(self, handle, meta):
        pass

    def node_id(self, handle, id):
        pass


class WavClient(Client):
    """N4 output sink: a Client whose audio_rendered appends every buffer to a float32 WAV file (slot s = channel s;
    the C writer of include/friendship_dispatch.h).  The reference leaves file output to the client (README.md:22-26)."""

    def __init__(self, path, n_channels, sample_rate):
        self._w = _lib.frd_wav_open(str(path).encode(), n_channels, int(sample_rate))
        if not self._w:
            raise OSError(f"cannot create {path}")
        self.frames = 0

    def audio_rendered(self, buffer, idx):
        buffer = np.ascontiguousarray(buffer, dtype=np.float32)
        if _lib.frd_wav_write(self._w, buffer.ctypes.data, buffer.shape[0], buffer.shape[1]) != 0:
            raise OSError(_lib.frd_wav_error(self._w).decode())
        self.frames += buffer.shape[1]

    def close(self):
        if self._w:
            w, self._w = self._w, None
            if _lib.frd_wav_close(w) != 0:
                raise OSError("wav: close failed")


class Dispatch:
    def __init__(self, client=None, device=0, flags=0, n_devices=0):
        """n_devices > 1: the Dispatch's ONE renderer drives devices device .. device + n_devices - 1 (csrc/multi.cu)."""
        self.client = client or Client()
        self._cb_audio = _AUDIO(self._on_audio)
        self._cb_meta = _NODEJSON(lambda u, h, s: self.client.node_meta(h, json.loads(s.decode())))
        self._cb_id = _NODEJSON(lambda u, h, s: self.client.node_id(h, json.loads(s.decode())))
        self._cs = _frd_client(None, self._cb_audio, self._cb_meta, self._cb_id)
        cfg = _cabi.frb_config(device, flags, 0, n_devices)
        self._h = _lib.frd_create(C.byref(cfg), C.byref(self._cs))
        if not self._h:
            msg = _lib.frb_last_error(None)
            raise _cabi.RendererError(_cabi.FRB_E_NO_DEVICE, msg.decode() if msg else "frd_create failed")

    def _on_audio(self, user, buf, n_slots, n_times, idx):
        n = n_slots * n_times
        arr = np.ctypeslib.as_array(buf, shape=(n,)).reshape(n_slots, n_times).copy() if n else np.zeros((n_slots, n_times), np.float32)
        self.client.audio_rendered(arr, idx)

    def _check(self, rc):
        if rc != 0:
            msg = _lib.frd_last_error(self._h)
            raise DispatchError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            _lib.frd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- OscToplevel: the one entry point of the reference, `Dispatch::dispatch(msg)` (dispatch.rs:109-160) ----
    _ROUTES = {
        "/routegraph/add_node": "add_node", "/routegraph/add_edge": "add_edge", "/routegraph/del_node": "del_node",
        "/routegraph/del_edge": "del_edge", "/routegraph/query_meta": "query_meta", "/routegraph/query_id": "query_id",
        "/renderer/render": "_render_msg", "/resman/add_dir": "add_dir",
    }

    def dispatch(self, address, *args):
        """One OSC message by its address — the `#[osc_address]` paths of `OscToplevel` / `OscRouteGraph` /
        `OscRenderer` / `OscResMan` (dispatch.rs:31-84) — with the message's argument tuple:
        `/routegraph/add_node (handle, EffectId)`, `/routegraph/add_edge (edge,)`, `/routegraph/del_node (handle,)`,
        `/routegraph/del_edge (edge,)`, `/routegraph/query_meta (handle,)`, `/routegraph/query_id (handle,)`,
        `/renderer/render (range, num_slots, inputs)`, `/resman/add_dir (dir,)`.  The byte encoding of OSC packets
        (the reference derives it with an external macro crate) is not part of this layer."""
        name = self._ROUTES.get("/" + "/".join(p for p in address.split("/") if p))
        if name is None:
            raise DispatchError(-108, f"no such OSC address: {address}")
        return getattr(self, name)(*args)

    def _render_msg(self, rng, n_slots, inputs=None):
        start, end = (rng.start, rng.stop) if isinstance(rng, range) else rng
        return self.render_range(start, end, n_slots, inputs)

    # ---- OscRouteGraph ----
    def add_node(self, handle, effect_id):
        self._check(_lib.frd_add_node(self._h, handle, effect_id.to_json().encode()))

    def add_edge(self, edge):
        self._check(_lib.frd_add_edge(self._h, _cabi.frb_edge(*edge)))

    def del_node(self, handle):
        self._check(_lib.frd_del_node(self._h, handle))

    def del_edge(self, edge):
        self._check(_lib.frd_del_edge(self._h, _cabi.frb_edge(*edge)))

    def query_meta(self, handle):
        self._check(_lib.frd_query_meta(self._h, handle))

    def query_id(self, handle):
        self._check(_lib.frd_query_id(self._h, handle))

    # ---- OscRenderer ----
    def render_range(self, start, end, n_slots, inputs=None):
        data, offs, n_rows = _cabi.CRendererBase._jagged(inputs)
        self._check(_lib.frd_render_range(self._h, start, end, n_slots, data.ctypes.data,
                                          offs.ctypes.data_as(C.POINTER(C.c_uint64)), n_rows))

    def render_stream(self, start, end, n_slots, block):
        """N4: [start, end) as consecutive audio_rendered callbacks of `block` samples, render / copy / client
        overlapped (include/friendship_dispatch.h)."""
        self._check(_lib.frd_render_stream(self._h, start, end, n_slots, block))

    # ---- OscResMan ----
    def add_dir(self, path):
        self._check(_lib.frd_add_dir(self._h, str(path).encode()))

    def adjlist(self):
        n = _lib.frd_adjlist_json(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        _lib.frd_adjlist_json(self._h, buf, n + 1)
        return json.loads(buf.value.decode())
