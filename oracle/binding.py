"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) — TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package.

Same Python surface as libfriendship_b200.B200Renderer so parity tests drive both with the same calls.
"""
import ctypes as C
import os
import sys

import importlib.util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The oracle's C shim mirrors include/friendship_b200.h (orc_* for frb_*), so it shares the ctypes declarations of that
# header — loaded BY PATH: importing the package would map the product library, and bench.py's reference arm must run
# on oracle/_build/liboracle.so alone.
if "libfriendship_b200._cabi" in sys.modules:
    # a test process that drives both renderers: one set of classes (RendererError raised by either is the same type)
    _cabi = sys.modules["libfriendship_b200._cabi"]
else:
    _spec = importlib.util.spec_from_file_location("_frb_cabi_decls", os.path.join(ROOT, "libfriendship_b200", "_cabi.py"))
    _cabi = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_cabi)
_LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_lib = C.CDLL(_LIB_PATH)
_cabi.declare(_lib, "orc")
_lib.orc_create.restype = C.c_void_p
_lib.orc_set_ext_mode.argtypes = [C.c_void_p, C.c_int]
_lib.orc_fill_buffer_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32,
                                    C.POINTER(C.c_double)]
_lib.orc_fill_buffer_mt.restype = C.c_int


class OracleRenderer(_cabi.CRendererBase):
    _lib = _lib
    _prefix = "orc"

    def __init__(self, ext_mode="fp64", flags=0, **_ignored):
        super().__init__(_lib.orc_create())
        _lib.orc_set_ext_mode(self._h, 0 if ext_mode == "fp64" else 1)
        _lib.orc_set_sparkle_delay.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_set_sparkle_delay(self._h, 1 if flags & _cabi.FLAG_SPARKLE_DELAY else 0)
        _lib.orc_set_sparkle_min.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_set_sparkle_min(self._h, 1 if flags & _cabi.FLAG_SPARKLE_MIN else 0)

    def fill_buffer_mt(self, n_slots, n_times, idx, n_threads):
        """CPU-baseline helper: multi-threaded evaluation of the current graph (inputs already fed)."""
        import numpy as np
        out = np.zeros((n_slots, n_times), dtype=np.float32)
        sec = C.c_double(0)
        self._check(_lib.orc_fill_buffer_mt(self._h, out.ctypes.data, n_slots, n_times, idx, n_threads, C.byref(sec)))
        return out, sec.value
