import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")
    # Build what the tests need if it is not there yet (the GPU box receives the prebuilt .so files).
    lib = os.path.join(ROOT, "libfriendship_b200", "lib", "libfriendship_b200.so")
    orc = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "__graft_entry__.py")])
    # The product package first: oracle/binding.py then shares its ABI declaration module (one RendererError class for
    # both renderers) instead of loading a private copy by path, as it does for bench.py's reference arm.
    import libfriendship_b200  # noqa: F401


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # no test may hang a GPU box (or the CPU suite): 10 minutes each at the very most (pytest-timeout, when installed)
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if not any(m.name == "timeout" for m in item.iter_markers()):
                item.add_marker(pytest.mark.timeout(600))
    # `-m gpu` on a box without a GPU (or vice versa) should skip, not fail.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
