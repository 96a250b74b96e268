"""The C ABI from plain C: tests/c/abi_smoke.c is compiled with gcc -std=c99 against include/*.h and linked with the
shared library.  CPU: planning-only run (graph ops, validation error, schedule dump, rendering refused).
GPU: the same binary renders the reference's render_delay test through Dispatch."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "libfriendship_b200", "lib")
EXE = os.path.join(ROOT, "build", "abi_smoke")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", EXE, "-L", LIBDIR, "-lfriendship_b200",
                           "-Wl,-rpath," + LIBDIR])


def run(device):
    r = subprocess.run([EXE, str(device)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, r.stdout


def test_headers_compile_as_c99_and_planning_mode_works():
    build()
    run(-1)


@pytest.mark.gpu
def test_c_client_renders_on_the_gpu():
    build()
    run(0)
