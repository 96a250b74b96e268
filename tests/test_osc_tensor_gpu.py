"""GPU: the matrix-product oscillator kernels (csrc/osc_tc.cuh K1T on tcgen05, csrc/osc_gemm.cuh K1G on mma.sync).  They
render a bank of big voices past its attack ramps; the resonator kernel K1 keeps the ramps and the small banks.  Checked
against the fp64 oracle (extension node: parity unpinned by the reference, tolerance 1e-5 of full scale = sum |amp| of a
voice), against K1 on the same bank, and for independence of how a render is cut into calls (bit-exact).
FRB_OSC_GEMM is read once per process, so the variants run in subprocesses."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.binding import OracleRenderer
from workloads.banks import detuned_bank, full_scale

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5


def bank_with_levels(n_voices, n_partials, seed=5):
    """detuned_bank with voices at very different levels (the fp16 operands are scaled per voice), random phases and
    some partials that do not decay."""
    bank, ids = detuned_bank(n_voices, n_partials, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    amp = bank["amp"].copy()
    for v in range(n_voices):
        lo, hi = bank["voice_offsets"][v], bank["voice_offsets"][v + 1]
        amp[lo:hi] *= (1.0, 1e-3, 40.0)[v % 3]
    bank["amp"] = amp.astype(np.float32)
    bank["phase"] = rng.uniform(0, 6.0, size=amp.shape).astype(np.float32)
    tau = bank["tau"].copy()
    tau[::11] = 0.0
    bank["tau"] = tau.astype(np.float32)
    return bank, ids


def render_voices(r, bank, n_voices, n, blocks=None):
    from libfriendship_b200 import KIND_OSCBANK
    r.define_oscbank(5, **bank)
    r.on_add_node(1, KIND_OSCBANK, 5)
    for v in range(n_voices):
        r.on_add_edge((1, 0, v, v))
    if blocks is None:
        return r.fill_buffer(n_voices, n, 0)
    parts, idx = [], 0
    for m in blocks:
        parts.append(r.fill_buffer(n_voices, m, idx))
        idx += m
    assert idx == n
    return np.concatenate(parts, axis=1)


def voice_full_scales(bank, n_voices):
    return np.array([np.abs(bank["amp"][bank["voice_offsets"][v]:bank["voice_offsets"][v + 1]]).sum() for v in range(n_voices)])


CHILD = r"""
import json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
from libfriendship_b200 import B200Renderer
from test_osc_tensor_gpu import bank_with_levels, render_voices
nv, npart, n = %(nv)d, %(npart)d, %(n)d
bank, _ = bank_with_levels(nv, npart)
r = B200Renderer()
out = render_voices(r, bank, nv, n, %(blocks)r)
np.save(%(path)r, out)
print(json.dumps(r.stats()))
"""


def run_child(tmp_path, mode, nv, npart, n, blocks=None, tag=""):
    path = str(tmp_path / f"out_{mode}{tag}.npy")
    env = dict(os.environ)
    if mode is None:
        env.pop("FRB_OSC_GEMM", None)
    else:
        env["FRB_OSC_GEMM"] = str(mode)
    code = CHILD % dict(root=ROOT, nv=nv, npart=npart, n=n, blocks=blocks, path=path)
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return np.load(path), json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode", [None, 2])
def test_big_voices_against_the_fp64_oracle_and_the_resonator_kernel(tmp_path, mode):
    """8 voices x 1,024 partials: the default policy picks the tensor-core kernel (tcgen05); mode 2 forces mma.sync."""
    nv, npart, n = 8, 1024, 40000
    bank, _ = bank_with_levels(nv, npart)
    out, stats = run_child(tmp_path, mode, nv, npart, n)
    assert stats["osc_tensor_launches"] > 0
    ref, stats0 = run_child(tmp_path, 0, nv, npart, n)
    assert stats0["osc_tensor_launches"] == 0
    fs = voice_full_scales(bank, nv)[:, None]
    assert np.isfinite(out).all()
    # the two kernels agree far inside the tolerance each is held to
    assert (np.abs(out.astype(np.float64) - ref) / fs).max() <= 4e-6
    # fp64 oracle on windows: start (attack ramps: the resonator kernel), the hand-over at t = 384, a tile boundary, the end
    o = OracleRenderer()
    from libfriendship_b200 import KIND_OSCBANK
    o.define_oscbank(5, **bank)
    o.on_add_node(1, KIND_OSCBANK, 5)
    for v in range(nv):
        o.on_add_edge((1, 0, v, v))
    for start in (0, 384 - 16, 16384 - 16, 32768 - 8, n - 32):
        w = o.fill_buffer(nv, 32, start)
        err = (np.abs(out[:, start:start + 32].astype(np.float64) - w) / fs).max()
        assert err <= TOL, (start, err)


def test_values_do_not_depend_on_how_the_render_is_cut(tmp_path):
    nv, npart, n = 4, 512, 36000
    whole, stats = run_child(tmp_path, None, nv, npart, n)
    assert stats["osc_tensor_launches"] > 0
    cut, _ = run_child(tmp_path, None, nv, npart, n, blocks=[512, 16000, 3, 16381, 3104], tag="cut")
    assert whole.tobytes() == cut.tobytes()


def test_small_banks_and_few_voices_stay_on_the_resonator_kernel(tmp_path):
    _, stats = run_child(tmp_path, None, 2, 4096, 20000)           # fewer than 4 voices
    assert stats["osc_tensor_launches"] == 0 and stats["osc_launches"] > 0
    _, stats = run_child(tmp_path, None, 8, 64, 20000, tag="b")    # small voices
    assert stats["osc_tensor_launches"] == 0 and stats["osc_launches"] > 0


def test_forced_on_a_ragged_bank(tmp_path):
    """Mode 3 forces tcgen05 on every bank with the 16-record layout: voices of different sizes, one of them empty-ish."""
    from libfriendship_b200 import KIND_OSCBANK
    sizes = [40, 9, 700, 16]
    rng = np.random.Generator(np.random.PCG64(11))
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    npart = int(off[-1])
    bank = dict(voice_offsets=off, sample_rate=48000.0,
                freq_hz=rng.uniform(20, 23000, npart), amp=rng.uniform(0.01, 1, npart).astype(np.float32),
                phase=rng.uniform(0, 6, npart).astype(np.float32), attack=rng.uniform(0, 300, npart).astype(np.float32),
                tau=rng.uniform(2000, 90000, npart).astype(np.float32))
    code = r"""
import sys, json
sys.path.insert(0, %r)
import numpy as np
from libfriendship_b200 import B200Renderer, KIND_OSCBANK
d = np.load(%r, allow_pickle=True).item()
r = B200Renderer()
r.define_oscbank(5, **d)
r.on_add_node(1, KIND_OSCBANK, 5)
for v in range(4): r.on_add_edge((1, 0, v, v))
np.save(%r, r.fill_buffer(4, 20000, 0))
print(json.dumps(r.stats()))
"""
    bpath, opath = str(tmp_path / "bank.npy"), str(tmp_path / "o.npy")
    np.save(bpath, bank, allow_pickle=True)
    env = dict(os.environ, FRB_OSC_GEMM="3")
    p = subprocess.run([sys.executable, "-c", code % (ROOT, bpath, opath)], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["osc_tensor_launches"] > 0
    out = np.load(opath)
    o = OracleRenderer()
    o.define_oscbank(5, **bank)
    o.on_add_node(1, KIND_OSCBANK, 5)
    for v in range(4):
        o.on_add_edge((1, 0, v, v))
    fs = voice_full_scales(bank, 4)[:, None]
    for start in (0, 300, 16384 - 16, 20000 - 32):
        w = o.fill_buffer(4, 32, start)
        assert (np.abs(out[:, start:start + 32].astype(np.float64) - w) / fs).max() <= TOL, start


def test_no_tensor_flag_keeps_a_big_bank_on_the_resonator_kernel():
    """FRB_FLAG_NO_TENSOR_OSC: a creation flag for callers that stream a big bank in short real-time calls."""
    from libfriendship_b200 import B200Renderer, FLAG_NO_TENSOR_OSC
    nv, npart, n = 4, 512, 20000
    bank, _ = bank_with_levels(nv, npart)
    plain = B200Renderer(flags=FLAG_NO_TENSOR_OSC)
    a = render_voices(plain, bank, nv, n)
    assert plain.stats()["osc_tensor_launches"] == 0 and plain.stats()["osc_launches"] > 0
    if "FRB_OSC_GEMM" not in os.environ:                  # the default policy picks the tensor-core kernel for this bank
        tensor = B200Renderer()
        b = render_voices(tensor, bank, nv, n)
        assert tensor.stats()["osc_tensor_launches"] > 0
        fs = voice_full_scales(bank, nv)[:, None]
        assert (np.abs(a.astype(np.float64) - b) / fs).max() <= 4e-6
