"""CPU: bench.py / tools call into the package with signatures that exist (no GPU needed to catch a renamed or
shadowed method), and the bench line's static parts follow the driver's contract."""
import ast
import inspect
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _calls_on(tree, base):
    """(method, n_positional, keywords) for every call of the form <base>.<method>(...), base like 'sr' or 'sr.r'."""
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and ast.unparse(node.func.value) == base:
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue
            kws = [k.arg for k in node.keywords if k.arg is not None]
            has_splat = any(k.arg is None for k in node.keywords)
            out.append((node.func.attr, len(node.args), kws, has_splat))
    return out


def _check(cls, calls, where):
    for name, n_pos, kws, has_splat in calls:
        fn = getattr(cls, name, None)
        assert fn is not None, f"{where}: {cls.__name__}.{name} does not exist"
        sig = inspect.signature(fn)
        if has_splat:
            continue
        try:
            sig.bind(None, *([0] * n_pos), **{k: 0 for k in kws})
        except TypeError as e:
            raise AssertionError(f"{where}: {cls.__name__}.{name}{sig} called with {n_pos} positional + {kws}: {e}")


def test_bench_and_tools_call_existing_methods():
    from libfriendship_b200 import B200Renderer
    from libfriendship_b200.sharded import ShardedRenderer
    for rel in ("bench.py", "tools/render_cfg5.py"):
        tree = ast.parse(open(os.path.join(ROOT, rel)).read())
        sr_calls = _calls_on(tree, "sr")
        assert sr_calls, rel
        _check(ShardedRenderer, sr_calls, rel)
        _check(B200Renderer, _calls_on(tree, "sr.r"), rel)


def test_bench_line_static_contract():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"e2e"', '"h2d_bytes_per_step"',
                '"d2h_bytes_per_step"', '"gpu_launches"', '"clocks"', '"roofline"', '"cpu_baseline"', '"impl"'):
        assert key in src, key
    assert "/root/reference" not in src


def test_reference_arm_prints_exactly_one_json_line():
    """`bench.py --impl reference` runs on host cores only (no GPU): stdout must be ONE JSON line with the contract's
    keys, whatever libraries print (stdout is reserved for it)."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, NCCL_DEBUG="VERSION")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "partial-samples/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cfg4" in d["config"]["workload"]


def test_reference_arm_never_maps_the_product_library():
    """VERDICT r1 'weak' 8: the reference arm's imports (workloads/, oracle/binding.py) must not load libfriendship_b200.so —
    the oracle binding shares the ctypes declarations of the C header by loading that one file by path."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import bench\n"
            "from oracle.binding import OracleRenderer\n"
            "from workloads.banks import build_voice_mix_graph, detuned_bank\n"
            "bank, ids = detuned_bank(1, 8)\n"
            "r = OracleRenderer(ext_mode='f32'); build_voice_mix_graph(r, bank, ids); r.fill_buffer(1, 8, 0)\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'liboracle.so' in maps, 'oracle not loaded'\n"
            "assert 'libfriendship_b200.so' not in maps, 'product library mapped by the reference arm'\n"
            "assert 'libfriendship_b200' not in sys.modules\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]


def test_parity_windows_cover_start_taps_middle_and_end():
    sys.path.insert(0, ROOT)
    import bench
    wins = bench.parity_windows(480000, 64)
    assert len(wins) >= 4
    starts = [w[0] for w in wins]
    assert 0 in starts and any(4790 <= s <= 4800 for s in starts) and any(7120 <= s <= 7131 for s in starts)
    assert any(s == 240000 for s in starts) and wins[-1] == (480000 - 32, 32)
    for s0, n in wins:
        assert 0 <= s0 and s0 + n <= 480000
    assert all(s0 + n <= 600 for s0, n in bench.parity_windows(600, 4))      # a debug-sized run stays in range
