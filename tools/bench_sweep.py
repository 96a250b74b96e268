"""Runs bench.py with a list of extra-argument variants and prints one compact line per run (GPU box helper)."""
import json
import subprocess
import sys

variants = sys.argv[1:] or [""]
for v in variants:
    cmd = [sys.executable, "bench.py", "--steps", "2", "--warmup", "1", "--no-cpu-baseline"] + v.split()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"[{v}] value={d['value']:.4e} ms/step={d['ms_per_step']:.1f} e2e={d['e2e']['value']:.4e} "
              f"exec_frac={d['roofline']['executed_frac']:.3f} alg_frac={d['roofline']['frac']:.3f} clocks={d['clocks']['sm_mhz']}", flush=True)
    except Exception as e:
        print(f"[{v}] FAILED {e}: {r.stderr[-500:]}", flush=True)
