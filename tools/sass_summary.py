"""Static SASS mnemonic counts per kernel of the built library (no GPU needed):
python tools/sass_summary.py > profiles/sass_summary_rN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "libfriendship_b200", "lib", "libfriendship_b200.so")
KEYS = ["FFMA", "FADD", "FMUL", "LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "SHFL", "MUFU", "BAR", "HMMA", "UTCHMMA", "UTCBAR", "LDTM", "F2FP"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    rows = []
    for p in re.split(r"\n\s*Function : ", txt)[1:]:
        name = p.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        i = dem.find(">(")
        dem = dem[:i + 1] if i >= 0 else dem.split("(")[0]
        for junk in ("void ", "frb::", "(bool)", "(int)", " "):
            dem = dem.replace(junk, "")
        ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)([^;]*);", p)
        c = collections.Counter(m.split(".")[0] for m, _ in ins)
        ffma_u = sum(1 for m, a in ins if m.startswith("FFMA") and re.search(r"\bUR\d", a))
        rows.append((dem, len(ins), c, ffma_u))
    w = sys.stdout.write
    w("SASS mnemonic counts per kernel of the shipped libfriendship_b200.so (cuobjdump -sass, sm_100a; static instruction counts;\n")
    w("tools/sass_summary.py).  FFMA.u = FFMAs with a uniform-register operand.\n")
    w("osc_kernel<P,ATTACK>: P partials per thread; dfcomb_kernel<EXC,BULK>: BULK 1/2 = x tile / x tile + tap window by UBLKCP + mbarrier\n")
    w("(built for the A/B of profiles/k4_bulk_ab_r2.jsonl; the default is BULK = 0, LDGSTS).  The stage JIT's kernels are generated at\n")
    w("run time (NVRTC) and are not in this file: FRB_JIT_DUMP=<dir> writes their CUDA source.\n")
    w("osc_tc_kernel = K1T (tcgen05: UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, SYNCS = mbarrier);\n")
    w("osc_gemm_kernel = K1G (mma.sync: HMMA.16816.F32); F2FP = the fp32 -> 2 x fp16 conversions of the operand split.\n\n")
    w("%-34s %6s %6s " % ("kernel", "instrs", "FFMA.u") + " ".join("%7s" % k for k in KEYS) + "\n")
    for dem, n, c, fu in sorted(rows):
        w("%-34s %6d %6d " % (dem[:34], n, fu) + " ".join("%7d" % c.get(k, 0) for k in KEYS) + "\n")


if __name__ == "__main__":
    main()
