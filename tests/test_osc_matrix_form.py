"""CPU: the mathematics of the tensor-core oscillator kernels (csrc/osc_gemm.cuh, csrc/osc_tc.cuh), restated in numpy and
checked against the fp64 closed form.  No CUDA here — this pins the derivation and the precision budget the kernels rely on:

  amp rho^t sin(w t + phi), t = 128 b + j  =  [E_b sin T_b] [rho^j cos w j] + [E_b cos T_b] [rho^j sin w j]
  operands fp16 hi + lo (22 bits), A scaled per voice into [2^13, 2^14), W by 2^10, products lo*hi + hi*lo + hi*hi summed in fp32,
  rows / columns by complex rotation (8 steps) from anchors.

The GPU kernels themselves are checked in tests/test_osc_tensor_gpu.py."""
import numpy as np

from workloads.banks import detuned_bank

N_BLK = 128           # samples per block (GEMM N)
WSCALE = 1024.0


def split16(x):
    """x (fp32) -> fp16 hi, lo with hi + lo = x to 22 bits (what gm_split does)."""
    x = x.astype(np.float32)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def voice_operands(freq, amp, phase, tau, sr, b0, n_blocks, rotate=True):
    """A [n_blocks, 2P] and W [2P, 128] of one voice past its attack ramps, in fp32, built like the kernels build them:
    anchors every 8 rows / columns (fp64 phase reduced exactly, then fp32), the 7 in between by fp32 complex rotation."""
    fr = np.mod(freq / sr, 1.0)
    rho_log2 = np.where(tau > 0, -np.log2(np.e) / np.where(tau > 0, tau, 1.0), 0.0)     # rho^n = 2^(rho_log2 n)

    def anchor(n, ph0, a):
        turns = np.mod(fr * n + ph0 / (2 * np.pi), 1.0)
        e = (a * np.exp2(rho_log2 * n)).astype(np.float32)
        return (e * np.sin(2 * np.pi * turns).astype(np.float32)).astype(np.float32), \
               (e * np.cos(2 * np.pi * turns).astype(np.float32)).astype(np.float32)

    def rot(step):
        ang = 2 * np.pi * np.mod(fr * step, 1.0)
        r = np.exp2(rho_log2 * step)
        return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)

    P = len(freq)
    A = np.zeros((n_blocks, 2 * P), np.float32)
    W = np.zeros((2 * P, N_BLK), np.float32)
    scale = 2.0 ** (14 - np.frexp(np.abs(amp).max())[1])
    cr, ci = rot(8 * N_BLK)
    dr, di = rot(8)
    for g in range(8):
        # rows g, g + 8, ... of every 64-row group: anchor at the group's first, then rotate
        for base in range(0, n_blocks, 64):
            s, c = anchor((b0 + base + g) * float(N_BLK), phase, amp * scale)
            for i in range(8):
                r = base + g + 8 * i
                if r < n_blocks:
                    A[r, 0::2], A[r, 1::2] = s, c
                if rotate:
                    s, c = (s * cr + c * ci).astype(np.float32), (c * cr - s * ci).astype(np.float32)
                else:
                    s, c = anchor((b0 + r + 8) * float(N_BLK), phase, amp * scale)
        for base in range(0, N_BLK, 64):
            s, c = anchor(float(base + g), 0.0, np.full(P, WSCALE))
            for i in range(8):
                j = base + g + 8 * i
                W[0::2, j], W[1::2, j] = c, s
                c, s = (c * dr - s * di).astype(np.float32), (s * dr + c * di).astype(np.float32)
    return A, W, scale


def product(A, W, fmt, chunk_partials=64):
    """out = A W with the operands rounded the way `fmt` says; fp32 sums, one chunk of partials at a time."""
    out = np.zeros((A.shape[0], W.shape[1]), np.float32)
    for k0 in range(0, A.shape[1], 2 * chunk_partials):
        a, w = A[:, k0:k0 + 2 * chunk_partials], W[k0:k0 + 2 * chunk_partials]
        if fmt == "fp16x3":
            ah, al = split16(a)
            wh, wl = split16(w)
            f = lambda x: x.astype(np.float32)
            part = f(al) @ f(wh) + f(ah) @ f(wl) + f(ah) @ f(wh)
        elif fmt == "fp16x1":
            part = a.astype(np.float16).astype(np.float32) @ w.astype(np.float16).astype(np.float32)
        else:
            part = a @ w
        out += part.astype(np.float32)
    return out


def closed_form(freq, amp, phase, tau, sr, t):
    env = np.where(tau[:, None] > 0, np.exp(-t[None, :] / np.where(tau > 0, tau, 1.0)[:, None]), 1.0)
    turns = np.mod(np.mod(freq / sr, 1.0)[:, None] * t[None, :], 1.0)
    return (amp[:, None] * env * np.sin(2 * np.pi * turns + phase[:, None])).sum(axis=0)


def one_voice(n_partials=512, level=1.0, seed=3):
    bank, _ = detuned_bank(1, n_partials, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    freq = bank["freq_hz"].astype(np.float64)
    amp = bank["amp"].astype(np.float64) * level
    phase = rng.uniform(0, 6.0, n_partials)
    tau = bank["tau"].astype(np.float64)
    tau[::11] = 0.0
    return freq, amp, phase, tau, float(bank["sample_rate"])


def render(freq, amp, phase, tau, sr, b0, n_blocks, fmt, rotate=True):
    A, W, scale = voice_operands(freq, amp, phase, tau, sr, b0, n_blocks, rotate)
    return product(A, W, fmt).astype(np.float64).reshape(-1) / (scale * WSCALE)


def test_matrix_form_matches_the_closed_form():
    for level in (1.0, 1e-3, 40.0):                       # the per-voice scale keeps fp16 in range at any level
        freq, amp, phase, tau, sr = one_voice(level=level)
        b0, nb = 3, 128                                     # blocks 3 .. 130: past the attack ramps (t >= 384)
        t = (b0 * N_BLK + np.arange(nb * N_BLK)).astype(np.float64)
        ref = closed_form(freq, amp, phase, tau, sr, t)
        fs = np.abs(amp).sum()
        err = np.abs(render(freq, amp, phase, tau, sr, b0, nb, "fp16x3") - ref).max() / fs
        assert err <= 1e-6, (level, err)
        # the fp32 product of the same operands is no better: the error left is the operands' (anchors, rotations), not the split's
        err32 = np.abs(render(freq, amp, phase, tau, sr, b0, nb, "fp32") - ref).max() / fs
        assert err <= 2 * err32 + 2e-7, (level, err, err32)


def test_one_fp16_pass_is_not_enough_and_the_scale_matters():
    freq, amp, phase, tau, sr = one_voice()
    b0, nb = 3, 64
    t = (b0 * N_BLK + np.arange(nb * N_BLK)).astype(np.float64)
    ref = closed_form(freq, amp, phase, tau, sr, t)
    fs = np.abs(amp).sum()
    one = np.abs(render(freq, amp, phase, tau, sr, b0, nb, "fp16x1") - ref).max() / fs
    three = np.abs(render(freq, amp, phase, tau, sr, b0, nb, "fp16x3") - ref).max() / fs
    assert one > 1e-5 > 30 * three                         # a single pass misses the bar the kernels are held to
    # without the per-voice scale a quiet voice falls into fp16's subnormals: 2^-24 spacing against amplitudes of 1e-7
    fq, am, ph, ta, _ = one_voice(level=1e-4)
    A, W, scale = voice_operands(fq, am, ph, ta, sr, b0, nb)
    unscaled = product(A / np.float32(scale), W / np.float32(WSCALE), "fp16x3").astype(np.float64).reshape(-1)
    ref_q = closed_form(fq, am, ph, ta, sr, t)
    assert np.abs(unscaled - ref_q).max() / np.abs(am).sum() > 1e-4


def test_rotation_between_anchors_costs_nothing_measurable():
    freq, amp, phase, tau, sr = one_voice(n_partials=256)
    b0, nb = 3, 64
    t = (b0 * N_BLK + np.arange(nb * N_BLK)).astype(np.float64)
    ref = closed_form(freq, amp, phase, tau, sr, t)
    fs = np.abs(amp).sum()
    rotated = np.abs(render(freq, amp, phase, tau, sr, b0, nb, "fp16x3", rotate=True) - ref).max() / fs
    direct = np.abs(render(freq, amp, phase, tau, sr, b0, nb, "fp16x3", rotate=False) - ref).max() / fs
    assert rotated <= direct + 3e-7 and rotated <= 1e-6
