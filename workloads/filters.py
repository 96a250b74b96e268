"""cfg3 building blocks: RBJ low-pass biquads + feedback delays per voice."""
import numpy as np


def rbj_lowpass(fc, q, sr=48000.0):
    w0 = 2.0 * np.pi * np.asarray(fc, dtype=np.float64) / sr
    alpha = np.sin(w0) / (2.0 * np.asarray(q, dtype=np.float64))
    cw = np.cos(w0)
    a0 = 1.0 + alpha
    b0 = (1.0 - cw) / 2.0 / a0
    b1 = (1.0 - cw) / a0
    b2 = b0
    a1 = -2.0 * cw / a0
    a2 = (1.0 - alpha) / a0
    f = lambda x: np.asarray(x, dtype=np.float32)
    return f(b0), f(b1), f(b2), f(a1), f(a2)


def cfg3_filters(n_voices, sr=48000.0):
    """fc log-spaced 200 Hz - 8 kHz by voice, Q from 0.707 to 4; D = 100 + v mod 900, g = 0.7."""
    v = np.arange(n_voices)
    frac = v / max(n_voices - 1, 1)
    fc = 200.0 * (8000.0 / 200.0) ** frac
    q = 0.707 + (4.0 - 0.707) * ((v * 7) % 11) / 10.0
    delay = (100 + v % 900).astype(np.uint32)
    gain = np.full(n_voices, 0.7, dtype=np.float32)
    return rbj_lowpass(fc, q, sr), delay, gain


def build_cfg3_graph(r, n_voices, excitation="input", bank=None, mix_to_one=False):
    """excitation 'input': external input slot v feeds voice v; 'osc': OscBank voice v.  Each voice:
    excitation -> DirectForm (biquad) -> FbDelay -> output slot v (or a Sum2 chain to slot 0)."""
    from .kinds import KIND_DIRECTFORM, KIND_FBDELAY, KIND_OSCBANK, KIND_SUM2
    (b0, b1, b2, a1, a2), delay, gain = cfg3_filters(n_voices)
    r.define_directform(11, b0, b1, b2, a1, a2)
    r.define_fbdelay(12, delay, gain)
    H_DF, H_FB, H_OSC = 10, 11, 12
    r.on_add_node(H_DF, KIND_DIRECTFORM, 11)
    r.on_add_node(H_FB, KIND_FBDELAY, 12)
    if excitation == "osc":
        r.define_oscbank(13, **bank)
        r.on_add_node(H_OSC, KIND_OSCBANK, 13)
    for v in range(n_voices):
        src = (0, v) if excitation == "input" else (H_OSC, v)
        r.on_add_edge((src[0], H_DF, src[1], v))
        r.on_add_edge((H_DF, H_FB, v, v))
    if not mix_to_one:
        for v in range(n_voices):
            r.on_add_edge((H_FB, 0, v, v))
        return n_voices
    h = 100
    total = (H_FB, 0)
    for v in range(1, n_voices):
        r.on_add_node(h, KIND_SUM2)
        r.on_add_edge((total[0], h, total[1], 0))
        r.on_add_edge((H_FB, h, v, 1))
        total = (h, 0)
        h += 1
    r.on_add_edge((total[0], 0, total[1], 0))
    return 1
