"""tests/aux_signals.py -- deterministic test signals for the pre-processor and I/Q generator."""
import numpy as np

N_BLOCK = 128
FS = 44100.0


def _rng(seed):
    return np.random.default_rng(seed)


def pp_case(n_channels, n_blocks, seed=7):
    """Complex IF tones (one strong line each) with small noise.  Channel c % 4: 0 aligned, 1 Q one sample late (needs
    correction +1), 2 I one sample late (the detector walks +1 -> -1), 3 noise only (detector never decides).
    Returns (I, Q) int16 [C, S]."""
    ns = n_blocks * N_BLOCK
    r = _rng(seed)
    t = np.arange(-1, ns)
    I = np.empty((n_channels, ns), np.int16)
    Q = np.empty((n_channels, ns), np.int16)
    for c in range(n_channels):
        f = 2000.0 + 613.0 * (c % 17)
        amp = 6000.0 + 900.0 * (c % 5)
        ph = 2.0 * np.pi * f / FS * t + 0.1 * c
        re, im = amp * np.cos(ph), amp * np.sin(ph)
        kind = c % 4
        if kind == 3:
            re, im = np.zeros_like(re), np.zeros_like(im)
        re = re + r.normal(0.0, 30.0, re.shape)
        im = im + r.normal(0.0, 30.0, im.shape)
        i_sig = re[:-1] if kind == 2 else re[1:]
        q_sig = im[:-1] if kind == 1 else im[1:]
        I[c] = np.clip(np.round(i_sig), -32768, 32767).astype(np.int16)
        Q[c] = np.clip(np.round(q_sig), -32768, 32767).astype(np.int16)
    return I, Q


def pp_events(n_channels, n_blocks):
    """Setter schedule exercising every public function; (channel | None, block, name[, arg])."""
    ev = [(None, 0, "startAutoI2SerrorDetection")]
    for c in range(n_channels):
        k = c % 8
        if k == 4:
            ev.append((c, n_blocks // 3, "swapIQ", 1))
        if k == 5:
            ev.append((c, n_blocks // 2, "setI2SerrorCompensation", -1))
        if k == 6:
            ev.append((c, n_blocks // 4, "stopAutoI2SerrorDetection"))
            ev.append((c, n_blocks // 2, "setI2SerrorCompensation", 1))
            ev.append((c, 3 * n_blocks // 4, "swapIQ", 1))
        if k == 7:
            ev.append((c, n_blocks // 5, "setI2SerrorCompensation", 1))
            ev.append((c, 2 * n_blocks // 5, "startAutoI2SerrorDetection"))
    return ev


def iq_case(n_channels, n_blocks, seed=11):
    """Real audio-band input: two tones + noise, with full-scale clicks (exercises the int16 wrap of the output cast)."""
    ns = n_blocks * N_BLOCK
    r = _rng(seed)
    t = np.arange(ns)
    X = np.empty((n_channels, ns), np.int16)
    for c in range(n_channels):
        f1, f2 = 400.0 + 137.0 * (c % 23), 2500.0 + 91.0 * (c % 31)
        x = 9000.0 * np.cos(2 * np.pi * f1 / FS * t + c) + 5000.0 * np.sin(2 * np.pi * f2 / FS * t) + r.normal(0.0, 200.0, ns)
        clicks = r.integers(0, ns, size=max(1, ns // 3000))
        x[clicks] = np.where(r.random(len(clicks)) < 0.5, 32767.0, -32768.0)
        if c % 7 == 3:
            x = x * 3.4  # drive the rails: clipped input, Hilbert output beyond int16 -> the cast wraps
        X[c] = np.clip(np.round(x), -32768, 32767).astype(np.int16)
    return X


def iq_events(n_channels, n_blocks):
    ev = []
    for c in range(n_channels):
        if c % 3 == 1:
            ev.append((c, 0, "setGainBalance", 1.0 + 0.01 * (c % 10 + 1)))
        if c % 5 == 2:
            ev.append((c, n_blocks // 2, "setGainBalance", 0.93))
    return ev


def block_crcs(a, b):
    """uint32 [C, n_blocks]: CRC-32 of each block's bytes of plane a followed by plane b."""
    import zlib
    nch, ns = a.shape
    out = np.empty((nch, ns // N_BLOCK), np.uint32)
    for c in range(nch):
        for k in range(ns // N_BLOCK):
            z = slice(k * N_BLOCK, (k + 1) * N_BLOCK)
            out[c, k] = zlib.crc32(np.ascontiguousarray(b[c, z]).tobytes(), zlib.crc32(np.ascontiguousarray(a[c, z]).tobytes()))
    return out
