"""tests/signals.py -- synthetic IF signals and setter sequences for the five BASELINE.json configs.

SURVEY.md section 8(d) defines the workloads; this module produces, for any subset of channels and any
block window, the int16 I/Q pair the receiver consumes (q = rint(clip(x, +-1) * 32767)) and the list of
setter events (channel, block, setter, args...) that configures each channel.  Everything is a pure
function of (config, channel index, sample index), so a sampled channel can be regenerated for the oracle
without generating the other 262 143.
"""
import numpy as np

FS = 44100.0
N_BLOCK = 128
LSB, USB, CW_LSB, CW_USB, AM, SAM, WSPR = range(7)
AUDIO_AM, AUDIO_CW, AUDIO_WSPR, AUDIO_2100, AUDIO_2300, AUDIO_2500, AUDIO_2700 = range(7)
AGC_MEDIUM = 2
TUNING = {LSB: 8390.0, USB: 5390.0, CW_LSB: 7390.0, CW_USB: 6390.0, AM: 6890.0, SAM: 6890.0, WSPR: 5390.0}
CONFIG_CHANNELS = {1: 1, 2: 4096, 3: 65536, 4: 16384, 5: 262144}
BLOCKS_10S = 3446
BLOCKS_120S = 41344
_M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def chash(config, channel, salt=0):
    return splitmix64((0x5D120000 + config) ^ ((channel * 0x9E3779B97F4A7C15) & _M64) ^ (salt << 48))


def _unit(h):
    return (h >> 11) / float(1 << 53)


def channel_mode(config, c):
    if config == 1:
        return USB
    if config == 2:
        return [CW_LSB, CW_USB, LSB, USB][chash(2, c) % 4]
    if config == 3:
        return SAM
    if config == 4:
        return chash(4, c) % 7
    if config == 5:
        return WSPR
    raise ValueError(config)


def channel_events(config, c, row):
    """Setter sequence for channel index `c` of `config`, addressed to request row `row`, all at block 0."""
    m = channel_mode(config, c)
    ev = []

    def add(name, *a):
        ev.append((row, 0, name) + tuple(a))

    if config == 1:
        add("setDemodMode", USB); add("disableNoiseBlanker"); add("enableAudioFilter"); add("setAudioFilter", AUDIO_2700)
        add("setInputGain", 1.0); add("setOutputGain", 0.5); add("setMute", 0)
    elif config == 2:
        add("setDemodMode", m); add("enableNoiseBlanker"); add("setNoiseBlankerThresholdDb", 10.0)
        add("setAGCmode", AGC_MEDIUM); add("enableAudioFilter")
        add("setAudioFilter", AUDIO_CW if m in (CW_LSB, CW_USB) else AUDIO_2700); add("setMute", 0)
    elif config == 3:
        add("setDemodMode", SAM); add("disableNoiseBlanker"); add("setAGCmode", AGC_MEDIUM)
        add("enableAudioFilter"); add("setAudioFilter", AUDIO_AM); add("setMute", 0)
    elif config == 4:
        add("setDemodMode", m); add("enableNoiseBlanker"); add("setNoiseBlankerThresholdDb", 10.0)
        add("setAGCmode", AGC_MEDIUM); add("enableAudioFilter")
        filt = {LSB: AUDIO_2700, USB: AUDIO_2700, CW_LSB: AUDIO_CW, CW_USB: AUDIO_CW, AM: AUDIO_AM, SAM: AUDIO_AM,
                WSPR: AUDIO_WSPR}[m]
        add("setAudioFilter", filt); add("enableALSfilter")
        add("setALSfilterNotch" if (chash(4, c, 1) & 1) == 0 else "setALSfilterPeak"); add("setALSfilterAdaptive")
        add("setMute", 0)
    elif config == 5:  # BareBonesWSPR.ino:87-102,129
        add("enableAGC"); add("setAGCmode", AGC_MEDIUM); add("disableALSfilter"); add("disableNoiseBlanker")
        add("setNoiseBlankerThresholdDb", 10.0); add("setInputGain", 1.0); add("setOutputGain", 0.5)
        add("setIQgainBalance", 1.020); add("setAudioFilter", AUDIO_WSPR); add("setDemodMode", WSPR); add("setMute", 0)
    return ev


def _audio_to_if(mode, f_audio):
    """IF frequency that demodulates to `f_audio` in `mode`."""
    if mode in (USB, CW_USB, WSPR):
        return TUNING[mode] + f_audio
    return TUNING[mode] - f_audio


def channel_signal(config, c, n0, n):
    """Complex IF signal samples [n0, n0+n) of channel c (float64 complex), before quantisation."""
    t = np.arange(n0, n0 + n, dtype=np.float64)
    m = channel_mode(config, c)
    w = 2.0 * np.pi / FS
    x = np.zeros(n, np.complex128)
    sigma = 0.01
    if config == 1:
        for fa in (700.0, 1900.0):
            x += 0.2 * np.exp(1j * w * _audio_to_if(USB, fa) * t)
    elif config in (2, 4):
        if m in (LSB, USB):
            f1 = 300.0 + 900.0 * _unit(chash(config, c, 2)); f2 = 1300.0 + 1200.0 * _unit(chash(config, c, 3))
            for fa in (f1, f2):
                x += 0.2 * np.exp(1j * w * _audio_to_if(m, fa) * t)
        elif m in (CW_LSB, CW_USB):
            key = ((t // (FS / 40.0)).astype(np.int64) & 1) == 0  # on/off at 20 Hz
            x += 0.3 * key * np.exp(1j * w * _audio_to_if(m, 700.0) * t)
        elif m in (AM, SAM):
            df = 100.0 * _unit(chash(config, c, 4)) - 50.0
            x += 0.3 * (1.0 + 0.5 * np.cos(w * 1000.0 * t)) * np.exp(1j * w * (6890.0 + df) * t)
        elif m == WSPR:
            x += 0.2 * np.exp(1j * w * _audio_to_if(WSPR, 1500.0) * t)
        if config == 4:  # interferer at audio 1 kHz
            x += 0.1 * np.exp(1j * w * _audio_to_if(m, 1000.0) * t)
        off = chash(config, c, 5) % 11025  # impulses: 3-sample bursts of 0.9 FS on both rails
        ph = (t.astype(np.int64) - off) % 11025
        burst = ph < 3
        x[burst] = 0.9 + 0.9j
    elif config == 3:
        df = 100.0 * _unit(chash(3, c, 4)) - 50.0
        x += 0.3 * (1.0 + 0.5 * np.cos(w * 1000.0 * t)) * np.exp(1j * w * (6890.0 + df) * t)
    elif config == 5:
        sigma = 0.05
        sym = (t // 8192).astype(np.int64)
        tone = np.array([chash(5, c, 8 + int(s)) % 4 for s in np.unique(sym)], np.float64)
        tone = tone[sym - sym.min()]
        # continuous-phase 4-FSK around audio 1500 Hz, spacing 1.4648 Hz (phase approximated per sample)
        f = _audio_to_if(WSPR, 1500.0) + 1.4648 * (tone - 1.5)
        x += 0.05 * np.exp(1j * w * f * t)
    rng = np.random.Generator(np.random.Philox(key=[0x5D120000 + config, c]))
    rng.bit_generator.advance(int(n0) * 4)  # coarse but deterministic for (c, n0)
    noise = rng.standard_normal(2 * n)
    x += sigma * (noise[0::2] + 1j * noise[1::2])
    return x


def quantise(x):
    i = np.rint(np.clip(x.real, -1.0, 1.0) * 32767.0).astype(np.int16)
    q = np.rint(np.clip(x.imag, -1.0, 1.0) * 32767.0).astype(np.int16)
    return i, q


def make(config, channels, n_blocks, start_block=0):
    """Returns (I int16 [len(channels), S], Q, events) for the given channel indices of `config`."""
    ns = n_blocks * N_BLOCK
    I = np.empty((len(channels), ns), np.int16)
    Q = np.empty((len(channels), ns), np.int16)
    events = []
    for row, c in enumerate(channels):
        I[row], Q[row] = quantise(channel_signal(config, int(c), start_block * N_BLOCK, ns))
        events += channel_events(config, int(c), row)
    return I, Q, events


def sample_channels(config, n_total, n_want, n_shards=8):
    """Sampled channel subset: covers every mode and the first/last channel of every shard."""
    picks = set()
    for g in range(n_shards):
        picks.add(n_total * g // n_shards)
        picks.add(n_total * (g + 1) // n_shards - 1)
    seen = set()
    for c in range(min(n_total, 4096)):
        mm = channel_mode(config, c)
        if mm not in seen:
            seen.add(mm); picks.add(c)
    c = 0
    while len(picks) < min(n_want, n_total):
        picks.add(splitmix64(c + 77 * config) % n_total); c += 1
    return sorted(picks)


def sam_lock_unlock_case(n_channels, n_blocks, seed=3):
    """SAM channels driven through lock -> out of lock -> envelope fallback -> re-lock (C:130-143, C:738-747).
    The carrier sits near 6 890 Hz, jumps by +1.5 kHz (outside the lock detector's 5 890..7 890 Hz window) for the second
    fifth of the run, comes back, disappears (noise only) for the fourth fifth and returns.  Returns (I, Q, events)."""
    ns = n_blocks * N_BLOCK
    t = np.arange(ns, dtype=np.float64)
    w = 2.0 * np.pi / FS
    seg = ns // 5
    I = np.empty((n_channels, ns), np.int16); Q = np.empty_like(I)
    ev = []
    for c in range(n_channels):
        df = 100.0 * _unit(chash(3, c, 4 + seed)) - 50.0
        f = np.full(ns, 6890.0 + df)
        f[seg:2 * seg] += 1500.0 + 20.0 * c
        amp = np.full(ns, 0.3)
        amp[3 * seg:4 * seg] = 0.0
        ph = np.cumsum(w * f)
        x = amp * (1.0 + 0.5 * np.cos(w * 1000.0 * t)) * np.exp(1j * ph)
        rng = np.random.Generator(np.random.Philox(key=[0x5A3 + seed, c]))
        noise = rng.standard_normal(2 * ns)
        x = x + 0.01 * (noise[0::2] + 1j * noise[1::2])
        I[c], Q[c] = quantise(x)
        ev += channel_events(3, c, c)
        if c % 3 == 1:  # some channels with the blanker on as well (the general ENV launch instead of the lean one)
            ev += [(c, 0, "enableNoiseBlanker"), (c, 0, "setNoiseBlankerThresholdDb", 10.0)]
    return I, Q, ev
