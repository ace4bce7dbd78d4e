"""oracle/ref_client.py -- TEST INFRASTRUCTURE (not product code).

Python client for oracle/_ref/refsdr, the unmodified reference AudioSDR.{h,cpp}
compiled on the host (see oracle/ref_driver.cpp).  Writes the request file,
runs the binary (one forked process per channel), parses the response.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFSDR = os.path.join(HERE, "_ref", "refsdr")
ALL = 0xFFFFFFFF
N_BLOCK = 128

# Setter opcodes: same numbering as include/sdr_batch.h (enum sdr_setter) and oracle/sdr_oracle.h.
OPS = dict(
    setMute=1, setInputGain=2, setIQgainBalance=3, setDemodMode=4, enableAudioFilter=5,
    disableAudioFilter=6, setOutputGain=7, setAudioFilter=8, enableALSfilter=9, disableALSfilter=10,
    setALSfilterNotch=11, setALSfilterPeak=12, setALSfilterAdaptive=13, setALSfilterStatic=14,
    setALSfilterParams=15, enableAGC=16, disableAGC=17, setAGCthreshold=18, setAGCslope=19, setAGCmode=20,
    setAGCkneeWidth=21, setAGCattackTime=22, setAGCreleaseTime=23, setAGChangTime=24, setAGCstaticGain=25,
    enableNoiseBlanker=26, disableNoiseBlanker=27, setNoiseBlankerThreshold=28,
    setNoiseBlankerThresholdDb=29, init=30, oracle_identity_IF=100,
)
STATUS_FIELDS = ["tuning_offset", "mode", "agc_active", "nb_detected", "sam_freq", "sam_locked",
                 "am_carrier", "bpf_lower", "bpf_upper", "muted", "audio_filter", "agc_enabled",
                 "nb_enabled", "als_enabled", "agc_gain", "nb_avg"]


def available():
    return os.access(REFSDR, os.X_OK)


def pack_events(events):
    """events: iterable of (channel, block, opname|opcode, a0, a1, a2) with trailing args optional."""
    out = bytearray()
    n = 0
    for ev in events:
        ev = tuple(ev) + (0.0,) * (6 - len(ev))
        ch, blk, op, a0, a1, a2 = ev
        op = OPS[op] if isinstance(op, str) else int(op)
        out += struct.pack("<IIIfff", int(ch) & 0xFFFFFFFF, int(blk), op, float(a0), float(a1), float(a2))
        n += 1
    return bytes(out), n


def write_request(path, I, Q, events):
    I = np.ascontiguousarray(I, dtype=np.int16)
    Q = np.ascontiguousarray(Q, dtype=np.int16)
    assert I.shape == Q.shape and I.ndim == 2 and I.shape[1] % N_BLOCK == 0
    ev, n_ev = pack_events(events)
    with open(path, "wb") as f:
        f.write(b"REFSDR01" + struct.pack("<IIII", I.shape[0], I.shape[1] // N_BLOCK, n_ev, 0))
        f.write(ev)
        f.write(I.tobytes())
        f.write(Q.tobytes())


def run(I, Q, events, jobs=None):
    """Returns dict(audio f32 [C,S], pcm i16 [C,S], status f32 [C,16])."""
    if not available():
        raise RuntimeError("oracle/_ref/refsdr is not built (run `make -C oracle ref` where /root/reference exists)")
    jobs = jobs or os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as d:
        req, resp = os.path.join(d, "req.bin"), os.path.join(d, "resp.bin")
        write_request(req, I, Q, events)
        subprocess.run([REFSDR, "run", req, resp, str(jobs)], check=True)
        raw = open(resp, "rb").read()
    assert raw[:8] == b"REFOUT01"
    nch, nblk, nst, _ = struct.unpack("<IIII", raw[8:24])
    ns = nblk * N_BLOCK
    off = 24
    audio = np.frombuffer(raw, np.float32, nch * ns, off).reshape(nch, ns).copy(); off += nch * ns * 4
    pcm = np.frombuffer(raw, np.int16, nch * ns, off).reshape(nch, ns).copy(); off += nch * ns * 2
    status = np.frombuffer(raw, np.float32, nch * nst, off).reshape(nch, nst).copy()
    return dict(audio=audio, pcm=pcm, status=status)


def bench(I, Q, events, seconds, jobs):
    """Times update() only, one worker process per job; returns the driver's JSON dict."""
    import json
    with tempfile.TemporaryDirectory() as d:
        req = os.path.join(d, "req.bin")
        write_request(req, I, Q, events)
        out = subprocess.run([REFSDR, "bench", req, str(seconds), str(jobs)], check=True,
                             capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])
