"""oracle/oracle_lib.py -- TEST INFRASTRUCTURE (not product code).

ctypes binding of oracle/libsdr_oracle.so, the plain-C restatement of the reference chain
(oracle/sdr_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsdr_oracle.so")
N_BLOCK = 128
N_STATUS = 16
N_TAPS = 12
TAPS = ["in_i", "in_q", "nb_i", "nb_q", "if_i", "if_q", "dm_i", "dm_q", "demod", "audf", "agc", "als"]
_lib = None


class Event(C.Structure):
    _fields_ = [("channel", C.c_uint32), ("block", C.c_uint32), ("opcode", C.c_uint32),
                ("a0", C.c_float), ("a1", C.c_float), ("a2", C.c_float)]


def build():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        L.ora_new.restype = C.c_void_p
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_apply.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float]
        L.ora_set_taps.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_update_i16.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ora_update_f32.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ora_status.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_get_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_run.restype = C.c_double
        L.ora_run.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _events(events):
    from .ref_client import OPS
    evs = []
    for ev in events:
        ev = tuple(ev) + (0.0,) * (6 - len(ev))
        ch, blk, op, a0, a1, a2 = ev
        op = OPS[op] if isinstance(op, str) else int(op)
        evs.append(Event(int(ch) & 0xFFFFFFFF, int(blk), op, float(a0), float(a1), float(a2)))
    arr = (Event * max(len(evs), 1))(*evs)
    return arr, len(evs)


def run(I, Q, events, threads=1, want_pcm=True):
    """I/Q: [C, S] int16 (wire format) or float32.  Returns dict(audio, pcm, status, seconds)."""
    L = lib()
    I = np.ascontiguousarray(I)
    Q = np.ascontiguousarray(Q)
    assert I.shape == Q.shape and I.dtype == Q.dtype and I.ndim == 2 and I.shape[1] % N_BLOCK == 0
    fmt = {np.dtype(np.int16): 0, np.dtype(np.float32): 1}[I.dtype]
    nch, ns = I.shape
    audio = np.empty((nch, ns), np.float32)
    pcm = np.empty((nch, ns), np.int16) if want_pcm else None
    status = np.zeros((nch, N_STATUS), np.float32)
    ev, nev = _events(events)
    secs = L.ora_run(nch, ns // N_BLOCK, C.cast(ev, C.c_void_p), nev, fmt, I.ctypes.data, Q.ctypes.data,
                     audio.ctypes.data, pcm.ctypes.data if want_pcm else None, status.ctypes.data, int(threads))
    return dict(audio=audio, pcm=pcm, status=status, seconds=secs)


class Channel:
    """Single-channel stepping interface with stage taps (for stage-level known-answer tests)."""

    def __init__(self):
        self.L = lib()
        self.p = self.L.ora_new()
        self.taps = np.zeros((N_TAPS, N_BLOCK), np.float32)
        self.L.ora_set_taps(self.p, self.taps.ctypes.data)

    def apply(self, op, a0=0.0, a1=0.0, a2=0.0):
        from .ref_client import OPS
        op = OPS[op] if isinstance(op, str) else int(op)
        assert self.L.ora_apply(self.p, op, a0, a1, a2) == 0

    def update(self, I, Q):
        I = np.ascontiguousarray(I); Q = np.ascontiguousarray(Q)
        audio = np.empty(N_BLOCK, np.float32); pcm = np.empty(N_BLOCK, np.int16)
        fn = self.L.ora_update_i16 if I.dtype == np.int16 else self.L.ora_update_f32
        fn(self.p, I.ctypes.data, Q.ctypes.data, audio.ctypes.data, pcm.ctypes.data)
        return audio, pcm

    def status(self):
        s = np.zeros(N_STATUS, np.float32)
        self.L.ora_status(self.p, s.ctypes.data)
        return s

    def phases(self):
        a = C.c_float(); b = C.c_float()
        self.L.ora_get_phases(self.p, C.byref(a), C.byref(b))
        return a.value, b.value

    def __del__(self):
        try:
            self.L.ora_free(self.p)
        except Exception:
            pass
