"""oracle/aux_lib.py -- TEST INFRASTRUCTURE.  ctypes access to the C restatement of the pre-processor and the I/Q
generator (oracle/sdr_aux_oracle.c) and a client of the host-compiled unmodified reference (oracle/_ref/refaux).

kind: "pp" (AudioSDRpreProcessor, inputs I and Q) or "iq" (AudioIQgenerator, one real input).
events: list of (channel | None, block, opcode_name, arg)."""
import ctypes as C
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
N_BLOCK = 128
KIND = {"pp": 1, "iq": 2, "grab": 3}
OPS = {"pp": {"startAutoI2SerrorDetection": 1, "stopAutoI2SerrorDetection": 2, "setI2SerrorCompensation": 3, "swapIQ": 4},
       "iq": {"setGainBalance": 1}, "grab": {}}
PP_STATUS = ("auto_detect", "correction", "failure_count", "success_count", "saved_sample", "swap")
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libsdr_aux_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
        _lib = C.CDLL(so)
        _lib.ora_aux_run.restype = C.c_int
        _lib.ora_aux_run.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
        _lib.ora_aux_power128.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ora_grab_run.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ora_grab_run.restype = None
        _lib.ora_grab_spectrum.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.ora_grab_spectrum.restype = None
    return _lib


def _pack_events(kind, events):
    ev = sorted(events, key=lambda e: e[1])
    out = b""
    for e in ev:
        ch = 0xFFFFFFFF if e[0] is None else int(e[0])
        out += struct.pack("<IIIf", ch, int(e[1]), OPS[kind][e[2]], float(e[3]) if len(e) > 3 else 0.0)
    return out, len(ev)


def run(kind, planes, events=(), threads=4):
    """planes: (I, Q) int16 [C, S] for "pp"; (X,) for "iq".  Returns (out0, out1, status[C, 8] int32)."""
    L = lib()
    a = np.ascontiguousarray(planes[0], np.int16)
    b = np.ascontiguousarray(planes[1], np.int16) if kind == "pp" else None
    nch, ns = a.shape
    assert ns % N_BLOCK == 0
    o0, o1 = np.empty_like(a), np.empty_like(a)
    st = np.zeros((nch, 8), np.int32)
    blob, nev = _pack_events(kind, events)
    buf = C.create_string_buffer(blob, len(blob)) if nev else None
    L.ora_aux_run(KIND[kind], nch, ns // N_BLOCK, C.cast(buf, C.c_void_p) if nev else None, nev, a.ctypes.data,
                  b.ctypes.data if b is not None else None, o0.ctypes.data, o1.ctypes.data, st.ctypes.data, threads)
    return o0, o1, st


def grab_run(I, Q):
    """AudioGrabberComplex256: n_blocks updates from a fresh object, then grab().  Returns (out int32 [C, 512] with -1 where
    grab() delivers nothing, flags int32 [C, 2] = {newDataAvailable before the grab, buffer valid})."""
    I = np.ascontiguousarray(I, np.int16); Q = np.ascontiguousarray(Q, np.int16)
    nch, ns = I.shape
    out = np.empty((nch, 512), np.int32); flags = np.empty((nch, 2), np.int32)
    lib().ora_grab_run(nch, ns // N_BLOCK, I.ctypes.data, Q.ctypes.data, out.ctypes.data, flags.ctypes.data)
    return out, flags


def grab_spectrum(snap):
    """Power spectra float32 [C, 256] of snapshots int16 [C, 512] (interleaved re, im): the oracle of sdr_grabber_spectrum."""
    snap = np.ascontiguousarray(snap, np.int16)
    out = np.empty((snap.shape[0], 256), np.float32)
    lib().ora_grab_spectrum(snap.shape[0], snap.ctypes.data, out.ctypes.data)
    return out


def ref_grab_run(I, Q, jobs=8):
    """The same from the unmodified reference (oracle/_ref/refaux kind 3)."""
    _, _, st = ref_run("grab", (I, Q), (), jobs=jobs)
    return st[:, 8:520].copy(), st[:, :2].copy()


def power128(I, Q):
    p = np.empty(128, np.float32)
    I = np.ascontiguousarray(I, np.int16); Q = np.ascontiguousarray(Q, np.int16)
    lib().ora_aux_power128(I.ctypes.data, Q.ctypes.data, p.ctypes.data)
    return p


# ---- the unmodified reference, host-compiled (oracle/_ref/refaux)
REFAUX = os.path.join(HERE, "_ref", "refaux")


def ref_available():
    return os.path.exists(REFAUX)


def _write_request(path, kind, planes, events):
    a = np.ascontiguousarray(planes[0], np.int16)
    nch, ns = a.shape
    blob, nev = _pack_events(kind, events)
    with open(path, "wb") as f:
        f.write(b"REFAUX01" + struct.pack("<IIII", KIND[kind], nch, ns // N_BLOCK, nev))
        f.write(blob)
        f.write(a.tobytes())
        if kind in ("pp", "grab"):
            f.write(np.ascontiguousarray(planes[1], np.int16).tobytes())
    return nch, ns


def ref_run(kind, planes, events=(), jobs=8):
    with tempfile.TemporaryDirectory() as d:
        req, resp = os.path.join(d, "req"), os.path.join(d, "resp")
        nch, ns = _write_request(req, kind, planes, events)
        subprocess.check_call([REFAUX, "run", req, resp, str(jobs)])
        raw = open(resp, "rb").read()
    assert raw[:8] == b"REFAUO01"
    k, c, nb, nst = struct.unpack("<IIII", raw[8:24])
    assert (k, c, nb) == (KIND[kind], nch, ns // N_BLOCK)
    off = 24
    o0 = np.frombuffer(raw, np.int16, nch * ns, off).reshape(nch, ns); off += nch * ns * 2
    o1 = np.frombuffer(raw, np.int16, nch * ns, off).reshape(nch, ns); off += nch * ns * 2
    st = np.frombuffer(raw, np.int32, nch * nst, off).reshape(nch, nst)
    return o0.copy(), o1.copy(), st.copy()


def ref_bench(kind, planes, events=(), seconds=5.0, jobs=1):
    with tempfile.TemporaryDirectory() as d:
        req = os.path.join(d, "req")
        _write_request(req, kind, planes, events)
        out = subprocess.check_output([REFAUX, "bench", req, str(seconds), str(jobs)], text=True)
    return json.loads(out.strip().splitlines()[-1])
