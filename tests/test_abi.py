"""CPU tier: the C-ABI library loads and exports every symbol include/sdr_batch.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sdr_batch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdr_batch_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    f = declared_functions()
    for name in ("sdr_batch_create", "sdr_batch_configure", "sdr_batch_process_device", "sdr_batch_process_host",
                 "sdr_batch_set", "sdr_batch_get_status", "sdr_batch_destroy"):
        assert name in f


def test_cuda_library_exports_every_declared_symbol():
    from audiosdr_b200 import api, build
    path = build.build_library()  # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert set(api.EXPORTS) == set(declared_functions())
    assert b"sm_100a" in ctypes.cast(lib.sdr_batch_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_setter_ids_agree_between_header_python_and_oracle():
    from audiosdr_b200 import api
    from oracle import ref_client as rc
    src = open(os.path.join(ROOT, "include", "sdr_batch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = src[:src.index("} sdr_setter;")]
    body = body[body.rindex("typedef enum {") + len("typedef enum {"):]
    val, ids = 0, {}
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            name, v = [s.strip() for s in item.split("=")]
            val = int(v)
        else:
            name = item
        ids[name.replace("SDR_SET_", "")] = val
        val += 1
    assert ids == api.SETTERS
    assert {k: v for k, v in rc.OPS.items() if k != "oracle_identity_IF"} == api.SETTERS


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or execute anything under oracle/ (or the emulation scaffold)."""
    pkg = os.path.join(ROOT, "audiosdr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in src.replace("/* test", "") or f == "sdr_tables.inc", f
                assert "libsdr_emu" not in src and "libsdr_oracle" not in src, f


def test_no_device_no_fallback(monkeypatch):
    """Without the CUDA library the product raises; it never computes on the CPU."""
    from audiosdr_b200 import api
    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(api, "lib_path", lambda: "/nonexistent/libsdr_batch.so")
    with pytest.raises(api.SdrError):
        api.SdrBatch(4)
