"""CPU tier: the generated constant tables (filter designs, Hilbert half, sine LUT)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_inc(path):
    src = open(path).read()
    tabs = {}
    for m in re.finditer(r"SDR_TAB_(\w+)\[(\d+)\] = \{([^}]*)\}", src):
        vals = np.array([int(v.strip().rstrip("u"), 16) for v in m.group(3).split(",") if v.strip()], np.uint32)
        assert vals.size == int(m.group(2))
        tabs[m.group(1)] = vals.view(np.float32)
    return tabs


def test_product_and_oracle_tables_identical():
    a = open(os.path.join(ROOT, "oracle", "sdr_oracle_tables.inc")).read()
    b = open(os.path.join(ROOT, "audiosdr_b200", "csrc", "sdr_tables.inc")).read()
    assert a == b


def test_table_shapes_and_known_values():
    t = parse_inc(os.path.join(ROOT, "audiosdr_b200", "csrc", "sdr_tables.inc"))
    assert len(t) == 17 and t["HILBERT"].size == 64 and t["SINE"].size == 257
    k = np.arange(257)
    assert np.array_equal(t["SINE"], np.round(np.sin(2 * np.pi * k / 256), 8).astype(np.float32))
    assert t["HILBERT"][63] == np.float32(-0.6365587) and t["HILBERT"][0] == np.float32(-0.003780058)
    # the reference's data quirks are load-bearing
    assert t["AUDIO_WSPR"][8] < 0          # sign typo in wspr_coefs row 2
    assert t["AUDIO_CW"][17] == np.float32(-1.963497179540541810)  # permuted last row of bw470_coefs
    assert t["AUDIO_AM"][3] == np.float32(1.907327579454288770)   # the second (active) bw3900 definition
    for name, v in t.items():           # every biquad section is stable: |a2| < 1
        if v.size == 20 and name != "AUDIO_CW":
            assert np.all(np.abs(v.reshape(4, 5)[:, 4]) < 1.0), name


@pytest.mark.skipif(not os.path.exists("/root/reference/SRC/AudioSDRlib/AudioSDR.h"), reason="reference tree not present")
def test_tables_match_reference_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
