"""CPU tier: the FP64-free / divide-free forms used by the kernels are EXACTLY the reference's double expressions.

tests/emu/exhaustive_lut.cpp walks every float (or int16) of each operand range and compares bit for bit:
sine-table index (H:364), Q15 input scaling (C:68), AGC table index (C:419), int16 output truncation (C:160),
cosine argument Phase + PI/2 (H:376), arctangent +-PI (H:396-397) and the PLL wrap compares (C:735-736)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exhaustive_exact_forms(tmp_path):
    exe = str(tmp_path / "exhaustive_lut")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=gnu++17", "-pthread",
                    os.path.join(ROOT, "tests", "emu", "exhaustive_lut.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout
