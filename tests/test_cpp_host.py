"""CPU tier: the C++ host class (include/SdrBatch.hpp) compiles, links against the product library and, without a
GPU, fails loudly instead of falling back to anything."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_class_builds_and_links(tmp_path):
    from audiosdr_b200 import build
    lib = build.build_library()
    exe = str(tmp_path / "host_class_smoke")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "host_class_smoke.cpp"), "-o", exe,
                    "-L" + os.path.dirname(lib), "-lsdr_batch", "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sm_100a" in r.stdout and ("NO_DEVICE" in r.stdout or "GPU_OK" in r.stdout)


def test_cpp_aux_classes_build_and_link(tmp_path):
    """include/SdrAux.hpp (PreProcessorBatch, IQGeneratorBatch) against libsdr_aux.so; same no-fallback rule."""
    from audiosdr_b200 import build
    lib = build.build_aux_library()
    exe = str(tmp_path / "aux_class_smoke")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "aux_class_smoke.cpp"), "-o", exe,
                    "-L" + os.path.dirname(lib), "-lsdr_aux", "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sm_100a" in r.stdout and ("NO_DEVICE" in r.stdout or "GPU_OK 1" in r.stdout)
