"""Host-only: the PLY passthrough of the PointCloud3f shim (tests/cpp/test_ply.cpp) — round trips in ascii and
binary, a hand-written file with another property layout and a face element, append / clear, a missing file."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ply_passthrough(tmp_path):
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    exe = str(tmp_path / "test_ply")
    lib = os.path.join(ROOT, "cilantro_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_ply.cpp"), "-o", exe, "-L", lib, "-lcilantro_b200",
                           f"-Wl,-rpath,{lib}"], env=env)
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PLY checks passed" in out.stdout
