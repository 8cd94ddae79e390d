"""csrc/ring_book.h (the page-locked rings of b200jpg_decode_files) under a multi-threaded stress: no region is
overwritten while in flight, no deadlock -- in particular an empty ring accepts a request that needs the wrap-around
(the case that once hung the whole-file engine).  CPU only; builds tests/cpp/ring_stress.cpp with g++."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [("4", "20000", "4096"), ("1", "50000", "1024"), ("8", "5000", "100000")])
def test_ring_book_stress(tmp_path, args):
    exe = str(tmp_path / "ring_stress")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "ring_stress.cpp")])
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("ok ")
