"""ent_unstuff (csrc/entropy_host.h) -- the host loop that removes byte stuffing from a scan before it is uploaded for
device entropy decoding -- exists in three forms (scalar, AVX2, AVX-512 VBMI2, chosen at run time).  All of them against a
byte-at-a-time restatement on 20000 random buffers (tests/cpp/unstuff_test.cpp).  CPU only; forms the CPU lacks fall
back to the next one, so the test passes on any x86-64."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unstuff_variants_agree(tmp_path):
    exe = str(tmp_path / "unstuff_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "unstuff_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok "), out.stdout[-2000:] + out.stderr[-2000:]
