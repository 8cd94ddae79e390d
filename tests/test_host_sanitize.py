"""AddressSanitizer + UndefinedBehaviorSanitizer over the host half of the decoder (csrc/host_decoder.cpp: marker parser,
Huffman decoder, scan driver, sparse-stream writer, device-scan payload builder) on every fixture, the reference's 111
crash-test files included -- the analogue of tests/crashtest/mod.rs ("decoding must not panic"), with the sanitizers
standing in for Rust's bounds checks.  CPU only."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "jpeg_decoder_b200")


def test_host_decoder_is_clean_under_asan_ubsan(tmp_path):
    from jpeg_decoder_b200 import build
    build.build()   # b200jpg_update_component_sizes / b200jpg_choose_idct_size come from the product library
    exe = str(tmp_path / "hd_san")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-fno-omit-frame-pointer", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_decoder_sanitize.cpp"),
                           os.path.join(PKG, "csrc", "host_decoder.cpp"), "-L" + PKG, "-lb200jpg", "-Wl,-rpath," + PKG])
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "**", "*.jpg"), recursive=True))
    files += sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "**", "*.jpeg"), recursive=True))
    assert len(files) > 150
    out = subprocess.run([exe, *files], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-6000:])
    last = out.stdout.strip().splitlines()[-1]
    assert last.startswith("ok ") and "ERROR" not in out.stderr, (last, out.stderr[-3000:])
    decoded = int(last.split()[1])
    assert decoded >= 50   # the well-formed fixtures decode; the crash-test files are mostly rejected, cleanly
