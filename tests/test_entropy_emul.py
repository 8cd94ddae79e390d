"""Device entropy decoding (csrc/entropy_dev.h), emulated on the CPU pass by pass with the functions the kernels
call (tests/cpp/entropy_emul.cpp): cold start, synchronisation passes, block prefix, write pass, DC prefix.

The property under test is the one the product relies on: whenever the device path ACCEPTS an image, its dense
coefficients equal what the host decoder -- the restatement of the reference's sequential Huffman loop,
src/huffman.rs + src/decoder.rs:1086-1172 -- produces, bit for bit; everything else is flagged and goes to the host.
Checked on every reftest / bench fixture, on synthetic BASELINE-shaped files and on ~1500 corrupted scans.
CPU only (the GPU kernels themselves are covered by tests/test_gpu_entropy.py)."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "jpeg_decoder_b200")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    from jpeg_decoder_b200 import build
    build.build()
    exe = str(tmp_path_factory.mktemp("emul") / "entropy_emul")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "entropy_emul.cpp"),
                           os.path.join(PKG, "csrc", "host_decoder.cpp"), "-L" + PKG, "-lb200jpg", "-Wl,-rpath," + PKG])
    return exe


def run(exe, args):
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    last = out.stdout.strip().splitlines()[-1]
    assert last.startswith("ok "), last
    return out.stdout


def test_fixtures_decode_identically(emul):
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "reftest", "**", "*.jpg"), recursive=True))
    files += sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "benches", "*.jpg")))
    assert len(files) > 30
    text = run(emul, files)
    # the baseline fixtures really take the device route (not merely "nothing mismatched because nothing ran")
    for name in ("tower.jpg", "rgb.jpg", "jpg-size-33x33.jpg", "jpg-cmyk-1.jpg", "grayscale_large.jpg", "16bit-qtables.jpg"):
        line = [l for l in text.splitlines() if l.split(":")[0].endswith(name)]
        assert line and "device == host" in line[0], (name, line)
    # progressive / restart-interval / non-interleaved files stay with the host decoder
    for name in ("tower_progressive.jpg", "restarts.jpg", "non-interleaved-mcu.jpg", "mjpeg.jpg"):
        line = [l for l in text.splitlines() if l.split(":")[0].endswith(name)]
        assert line and "host path" in line[0], (name, line)


def test_crashtest_files_never_mismatch(emul):
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "crashtest", "**", "*.jpg"), recursive=True))
    assert len(files) > 50
    run(emul, files)


@pytest.mark.parametrize("shape", [(1920, 1080, 2), (640, 480, 0), (333, 217, 1), (48, 1000, 2)])
def test_synthetic_baseline_shapes(emul, tmp_path, shape):
    from jpeg_decoder_b200 import workload
    w, h, ss = shape
    p = tmp_path / "s.jpg"
    p.write_bytes(workload.synth_jpeg(w, h, seed=1234, subsampling=ss))
    assert "device == host" in run(emul, [str(p)])


def test_corrupted_scans_are_flagged_or_identical(emul, tmp_path):
    from jpeg_decoder_b200 import workload
    p = tmp_path / "s.jpg"
    p.write_bytes(workload.synth_jpeg(333, 217, seed=5, subsampling=2))
    g = os.path.join(ROOT, "tests", "golden")
    text = run(emul, ["--corrupt", "300", "7", str(p), os.path.join(g, "benches", "tower.jpg"),
                      os.path.join(g, "reftest", "mozilla", "jpg-size-33x33.jpg"), os.path.join(g, "reftest", "mozilla", "jpg-cmyk-1.jpg"),
                      os.path.join(g, "reftest", "grayscale_large.jpg")])
    last = text.strip().splitlines()[-1].split()
    decoded, flagged = int(last[1]), int(last[7])
    assert decoded > 300 and flagged > 100, last  # both outcomes occur


def _dri_jpeg(w, h, subsampling, interval, seed=0):
    """Baseline JPEG with restart markers every `interval` MCUs (PIL's restart_marker_blocks)."""
    import io
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 90 * np.sin(x / 41 + c) + 40 * np.cos(y / 57 + c) for c in range(3)], -1) + rng.normal(0, 6, (h, w, 3))
    buf = io.BytesIO()
    Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(buf, "JPEG", quality=90, subsampling=subsampling, restart_marker_blocks=interval)
    return buf.getvalue()


@pytest.mark.parametrize("shape", [(640, 480, 0, 80), (1920, 1080, 2, 120), (200, 200, 2, 7), (333, 217, 1, 21)])
def test_restart_intervals(emul, tmp_path, shape):
    """Every restart interval is decoded as a scan of its own (src/decoder.rs:910-931)."""
    w, h, ss, ri = shape
    p = tmp_path / "d.jpg"
    p.write_bytes(_dri_jpeg(w, h, ss, ri))
    out = run(emul, [str(p)])
    assert "device == host" in out and "in 1 interval(s)" not in out, out


def test_corrupted_restart_interval_scans(emul, tmp_path):
    """Damage around restart markers: the reference only finds an RSTn that directly follows the interval's last MCU
    (anything else is an error there), so the device must not accept an interval with whole bytes to spare."""
    files = []
    for k, shape in enumerate([(640, 480, 0, 80), (200, 200, 2, 7)]):
        p = tmp_path / ("d%d.jpg" % k)
        p.write_bytes(_dri_jpeg(*shape, seed=k))
        files.append(str(p))
    text = run(emul, ["--descending", "--corrupt", "300", "3"] + files)
    last = text.strip().splitlines()[-1].split()
    assert int(last[1]) > 200 and int(last[7]) > 50, last


def test_eligibility_edges(emul, tmp_path):
    """What decides host vs device before any kernel runs (csrc/entropy_host.h): anything but stuffed bytes and the
    expected RSTn up to an EOI keeps the scan on the host; harmless oddities (bytes after EOI, spare bytes after the last
    MCU) do not.  Wherever the device route is taken the emulator also checks it against the host decoder."""
    from jpeg_decoder_b200 import workload
    base = workload.synth_jpeg(160, 120, seed=21, subsampling=2)
    assert base.endswith(b"\xff\xd9")
    sos = base.rfind(b"\xff\xda")
    mid = sos + 14 + (len(base) - sos - 16) // 2
    while base[mid - 1] == 0xFF or base[mid] == 0xFF:   # do not cut a stuffed pair
        mid += 1
    dri = _dri_jpeg(200, 200, 2, 7)
    rst = dri.find(b"\xff\xd1")
    cases = {
        "plain": (base, "device == host"),
        "bytes_after_eoi": (base + b"\x00\x01\x02garbage", "device == host"),
        "spare_bytes_before_eoi": (base[:-2] + b"\x00\x00\x00" + base[-2:], "device == host"),
        "rst_without_dri": (base[:mid] + b"\xff\xd0" + base[mid:], "host path"),
        "fill_bytes_before_eoi": (base[:-2] + b"\xff\xff\xd9", "host path"),
        "no_eoi": (base[:-2], "host path"),
        "other_marker_in_scan": (base[:mid] + b"\xff\xe0\x00\x02" + base[mid:], "host path"),
        "second_scan_follows": (base[:-2] + base[sos:], "host path"),
        "dri_plain": (dri, "device == host"),
        "dri_wrong_rst_number": (dri[:rst] + b"\xff\xd3" + dri[rst + 2:], "host path"),
        "dri_missing_rst": (dri[:rst] + dri[rst + 2:], "host path"),
    }
    for name, (data, want) in cases.items():
        p = tmp_path / (name + ".jpg")
        p.write_bytes(data)
        out = run(emul, [str(p)])
        line = out.strip().splitlines()[0]
        assert want in line, (name, line)


@pytest.mark.parametrize("passes", [0, 1, 2])
def test_exhausted_sync_budget_is_flagged_not_misdecoded(emul, tmp_path, passes):
    """When the synchronisation passes run out of budget with a change still pending, a successor's state and value
    count date from its predecessor's OLD state.  The write pass re-decodes from the new one: it usually ends in the same
    state (codes self-synchronise) with a DIFFERENT value count -- which would shift every later value offset.  The
    chain check therefore compares the count as well (ke_entropy.cu); the emulator mirrors it.  Property: with any
    budget, accepted => identical to the host decoder (the emulator's own check), and at least the big scan is not
    accepted with a budget this small."""
    from jpeg_decoder_b200 import workload
    files = []
    for k, (w, h, ss) in enumerate([(1920, 1080, 2), (640, 480, 0), (333, 217, 1)]):
        p = tmp_path / ("b%d.jpg" % k)
        p.write_bytes(workload.synth_jpeg(w, h, seed=40 + k, subsampling=ss))
        files.append(str(p))
    for order in ([], ["--descending"]):
        text = run(emul, order + ["--passes", str(passes)] + files)   # rc 0 = nothing accepted that differs from the host
        first = [l for l in text.splitlines() if l.split(":")[0].endswith("b0.jpg")][0]
        if passes == 0:
            assert "flagged" in first, first
