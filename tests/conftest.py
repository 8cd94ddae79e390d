import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def reftest_files(include_disabled=False):
    """The reference's reftest corpus (tests/reftest/mod.rs:9-16) minus lossless, as (jpg, png) pairs."""
    root = os.path.join(GOLDEN, "reftest")
    disabled = [l.strip() for l in open(os.path.join(root, "disabled.list")) if l.strip() and not l.startswith("#")]
    out = []
    for p in sorted(glob.glob(root + "/*.jpg") + glob.glob(root + "/mozilla/*.jpg")):
        rel = os.path.relpath(p, root)
        if rel in disabled and not include_disabled:
            continue
        out.append(p)
    return out


def bench_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "benches", "*.jpg")))


def load_png_like_reftest(png_path, pixel_format):
    """Golden PNG -> flat uint8 array the way tests/reftest/mod.rs:41-90 reads it (16-bit PNGs are
    stripped to 8 bits by the png crate's default transformations)."""
    from PIL import Image
    im = Image.open(png_path)
    if im.mode in ("I;16", "I;16B", "I"):
        a = (np.asarray(im).astype(np.uint32) >> 8).astype(np.uint8)
        return a.reshape(-1)
    if pixel_format == 0:
        return np.asarray(im.convert("L")).reshape(-1)
    return np.asarray(im.convert("RGB")).reshape(-1)


def cmyk_to_rgb(d):
    """tests/reftest/mod.rs:137-163 (f32 arithmetic)"""
    d = d.reshape(-1, 4).astype(np.float32) / np.float32(255.0)
    k = d[:, 3:4]
    cmy = d[:, :3] * (np.float32(1.0) - k) + k
    return ((np.float32(1.0) - cmy) * np.float32(255.0)).astype(np.uint8).reshape(-1)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def J():
    import jpeg_decoder_b200
    jpeg_decoder_b200.lib()
    return jpeg_decoder_b200
