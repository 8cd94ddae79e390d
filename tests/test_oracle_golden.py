"""Pins the oracle's whole path (feeder + worker + compute_image) on the reference's golden vectors:
every DCT-based JPEG of tests/reftest/images against its PNG within +-3 -- the reference's own
bound (tests/reftest/mod.rs:99) -- and the scaled decodes of rgb.jpg (tests/reftest/mod.rs:18-25).
Also: the product's C++ host decoder produces exactly the coefficients the oracle's feeder does,
and neither crashes on the reference's crashtest corpus (tests/crashtest/mod.rs).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, bench_files, cmyk_to_rgb, load_png_like_reftest, reftest_files


@pytest.mark.parametrize("arith", [0, 1])
def test_reftest_goldens(oracle_mod, arith):
    files = reftest_files()
    assert len(files) >= 33
    worst = 0
    for p in files:
        d = oracle_mod.Decoder(open(p, "rb").read(), arith)
        px = d.decode()
        info = d.info()
        if info.pixel_format == 3:
            px = cmyk_to_rgb(px)
        ref = load_png_like_reftest(p[:-4] + ".png", info.pixel_format)
        assert ref.size == px.size, p
        diff = int(np.abs(ref.astype(int) - px.astype(int)).max())
        worst = max(worst, diff)
        assert diff <= 3, (p, diff)
    assert worst <= 3


@pytest.mark.parametrize("req,png", [((500, 333), "rgb.png"), ((250, 167), "rgb_250x167.png"), ((125, 84), "rgb_125x84.png"),
                                     ((63, 42), "rgb_63x42.png")])
def test_reftest_scaled(oracle_mod, req, png):
    root = os.path.join(GOLDEN, "reftest")
    d = oracle_mod.Decoder(open(os.path.join(root, "rgb.jpg"), "rb").read())
    d.read_info()
    w, h = d.scale(*req)
    px = d.decode()
    ref = load_png_like_reftest(os.path.join(root, png), 2)
    assert (w, h) == req and ref.size == px.size
    assert np.abs(ref.astype(int) - px.astype(int)).max() <= 3


def test_read_info_then_decode(oracle_mod, J):
    """tests/lib.rs:34-50"""
    data = open(os.path.join(GOLDEN, "reftest", "mozilla", "jpg-progressive.jpg"), "rb").read()
    d = oracle_mod.Decoder(data)
    ref = d.decode()
    ref_info = d.info()
    d2 = oracle_mod.Decoder(data)
    d2.read_info()
    info = d2.info()
    assert (info.width, info.height, info.pixel_format, info.coding_process) == (ref_info.width, ref_info.height, ref_info.pixel_format, ref_info.coding_process)
    assert np.array_equal(d2.decode(), ref)
    # product host half: read_info, then entropy decode resumes on the same reader
    pd = J.Decoder(data)
    assert pd.info() is None
    pd.read_info()
    pi = pd.info()
    assert (pi.width, pi.height, pi.pixel_format, pi.coding_process) == (32, 32, J.PF_RGB24, J.CP_DCT_PROGRESSIVE)
    desc = pd.entropy_decode()
    assert desc.ncomp == 3


def test_host_decoder_matches_oracle_feeder(oracle_mod, J):
    files = reftest_files(include_disabled=True) + bench_files()
    checked = 0
    for p in files:
        data = open(p, "rb").read()
        o = oracle_mod.Decoder(data)
        try:
            o.decode()
            oerr = 0
        except oracle_mod.OracleError as e:
            oerr = e.code
        pd = J.Decoder(data)
        try:
            desc = pd.entropy_decode()
            perr = 0
        except J.B200JpgError as e:
            perr = -e.code
        assert oerr == perr, p
        if oerr:
            continue
        comps, qts = o.components()
        assert desc.ncomp == len(comps)
        assert desc.color_transform == o.color_transform()
        for i in range(desc.ncomp):
            for f in ("identifier", "h", "v", "tq", "dct_scale", "size_w", "size_h", "block_w", "block_h"):
                assert getattr(desc.comps[i], f) == getattr(comps[i], f)
            assert np.array_equal(pd.coefficients(desc, i), o.coefficients(i)), (p, i)
            assert np.array_equal(pd.qtable(desc, i), qts[i])
        checked += 1
    assert checked >= 38


def test_corrupted_progressive_scans_agree_with_the_oracle(oracle_mod, J):
    """The refinement scans of the product's host decoder walk a per-block non-zero map (host_decoder.cpp:
    refine_non_zeroes_map) where the reference -- and the oracle -- walk the coefficients (src/decoder.rs:1250-1298).
    On damaged progressive files (bit flips, scrambled and dropped runs inside the scans: repeated bands, runs that overshoot,
    corrections of coefficients that shifted to zero) both must still end the same way: the same error class, or the same
    coefficients bit for bit."""
    rng = np.random.default_rng(20261017)
    srcs = [os.path.join(GOLDEN, "benches", "tower_progressive.jpg")]
    srcs += [p for p in reftest_files() if "progressive" in os.path.basename(p) and os.path.getsize(p) > 4096][:3]
    agree_ok = agree_err = 0
    for p in srcs:
        data = bytearray(open(p, "rb").read())
        first_sos = data.find(b"\xff\xda")
        assert first_sos > 0
        for t in range(60):
            c = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                at = int(rng.integers(first_sos + 14, len(c) - 4))
                kind = int(rng.integers(0, 3))
                if kind == 0:
                    c[at] ^= 1 << int(rng.integers(0, 8))
                elif kind == 1:
                    n = int(rng.integers(1, 24))
                    c[at:at + n] = bytes(int(x) for x in rng.integers(0, 255, size=n))   # no 0xFF: stays inside the scan
                else:
                    del c[at:at + int(rng.integers(1, 40))]
            c = bytes(c)
            o = oracle_mod.Decoder(c)
            try:
                o.decode()
                oerr = 0
            except oracle_mod.OracleError as e:
                oerr = e.code
            pd = J.Decoder(c)
            try:
                desc = pd.entropy_decode()
                perr = 0
            except J.B200JpgError as e:
                perr = -e.code
            assert oerr == perr, (p, t, oerr, perr)
            if oerr:
                agree_err += 1
                continue
            for i in range(desc.ncomp):
                assert np.array_equal(pd.coefficients(desc, i), o.coefficients(i)), (p, t, i)
            agree_ok += 1
    assert agree_ok >= 20 and agree_ok + agree_err == 60 * len(srcs), (agree_ok, agree_err)


def test_crashtest_corpus(oracle_mod, J):
    """tests/crashtest/mod.rs:8-17: malformed files must produce errors, never crash; both decoders agree on the class."""
    files = sorted(f for f in glob.glob(os.path.join(GOLDEN, "crashtest", "**", "*"), recursive=True) if os.path.isfile(f))
    assert len(files) == 111
    for p in files:
        data = open(p, "rb").read()
        try:
            oracle_mod.Decoder(data).decode()
            oe = 0
        except oracle_mod.OracleError as e:
            oe = e.code
        try:
            J.Decoder(data).entropy_decode()
            pe = 0
        except J.B200JpgError as e:
            pe = -e.code
        assert pe == oe or (pe == 0 and oe != 0), (p, oe, pe)


def test_metadata(oracle_mod, J):
    """tests/lib.rs:52-170: ICC assembly rules, Exif, XMP -- oracle and product host decoder."""
    icc = os.path.join(GOLDEN, "icc")
    for mk in (lambda b: oracle_mod.Decoder(b), lambda b: J.Decoder(b)):
        def run(path):
            d = mk(open(path, "rb").read())
            if hasattr(d, "entropy_decode"):
                d.entropy_decode()
            else:
                d.decode()
            return d
        d = run(os.path.join(GOLDEN, "reftest", "mozilla", "jpg-srgb-icc.jpg"))
        assert d.icc_profile()[36:40] == b"acsp"
        prof = run(os.path.join(icc, "icc_chunk_order.jpeg")).icc_profile()
        assert len(prof) == 254 and all(prof[i - 1] == i for i in range(1, 255))
        for name in ("icc_chunk_seq_no_0", "icc_chunk_double_seq_no", "icc_chunk_count_mismatch", "icc_missing_chunk"):
            assert run(os.path.join(icc, name + ".jpeg")).icc_profile() is None
        d = run(os.path.join(GOLDEN, "reftest", "ycck.jpg"))
        assert d.exif_data()[:8] == b"\x49\x49\x2A\x00\x08\x00\x00\x00"
        assert d.xmp_data()[:9] == b"<?xpacket"


def test_buffer_limit(oracle_mod, J):
    data = open(os.path.join(GOLDEN, "reftest", "mozilla", "jpg-size-8x8.jpg"), "rb").read()
    d = oracle_mod.Decoder(data)
    d.set_max_decoding_buffer_size(8 * 8 * 3 - 1)
    with pytest.raises(oracle_mod.OracleError) as e:
        d.decode()
    assert e.value.code == oracle_mod.ERR_FORMAT
    pd = J.Decoder(data)
    pd.set_max_decoding_buffer_size(8 * 8 * 3 - 1)
    with pytest.raises(J.B200JpgError) as e2:
        pd.entropy_decode()
    assert e2.value.code == J.ERR_FORMAT
    pd2 = J.Decoder(data)
    pd2.set_max_decoding_buffer_size(8 * 8 * 3)
    pd2.entropy_decode()


def test_read_info_files(oracle_mod, J):
    """b200jpg_read_info_files (host only): same ImageInfo as the oracle's read_info for every fixture."""
    paths = reftest_files(include_disabled=True) + bench_files()
    files = [open(p, "rb").read() for p in paths]
    res = J.read_info_files(files, nthreads=3)
    assert len(res) == len(files)
    for p, data, (st, info, out_len) in zip(paths, files, res):
        od = oracle_mod.Decoder(data)
        try:
            od.read_info()
            oi = od.info()
        except oracle_mod.OracleError as e:
            assert st == -e.code, p
            continue
        assert st == 0, p
        assert (info.width, info.height, info.pixel_format, info.coding_process) == (oi.width, oi.height, oi.pixel_format, oi.coding_process)
        assert out_len == info.width * info.height * {0: 1, 1: 1, 2: 3, 3: 4}[info.pixel_format]


def test_oracle_output_is_frozen(oracle_mod):
    """tests/golden/oracle_hashes.json (scripts/make_oracle_hashes.py) holds the sha256 of the oracle's pixels for every
    fixture in both arithmetic variants, taken when the oracle was validated against the reference's goldens: the exact
    pin that the +-3 PNG comparison cannot give.  ORC_ARITH_SSSE3_NATIVE (the timed CPU baseline) must hash the same as
    the emulation."""
    import hashlib
    import json
    pins = json.load(open(os.path.join(GOLDEN, "oracle_hashes.json")))
    assert len(pins) >= 40
    for rel, entry in pins.items():
        data = open(os.path.join(GOLDEN, rel), "rb").read()
        for name, a in (("scalar", oracle_mod.ARITH_SCALAR), ("ssse3", oracle_mod.ARITH_SSSE3), ("ssse3", oracle_mod.ARITH_SSSE3_NATIVE)):
            want = entry[name]
            try:
                px = oracle_mod.Decoder(data, a).decode()
                got = {"sha256": hashlib.sha256(px.tobytes()).hexdigest(), "bytes": int(px.size)}
            except oracle_mod.OracleError as e:
                got = {"error": int(e.code)}
            assert got == want, (rel, name)
