"""The sparse block stream (csrc/sbs.h) written by the host entropy decoder expands -- with an independent numpy
reading of the format -- to exactly the dense coefficients the same decoder (and hence the oracle's feeder, see
test_oracle_golden.py) produces, for every reftest / bench file and for synthetic 4:2:0 / 4:4:4 / grey images,
in both block orders (interleaved = straight from the Huffman loop, planar = compacted dense buffers).  CPU only."""
import numpy as np
import pytest

from conftest import bench_files, reftest_files

UNZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                     28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                     54, 47, 55, 62, 63])


def expand_stream(desc, buf, order):
    """numpy restatement of kernel K0: stream -> per-component dense arrays (raster blocks, natural order)."""
    comps = [desc.comps[c] for c in range(desc.ncomp)]
    nbs = [int(c.block_w) * int(c.block_h) for c in comps]
    nb = sum(nbs)
    nb_pad = (nb + 31) // 32 * 32
    off_dc, off_voff = 8 * nb_pad, 10 * nb_pad
    off_vals = (off_voff + 4 * (nb_pad // 32 + 1) + 15) // 16 * 16
    bm = buf[:8 * nb_pad].view(np.uint64)
    dc = buf[off_dc:off_dc + 2 * nb_pad].view(np.int16)
    voff = buf[off_voff:off_voff + 4 * (nb_pad // 32 + 1)].view(np.uint32)
    vals = buf[off_vals:]
    dense = [np.zeros((n, 64), dtype=np.int16) for n in nbs]
    # scan-order block -> (component, raster block)
    where = []
    if order & 1 == 0:
        for c, n in enumerate(nbs):
            where += [(c, b) for b in range(n)]
    else:
        mcu_w = int(comps[0].block_w) // int(comps[0].h)
        mcu_h = int(comps[0].block_h) // int(comps[0].v)
        for my in range(mcu_h):
            for mx in range(mcu_w):
                for c, cp in enumerate(comps):
                    for vy in range(int(cp.v)):
                        for hx in range(int(cp.h)):
                            where.append((c, (my * int(cp.v) + vy) * int(cp.block_w) + mx * int(cp.h) + hx))
    assert len(where) == nb
    natural = bool(order & 2)
    at = 0
    for t in range(nb_pad):
        if t % 32 == 0:
            assert voff[t // 32] == at
        m = int(bm[t])
        wide = m & 1
        if t >= nb:
            assert m == 0
            continue
        c, b = where[t]
        dense[c][b, 0] = dc[t]
        for k in range(1, 64):
            if (m >> k) & 1:
                if wide:
                    v = np.array([vals[at], vals[at + 1]], dtype=np.uint8).view(np.int16)[0]
                    at += 2
                else:
                    v = np.int8(vals[at])
                    at += 1
                dense[c][b, k if natural else UNZIGZAG[k]] = v
    assert voff[nb_pad // 32] == at and off_vals + at <= buf.size
    return [d.reshape(-1) for d in dense]


def check(J, data, want_order=None):
    ref = J.Decoder(data)
    d0 = ref.entropy_decode()
    dec = J.Decoder(data)
    d1, buf, order = dec.entropy_decode_sbs()
    if want_order is not None:
        assert order == want_order
    assert d1.ncomp == d0.ncomp and buf.size % 16 == 0
    got = expand_stream(d1, buf, order)
    for c in range(d0.ncomp):
        assert np.array_equal(got[c], ref.coefficients(d0, c)), c
        assert np.array_equal(dec.qtable(d1, c), ref.qtable(d0, c))
    return order, buf.size, sum(g.size * 2 for g in got)


def test_streams_of_every_fixture(J):
    orders = set()
    for p in reftest_files() + bench_files():
        data = open(p, "rb").read()
        try:
            J.Decoder(data).entropy_decode()
        except J.B200JpgError:
            with pytest.raises(J.B200JpgError):
                J.Decoder(data).entropy_decode_sbs()
            continue
        orders.add(check(J, data)[0])
    assert orders == {0, 1}  # both block orders occur in the corpus


@pytest.mark.parametrize("w,h,sub,prog", [(256, 144, 2, False), (97, 61, 2, False), (64, 48, 0, False), (40, 24, 1, False),
                                          (96, 80, 0, True), (33, 17, 2, True)])
def test_streams_of_synthetic_images(J, w, h, sub, prog):
    from jpeg_decoder_b200 import workload
    data = workload.synth_jpeg(w, h, seed=11, subsampling=sub, progressive=prog)
    order, slen, dlen = check(J, data, want_order=0 if prog else 1)
    assert slen < dlen  # the stream is smaller than the dense coefficients


def test_wide_values_and_grey(J):
    """quality 100 forces AC values beyond int8 (wide blocks); a 1-component image takes the direct path too."""
    import io
    from PIL import Image
    rng = np.random.default_rng(5)
    img = (rng.integers(0, 2, size=(64, 72)) * 255).astype(np.uint8)
    for mode_img in (Image.fromarray(img, "L"), Image.fromarray(np.stack([img, img, 255 - img], axis=-1), "RGB")):
        b = io.BytesIO()
        mode_img.save(b, "JPEG", quality=100, subsampling=0)
        data = b.getvalue()
        dec = J.Decoder(data)
        d, buf, order = dec.entropy_decode_sbs()
        nb = sum(int(d.comps[c].block_w) * int(d.comps[c].block_h) for c in range(d.ncomp))
        bm = buf[:8 * ((nb + 31) // 32 * 32)].view(np.uint64)
        assert (bm & np.uint64(1)).any()  # some block is wide
        check(J, data, want_order=1)


def test_compacted_dense_buffers(J):
    """b200jpg_sbs_from_dense (what b200jpg_batch_run_host's host threads do): PLANAR | NATURAL streams of real and of
    adversarial coefficient buffers expand to the same coefficients."""
    from jpeg_decoder_b200 import workload
    for data in [workload.synth_jpeg(200, 120, seed=2, subsampling=2), open(bench_files()[0], "rb").read()]:
        dec = J.Decoder(data)
        d = dec.entropy_decode()
        buf, order = J.sbs_from_dense(d)
        assert order == 2
        got = expand_stream(d, buf, order)
        for c in range(d.ncomp):
            assert np.array_equal(got[c], dec.coefficients(d, c))
    # adversarial: every value class next to each other (zero, int8 limits, just beyond them, int16 limits)
    rng = np.random.default_rng(9)
    comps, _ = J.make_components(64, 48, [(1, 1)] * 3)
    keep = []
    coefs = [rng.choice(np.array([0, 0, 0, 1, -1, 127, -128, 128, -129, 32767, -32768], dtype=np.int16), 48 * 64) for _ in range(3)]
    coefs[1][:64 * 10] = 0  # all-zero blocks
    qts = [np.full(64, 1, dtype=np.uint16)] * 3
    d = J.make_image_desc(64, 48, comps, qts, coefs, J.CT_YCBCR, keep)
    buf, order = J.sbs_from_dense(d)
    got = expand_stream(d, buf, order)
    for c in range(3):
        assert np.array_equal(got[c], coefs[c])


def test_compaction_sse2_body_matches_avx512_body():
    """The two bodies of SbsWriter::put_dense_natural_run write identical streams (B200JPG_NO_AVX512 selects SSE2)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, zlib, numpy as np; sys.path.insert(0, %r); import jpeg_decoder_b200 as J; from jpeg_decoder_b200 import workload;"
            "rng = np.random.default_rng(3); out = [];\n"
            "for data in [workload.synth_jpeg(300, 200, seed=8, subsampling=2), workload.synth_jpeg(64, 64, seed=9, subsampling=0)]:\n"
            "    dec = J.Decoder(data); d = dec.entropy_decode(); buf, order = J.sbs_from_dense(d); out.append(zlib.crc32(buf.tobytes()))\n"
            "comps, _ = J.make_components(64, 48, [(1, 1)] * 3); keep = []\n"
            "coefs = [rng.choice(np.array([0, 0, 0, 1, -1, 127, -128, 128, -129, 32767, -32768], dtype=np.int16), 48 * 64) for _ in range(3)]\n"
            "d = J.make_image_desc(64, 48, comps, [np.full(64, 1, dtype=np.uint16)] * 3, coefs, J.CT_YCBCR, keep)\n"
            "buf, order = J.sbs_from_dense(d); out.append(zlib.crc32(buf.tobytes())); print(out)") % root
    runs = []
    for env_extra in ({}, {"B200JPG_NO_AVX512": "1"}):
        env = dict(os.environ, **env_extra)
        runs.append(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300))
        assert runs[-1].returncode == 0, runs[-1].stderr
    assert runs[0].stdout == runs[1].stdout and runs[0].stdout.startswith("[")
