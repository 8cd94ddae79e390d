"""GPU parity: every kernel variant against the CPU oracle, through the C ABI.  Integer work: the bar
is BIT-EXACT (hence within the +-1 LSB per RGB channel BASELINE.json asks for) -- for the scalar
arithmetic against the scalar oracle and for the SSSE3-emulation mode against the SSSE3 oracle.
Run with `pytest -m gpu` on a B200."""
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, bench_files, cmyk_to_rgb, load_png_like_reftest, reftest_files
from test_oracle_kat import COEFS, EXPECTED, QT, SATURATED

pytestmark = pytest.mark.gpu

VARIANTS = [("scalar", "auto"), ("scalar", "generic"), ("scalar", "fast"), ("ssse3", "auto")]


@pytest.fixture(scope="module")
def ctxs(J):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    kern = {"auto": J.KERNEL_AUTO, "generic": J.KERNEL_GENERIC, "fast": J.KERNEL_FAST}
    out = {}
    for arith, k in VARIANTS:
        a = J.ARITH_SSSE3 if arith == "ssse3" else J.ARITH_SCALAR
        # "fast" pins K1 only: the K2 fast paths cover two layouts, the tests below choose per case
        k2 = J.KERNEL_GENERIC if k == "generic" else J.KERNEL_AUTO
        out[(arith, k)] = J.Context(device=0, arith=a, k1_kernel=kern[k], k2_kernel=k2)
    yield out
    for c in out.values():
        c.close()


def oarith(oracle_mod, arith):
    return oracle_mod.ARITH_SSSE3 if arith == "ssse3" else oracle_mod.ARITH_SCALAR


def copy_comps(J, ocomps):
    arr = (J.Component * len(ocomps))()
    for i, c in enumerate(ocomps):
        for f, _ in J.Component._fields_:
            setattr(arr[i], f, getattr(c, f))
    return arr


def gpu_planes(J, ctx, comps, qts, coefs, rows_per_call=None):
    """Worker::start / append_row(s) / get_result for each component."""
    w = J.Worker(ctx)
    planes = []
    for i, c in enumerate(comps):
        w.start(i, c, qts[i])
        per_row = c.block_w * c.v * 64
        a = np.ascontiguousarray(coefs[i], dtype=np.int16).reshape(-1)
        nrows = a.size // per_row
        if rows_per_call is None:
            w.append_rows(i, a, nrows)
        else:
            for r in range(nrows):
                w.append_row(i, a[r * per_row:(r + 1) * per_row])
        planes.append(w.get_result(i))
    return w, planes


def random_coefs(rng, nblocks, kind):
    if kind == "photo":      # sparse, small, like real images
        c = (rng.standard_normal((nblocks, 64)) * 40).astype(np.int16)
        c[rng.random((nblocks, 64)) > 0.25] = 0
        c[:, 0] = rng.integers(-1024, 1024, nblocks)
    elif kind == "dense":
        c = rng.integers(-2048, 2048, (nblocks, 64)).astype(np.int16)
    elif kind == "extreme":  # i16 extremes: i32 wrapping paths
        c = rng.choice(np.array([-32768, 32767, 0, 1, -1, 2047, -2047], dtype=np.int16), (nblocks, 64))
    elif kind == "dc_only":  # zero-AC columns everywhere, DC large: the column shortcut (src/idct.rs:279-295)
        c = np.zeros((nblocks, 64), dtype=np.int16)
        c[:, :8] = rng.integers(-32768, 32767, (nblocks, 8))
        mask = rng.random((nblocks, 8)) < 0.5
        c[:, :8][mask] = 0
    else:
        raise ValueError(kind)
    return c.reshape(-1)


# ---------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS)
def test_k1_known_answer_tests(J, oracle_mod, ctxs, variant):
    """src/idct.rs:580-657 through Worker::start/append_row/get_result on the device."""
    ctx = ctxs[variant]
    comps, _ = J.make_components(8, 8, [(1, 1)])
    _, (p,) = gpu_planes(J, ctx, comps, [QT], [COEFS])
    if variant[0] == "scalar":
        assert p.tolist() == EXPECTED
    else:
        assert np.abs(p.astype(int) - np.array(EXPECTED)).max() <= 1
        assert np.array_equal(p.reshape(8, 8), oracle_mod.idct_block(COEFS, QT, arith=1))
    _, (z,) = gpu_planes(J, ctx, comps, [[666] * 64], [[0] * 64])
    assert z.tolist() == [128] * 64
    _, (s,) = gpu_planes(J, ctx, comps, [[65535] * 64], [[32767] * 64])
    if variant[0] == "scalar":
        assert s.tolist() == SATURATED
    else:
        assert np.array_equal(s.reshape(8, 8), oracle_mod.idct_block([32767] * 64, [65535] * 64, arith=1))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("kind,qmax", [("photo", 255), ("dense", 255), ("extreme", 65535), ("dc_only", 65535), ("dense", 65535)])
def test_k1_random_blocks_bit_exact(J, oracle_mod, ctxs, variant, kind, qmax):
    """~10^5 seeded blocks per case, two components with different tables and ragged sizes (partial tiles,
    rows that are not a multiple of the 128-block tile)."""
    ctx = ctxs[variant]
    rng = np.random.default_rng(zlib.crc32(("%s-%d" % (kind, qmax)).encode()))
    w, h = 1160, 648   # Y: 146x82 blocks (not a multiple of 128), C: 73x41
    comps, _ = J.make_components(w, h, [(2, 2), (1, 1)])
    ocomps, _ = oracle_mod.make_components(w, h, [(2, 2), (1, 1)])
    qts = [rng.integers(1, qmax + 1, 64).astype(np.uint16) for _ in comps]
    coefs = [random_coefs(rng, c.block_w * c.block_h, kind) for c in comps]
    _, planes = gpu_planes(J, ctx, comps, qts, coefs)
    want = oracle_mod.idct_planes(ocomps, qts, coefs, arith=oarith(oracle_mod, variant[0]))
    for p, q in zip(planes, want):
        assert p.shape == q.shape
        assert np.array_equal(p, q), "max |diff| = %d" % np.abs(p.astype(int) - q.astype(int)).max()


@pytest.mark.parametrize("variant", [("scalar", "auto"), ("ssse3", "auto")])
@pytest.mark.parametrize("scale", [4, 2, 1])
def test_k1_scaled_idct(J, oracle_mod, ctxs, variant, scale):
    """dequantize_and_idct_block_{4x4,2x2,1x1}, src/idct.rs:456-565 (Decoder::scale path)."""
    ctx = ctxs[variant]
    rng = np.random.default_rng(scale)
    comps, _ = J.make_components(333, 200, [(2, 1), (1, 1)], dct_scale=scale)
    ocomps, _ = oracle_mod.make_components(333, 200, [(2, 1), (1, 1)], dct_scale=scale)
    qts = [rng.integers(1, 65536, 64).astype(np.uint16) for _ in comps]
    for kind in ("photo", "extreme"):
        coefs = [random_coefs(rng, c.block_w * c.block_h, kind) for c in comps]
        _, planes = gpu_planes(J, ctx, comps, qts, coefs)
        want = oracle_mod.idct_planes(ocomps, qts, coefs, arith=oarith(oracle_mod, variant[0]))
        for p, q in zip(planes, want):
            assert np.array_equal(p, q)


def test_k1_append_row_by_row_and_partial(J, oracle_mod, ctxs):
    """append_row one MCU row at a time == append_rows; MCU rows never appended stay 0 (src/worker/rayon.rs:46)."""
    ctx = ctxs[("scalar", "auto")]
    rng = np.random.default_rng(5)
    comps, _ = J.make_components(200, 120, [(1, 2)])
    ocomps, _ = oracle_mod.make_components(200, 120, [(1, 2)])
    qt = rng.integers(1, 256, 64).astype(np.uint16)
    coefs = random_coefs(rng, comps[0].block_w * comps[0].block_h, "photo")
    _, (a,) = gpu_planes(J, ctx, comps, [qt], [coefs], rows_per_call=1)
    (want,) = oracle_mod.idct_planes(ocomps, [qt], [coefs])
    assert np.array_equal(a, want)
    per_row = comps[0].block_w * comps[0].v * 64
    w = J.Worker(ctx)
    w.start(0, comps[0], qt)
    w.append_rows(0, coefs[:3 * per_row], 3)
    part = w.get_result(0)
    ow = oracle_mod.Worker()
    ow.start(0, ocomps[0], qt)
    for r in range(3):
        ow.append_row(0, coefs[r * per_row:(r + 1) * per_row])
    assert np.array_equal(part, ow.get_result(0))
    assert part[3 * per_row:].max() == 0


def test_worker_contract_errors(J, ctxs):
    """assert!s of src/worker/immediate.rs:31,47 become ERR_INTERNAL."""
    ctx = ctxs[("scalar", "auto")]
    comps, _ = J.make_components(16, 16, [(1, 1)])
    w = J.Worker(ctx)
    with pytest.raises(J.B200JpgError) as e:
        w.append_row(0, np.zeros(64 * 2, dtype=np.int16))
    assert e.value.code == J.ERR_INTERNAL
    w.start(0, comps[0], [1] * 64)
    with pytest.raises(J.B200JpgError):
        w.append_row(0, np.zeros(64 * 3, dtype=np.int16))   # wrong length
    with pytest.raises(J.B200JpgError):
        w.start(0, comps[0], [1] * 64)                        # started twice without get_result
    with pytest.raises(J.B200JpgError) as e2:
        w.compute_image(1, 16, 16, J.CT_GRAYSCALE)           # "not all components have data"
    assert e2.value.code == J.ERR_FORMAT


# ---------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------
SAMPLINGS = {
    "444": [(1, 1), (1, 1), (1, 1)], "420": [(2, 2), (1, 1), (1, 1)], "422": [(2, 1), (1, 1), (1, 1)],
    "440": [(1, 2), (1, 1), (1, 1)], "411": [(4, 1), (1, 1), (1, 1)], "generic_v": [(1, 4), (1, 1), (1, 2)],
    "mixed": [(2, 2), (2, 1), (1, 2)], "luma_sub": [(1, 1), (2, 2), (2, 2)],
}
SIZES = [(1, 1), (2, 2), (3, 5), (8, 8), (15, 15), (16, 16), (17, 17), (33, 31), (64, 48), (160, 2), (2, 160), (16, 1), (32, 3), (48, 7), (64, 9),
         (2064, 17), (240, 135), (256, 144), (641, 479)]


def random_planes(rng, comps):
    return [rng.integers(0, 256, c.block_w * c.block_h * c.dct_scale * c.dct_scale).astype(np.uint8) for c in comps]


@pytest.mark.parametrize("variant", [("scalar", "auto"), ("scalar", "generic"), ("ssse3", "auto")])
@pytest.mark.parametrize("sname", sorted(SAMPLINGS))
def test_k2_upsample_ycbcr_bit_exact(J, oracle_mod, ctxs, variant, sname):
    """compute_image with random planes: every upsampler (src/upsampler.rs:119-250), odd sizes, 1-pixel edges."""
    ctx = ctxs[variant]
    rng = np.random.default_rng(zlib.crc32(sname.encode()))
    for (w, h) in SIZES:
        comps, _ = J.make_components(w, h, SAMPLINGS[sname])
        ocomps, _ = oracle_mod.make_components(w, h, SAMPLINGS[sname])
        planes = random_planes(rng, comps)
        got = J.compute_image(ctx, comps, planes, w, h, J.CT_YCBCR)
        want = oracle_mod.compute_image(ocomps, planes, w, h, oracle_mod.CT_YCBCR, arith=oarith(oracle_mod, variant[0]))
        assert got.shape == want.shape
        assert np.array_equal(got, want), (sname, w, h, int(np.abs(got.astype(int) - want.astype(int)).max()))


@pytest.mark.parametrize("variant", [("scalar", "auto"), ("ssse3", "auto")])
def test_k2_colour_transforms(J, oracle_mod, ctxs, variant):
    """RGB / CMYK / YCCK / None / grayscale crop: src/decoder.rs:1310-1332, 1391-1484."""
    ctx = ctxs[variant]
    rng = np.random.default_rng(11)
    oa = oarith(oracle_mod, variant[0])
    for (w, h) in [(5, 3), (64, 40), (250, 99)]:
        for ncomp, sampling, cts in [(3, [(1, 1)] * 3, [J.CT_RGB, J.CT_NONE, J.CT_YCBCR]),
                                     (3, [(2, 2), (1, 1), (1, 1)], [J.CT_RGB]),
                                     (4, [(1, 1)] * 4, [J.CT_CMYK, J.CT_YCCK, J.CT_NONE]),
                                     (4, [(2, 2), (1, 1), (1, 1), (2, 2)], [J.CT_CMYK, J.CT_YCCK]),
                                     (1, [(1, 1)], [J.CT_GRAYSCALE]), (1, [(2, 2)], [J.CT_GRAYSCALE])]:
            comps, _ = J.make_components(w, h, sampling)
            ocomps, _ = oracle_mod.make_components(w, h, sampling)
            planes = random_planes(rng, comps)
            for ct in cts:
                got = J.compute_image(ctx, comps, planes, w, h, ct)
                want = oracle_mod.compute_image(ocomps, planes, w, h, ct, arith=oa)
                assert np.array_equal(got, want), (w, h, ncomp, sampling, ct)


@pytest.mark.parametrize("variant", [("scalar", "auto"), ("scalar", "generic"), ("ssse3", "auto")])
@pytest.mark.parametrize("name", ["h2v2", "h2v1", "h1v2"])
def test_k2_hand_derived_upsampling_vectors(J, ctxs, variant, name):
    """The kernels against bytes worked out by hand from src/upsampler.rs (tests/hand_vectors.py): no oracle involved."""
    from hand_vectors import planes_for
    comps, planes, w, h, want = planes_for(J.make_components, name, np.random.default_rng(1))
    got = J.compute_image(ctxs[variant], comps, planes, w, h, J.CT_RGB).reshape(h, w, 3)
    assert np.array_equal(got[..., 1], want), (name, got[..., 1])


def test_k2_error_mapping(J, oracle_mod, ctxs):
    """choose_color_convert_func / choose_upsampler errors (src/decoder.rs:1344-1386, src/upsampler.rs:93-98)."""
    ctx = ctxs[("scalar", "auto")]
    rng = np.random.default_rng(3)
    comps, _ = J.make_components(32, 32, [(1, 1)] * 3)
    planes = random_planes(rng, comps)
    for ct, code in [(J.CT_GRAYSCALE, J.ERR_FORMAT), (J.CT_CMYK, J.ERR_FORMAT), (J.CT_YCCK, J.ERR_FORMAT), (J.CT_UNKNOWN, J.ERR_FORMAT),
                     (J.CT_JCS_BG_YCC, J.ERR_UNSUPPORTED), (J.CT_JCS_BG_RGB, J.ERR_UNSUPPORTED)]:
        with pytest.raises(J.B200JpgError) as e:
            J.compute_image(ctx, comps, planes, 32, 32, ct)
        assert e.value.code == code
        with pytest.raises(oracle_mod.OracleError) as oe:
            oracle_mod.compute_image(oracle_mod.make_components(32, 32, [(1, 1)] * 3)[0], planes, 32, 32, ct)
        assert -oe.value.code == code
    comps, _ = J.make_components(32, 32, [(3, 1), (2, 1), (1, 1)])   # 3/2 is not an integer ratio
    with pytest.raises(J.B200JpgError) as e:
        J.compute_image(ctx, comps, random_planes(rng, comps), 32, 32, J.CT_YCBCR)
    assert e.value.code == J.ERR_UNSUPPORTED
    with pytest.raises(J.B200JpgError) as e:                          # "not all components have data"
        J.compute_image(ctx, comps, [np.zeros(0, np.uint8)] * 3, 32, 32, J.CT_YCBCR)
    assert e.value.code == J.ERR_FORMAT


# ---------------------------------------------------------------------------------------------
# whole files: Decoder::decode on the GPU vs the oracle (bit-exact) and vs the reference's goldens (+-3)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS)
def test_whole_files_bit_exact_and_golden(J, oracle_mod, ctxs, variant):
    ctx = ctxs[variant]   # "fast" pins the TMA-fed K1, which takes every scale and both arithmetic variants since round 2
    oa = oarith(oracle_mod, variant[0])
    n = 0
    for p in reftest_files() + bench_files():
        data = open(p, "rb").read()
        od = oracle_mod.Decoder(data, oa)
        want = od.decode()
        d = J.Decoder(data, ctx)
        got = d.decode()
        info, oinfo = d.info(), od.info()
        assert (info.width, info.height, info.pixel_format, info.coding_process) == (oinfo.width, oinfo.height, oinfo.pixel_format, oinfo.coding_process)
        assert np.array_equal(got, want), (p, int(np.abs(got.astype(int) - want.astype(int)).max()))
        png = p[:-4] + ".png"
        if os.path.exists(png):
            px = cmyk_to_rgb(got) if info.pixel_format == 3 else got
            ref = load_png_like_reftest(png, info.pixel_format)
            assert np.abs(ref.astype(int) - px.astype(int)).max() <= 3, p   # tests/reftest/mod.rs:99
        n += 1
    assert n >= 37


def test_scalar_vs_ssse3_gap_is_the_references_own(J, ctxs):
    """The two arithmetic variants of the reference differ by a few LSB; both are reproduced."""
    data = open(os.path.join(GOLDEN, "benches", "tower.jpg"), "rb").read()
    a = J.Decoder(data, ctxs[("scalar", "auto")]).decode().astype(int)
    b = J.Decoder(data, ctxs[("ssse3", "auto")]).decode().astype(int)
    d = np.abs(a - b)
    assert 0 < d.max() <= 4


@pytest.mark.parametrize("req,png", [((500, 333), "rgb.png"), ((250, 167), "rgb_250x167.png"), ((125, 84), "rgb_125x84.png"), ((63, 42), "rgb_63x42.png")])
def test_scaled_decode(J, oracle_mod, ctxs, req, png):
    """tests/reftest/mod.rs:18-25 on the GPU."""
    root = os.path.join(GOLDEN, "reftest")
    data = open(os.path.join(root, "rgb.jpg"), "rb").read()
    d = J.Decoder(data, ctxs[("scalar", "auto")])
    d.read_info()
    assert d.scale(*req) == req
    got = d.decode()
    od = oracle_mod.Decoder(data)
    od.read_info()
    od.scale(*req)
    assert np.array_equal(got, od.decode())
    assert np.abs(load_png_like_reftest(os.path.join(root, png), 2).astype(int) - got.astype(int)).max() <= 3


def test_read_info_then_decode_gpu(J, ctxs):
    """tests/lib.rs:34-50"""
    data = open(os.path.join(GOLDEN, "reftest", "mozilla", "jpg-progressive.jpg"), "rb").read()
    ctx = ctxs[("scalar", "auto")]
    ref = J.Decoder(data, ctx).decode()
    d = J.Decoder(data, ctx)
    d.read_info()
    assert np.array_equal(d.decode(), ref)


# ---------------------------------------------------------------------------------------------
# batches
# ---------------------------------------------------------------------------------------------
def test_heterogeneous_batch_with_bad_image(J, oracle_mod, ctxs):
    """decode_batch over mixed geometries; a bad image gets its own status and does not poison the batch."""
    from jpeg_decoder_b200 import workload
    ctx = ctxs[("scalar", "auto")]
    keep, descs, wants = [], [], []
    jobs = [workload.synth_jpeg(320, 176, 1, 2), workload.synth_jpeg(96, 64, 2, 0), workload.synth_jpeg(161, 99, 3, 1),
            open(os.path.join(GOLDEN, "benches", "tower_grayscale.jpg"), "rb").read(), workload.synth_jpeg(640, 368, 4, 2, progressive=True)]
    decs = []
    for data in jobs:
        d = J.Decoder(data)
        decs.append(d)
        descs.append(d.entropy_decode())
        wants.append(oracle_mod.Decoder(data).decode())
    bad = J.ImageDesc()
    for f, _ in J.ImageDesc._fields_:
        setattr(bad, f, getattr(descs[0], f))
    bad.color_transform = J.CT_CMYK          # 3 components + CMYK: Format error
    descs.insert(2, bad)
    outs, st = J.decode_batch(ctx, descs)
    assert st[2] == J.ERR_FORMAT and [s for i, s in enumerate(st) if i != 2] == [0] * 5
    del outs[2]
    for got, want in zip(outs, wants):
        assert np.array_equal(got[:want.size], want)


@pytest.mark.parametrize("k1", ["generic", "fast"])
def test_full_size_batch_properties(J, oracle_mod, ctxs, k1):
    """BASELINE cfg2 geometry (1920x1080 4:2:0) at a reduced batch, device-resident path as bench.py uses it:
    every image bit-exact vs the oracle, and a checksum-of-checksums over the replicated batch."""
    import torch
    from jpeg_decoder_b200 import workload
    ctx = ctxs[("scalar", k1)]
    uniq = workload.build_unique("cfg2", 2)
    B = 12
    keep, descs = [], []
    for i in range(B):
        u = uniq[i % 2]
        descs.append(J.make_image_desc(u.width, u.height, u.components, u.qts, u.coefs, u.color_transform, keep))
    batch = J.Batch(ctx, descs)
    info = batch.info
    assert info.n_blocks == B * 48960 and info.n_pixels == B * 1920 * 1080
    assert info.k1_algorithmic_bytes == B * 9400320 and info.k2_algorithmic_bytes == B * 9354240
    dev = torch.device("cuda", 0)
    d_coefs = torch.zeros(info.coef_bytes, dtype=torch.uint8, device=dev)
    d_planes = torch.zeros(info.plane_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(info.out_bytes, dtype=torch.uint8, device=dev)
    for j in range(B):
        lay = batch.layout(j)
        for k, c in enumerate(uniq[j % 2].coefs):
            t = torch.from_numpy(c.view(np.uint8)).to(dev)
            d_coefs[lay["coef_off"][k]:lay["coef_off"][k] + t.numel()].copy_(t)
    torch.cuda.synchronize()
    batch.run_device(d_coefs.data_ptr(), d_planes.data_ptr(), d_out.data_ptr(), 3)
    ctx.synchronize()
    torch.cuda.synchronize()
    wants = [oracle_mod.hotpath_image(copy_ocomps(oracle_mod, u.components), u.qts, u.coefs, u.width, u.height, u.color_transform) for u in uniq]
    sums = []
    for j in range(B):
        lay = batch.layout(j)
        got = d_out[lay["out_off"]:lay["out_off"] + lay["out_len"]].cpu().numpy()
        assert np.array_equal(got, wants[j % 2]), j
        sums.append(int(got.astype(np.uint64).sum()))
    assert sums[0::2] == [sums[0]] * (B // 2) and sums[1::2] == [sums[1]] * (B // 2)
    # same batch through the host pipeline (H2D, K1, K2, D2H) gives the same bytes
    outs = [np.zeros(batch.layout(j)["out_len"], dtype=np.uint8) for j in range(B)]
    assert batch.run_host(outs) == [0] * B
    for j in range(B):
        assert np.array_equal(outs[j], wants[j % 2])
    batch.close()


BULK_SIZES = [(32, 1), (32, 2), (32, 37), (64, 9), (96, 200), (2048, 5), (2080, 4), (4096, 3), (4128, 37), (6176, 11), (1920, 1080)]


@pytest.mark.parametrize("k2_mode", [0, 1, 4])
def test_k2_420_bulk_copy_kernel_geometries(J, oracle_mod, ctxs, k2_mode):
    """The bulk-copy fed 4:2:0 kernel (k2_ycbcr420_tma: strips of 2048 pixels, row pairs chained through the shared
    memory ring) against the oracle at the geometries that stress it: one row, fewer row pairs than ring stages, several
    strips with a short last strip, odd heights; the load/store kernels (modes 1 and 4) must give the same bytes."""
    ctx = ctxs[("scalar", "auto")]
    rng = np.random.default_rng(420)
    J.lib().b200jpg_debug_set_kernel_modes(-1, k2_mode)
    try:
        for (w, h) in BULK_SIZES:
            comps, _ = J.make_components(w, h, SAMPLINGS["420"])
            ocomps, _ = oracle_mod.make_components(w, h, SAMPLINGS["420"])
            planes = random_planes(rng, comps)
            got = J.compute_image(ctx, comps, planes, w, h, J.CT_YCBCR)
            want = oracle_mod.compute_image(ocomps, planes, w, h, oracle_mod.CT_YCBCR)
            assert np.array_equal(got, want), (w, h, k2_mode)
    finally:
        J.lib().b200jpg_debug_set_kernel_modes(-1, -1)


def test_k2_420_bulk_copy_mixed_batch(J, oracle_mod, ctxs):
    """One batch mixing images on the bulk-copy path with ragged / 4:4:4 / gray ones: the strip table only lists the
    eligible images and the per-CTA ranges cross image boundaries."""
    from jpeg_decoder_b200 import workload
    ctx = ctxs[("scalar", "auto")]
    datas = [workload.synth_jpeg(w, h, 3000 + i, sub) for i, (w, h, sub) in enumerate(
        [(64, 40, 2), (2080, 36, 2), (50, 30, 2), (96, 64, 0), (4128, 20, 2), (32, 2, 2), (640, 480, 2)])]
    descs, keep = [], []
    for d in datas:
        dec = J.Decoder(d)
        keep.append(dec)
        descs.append(dec.entropy_decode())
    outs, st = J.decode_batch(ctx, descs * 3)
    assert st == [0] * (3 * len(datas))
    wants = [oracle_mod.Decoder(d).decode() for d in datas]
    for j, o in enumerate(outs):
        assert np.array_equal(o, wants[j % len(datas)]), j


def test_k1_more_than_four_quant_tables(J, oracle_mod, ctxs):
    """Only the first four 8-bit tables of a batch ride in the kernel-parameter constant bank; later ones are read
    from the packed table in global memory (qflags slot 0xff).  Six images with distinct tables."""
    from jpeg_decoder_b200 import workload
    ctx = ctxs[("scalar", "auto")]
    datas = [workload.synth_jpeg(160, 96, 100 + q, 2, quality=q) for q in (35, 50, 65, 75, 85, 92)]
    descs, keep = [], []
    for d in datas:
        dec = J.Decoder(d)
        keep.append(dec)
        descs.append(dec.entropy_decode())
    outs, st = J.decode_batch(ctx, descs)
    assert st == [0] * len(datas)
    for d, o in zip(datas, outs):
        assert np.array_equal(o, oracle_mod.Decoder(d).decode())


def copy_ocomps(oracle_mod, comps):
    arr = (oracle_mod.Component * len(comps))()
    for i, c in enumerate(comps):
        for f, _ in oracle_mod.Component._fields_:
            setattr(arr[i], f, getattr(c, f))
    return arr


def test_cfg3_geometry_444_fast_path(J, oracle_mod, ctxs):
    """BASELINE cfg3 geometry (3840x2160 4:4:4, no upsample path), one image, bit-exact."""
    from jpeg_decoder_b200 import workload
    data = workload.synth_jpeg(3840, 2160, 5000, 0)
    got = J.Decoder(data, ctxs[("scalar", "auto")]).decode()
    assert np.array_equal(got, oracle_mod.Decoder(data).decode())


def test_progressive_x_many(J, oracle_mod, ctxs):
    """BASELINE cfg4: tower_progressive.jpg replicated (host re-feeds coefficients once per component)."""
    data = open(os.path.join(GOLDEN, "benches", "tower_progressive.jpg"), "rb").read()
    want = oracle_mod.Decoder(data).decode()
    d = J.Decoder(data)
    desc = d.entropy_decode()
    outs, st = J.decode_batch(ctxs[("scalar", "auto")], [desc] * 16)
    assert st == [0] * 16
    for o in outs:
        assert np.array_equal(o, want)


def test_decode_files_batch(J, oracle_mod, ctxs):
    """b200jpg_decode_files: multi-threaded host Huffman + GPU worker path over every reference fixture, with broken
    files mixed in; per-image statuses, pixels bit-exact vs the oracle."""
    import glob
    ctx = ctxs[("scalar", "auto")]
    paths = reftest_files(include_disabled=True) + bench_files()
    paths += sorted(glob.glob(os.path.join(GOLDEN, "crashtest", "*.jpg")))[:6]
    files = [open(p, "rb").read() for p in paths]
    wants = []
    for data in files:
        try:
            wants.append(oracle_mod.Decoder(data).decode())
        except oracle_mod.OracleError as e:
            wants.append(-e.code)
    for nthreads in (1, 5):
        outs, st, infos = J.decode_files(ctx, files, nthreads=nthreads)
        for p, o, s_, w in zip(paths, outs, st, wants):
            if isinstance(w, int):
                assert s_ == w, (p, s_, w)
            else:
                assert s_ == 0, (p, s_)
                assert np.array_equal(o, w), p
    # many copies of one file: exercises chunking and the double-buffered arenas
    data = open(os.path.join(GOLDEN, "benches", "tower.jpg"), "rb").read()
    want = oracle_mod.Decoder(data).decode()
    outs, st, _ = J.decode_files(ctx, [data] * 70, nthreads=3)
    assert st == [0] * 70 and all(np.array_equal(o, want) for o in outs)


# ---------------------------------------------------------------------------------------------
# K0: sparse block streams (csrc/sbs.h, csrc/k0_expand.cu)
# ---------------------------------------------------------------------------------------------
def test_k0_expand_rebuilds_the_dense_slab(J, ctxs):
    """Kernel K0 turns the host decoder's sparse stream back into exactly the dense coefficients the reference pushes
    through Worker::append_row -- for every fixture (planar and interleaved order, wide blocks, odd geometries)."""
    from jpeg_decoder_b200 import workload
    ctx = ctxs[("scalar", "auto")]
    datas = [open(p, "rb").read() for p in reftest_files() + bench_files()]
    datas += [workload.synth_jpeg(w, h, seed=3, subsampling=s, progressive=pr)
              for (w, h, s, pr) in [(1920, 1080, 2, False), (97, 61, 2, False), (640, 360, 0, False), (320, 200, 1, False), (200, 120, 2, True)]]
    orders = set()
    for data in datas:
        ref = J.Decoder(data)
        try:
            d0 = ref.entropy_decode()
        except J.B200JpgError:
            continue
        if any(not d0.coefs[c] for c in range(d0.ncomp)):
            continue
        d1, buf, order = J.Decoder(data).entropy_decode_sbs()
        orders.add(order)
        dense = J.expand_sbs(ctx, d1, buf, order)
        for c in range(d0.ncomp):
            assert np.array_equal(dense[c], ref.coefficients(d0, c)), (len(data), c)
    assert orders == {0, 1}


@pytest.mark.parametrize("variant", [("scalar", "auto"), ("ssse3", "auto")])
def test_decode_batch_sbs_bit_exact(J, oracle_mod, ctxs, variant):
    """Streams in, pixels out (H2D -> K0 -> K1 -> K2 -> D2H over three CUDA streams): bit-exact vs the oracle, more
    images than one group, a malformed stream is rejected without poisoning the batch."""
    from jpeg_decoder_b200 import workload
    ctx = ctxs[variant]
    datas = [open(p, "rb").read() for p in bench_files()]
    datas += [workload.synth_jpeg(w, h, seed=21 + i, subsampling=s) for i, (w, h, s) in enumerate([(256, 144, 2), (130, 70, 0), (48, 48, 1)])]
    datas = datas * 6  # > 32 images: several groups
    descs, streams, wants, keep = [], [], [], []
    for data in datas:
        dec = J.Decoder(data)
        d, buf, order = dec.entropy_decode_sbs()
        keep.append(dec)
        descs.append(d)
        streams.append((buf, order))
        wants.append(oracle_mod.Decoder(data, oarith(oracle_mod, variant[0])).decode())
    bad = streams[5][0].copy()
    bad[0] ^= 0x10  # a bitmap bit without a value: offsets no longer add up
    streams[5] = (bad, streams[5][1])
    outs, st = J.decode_batch_sbs(ctx, descs, streams)
    for i, (o, s_, w) in enumerate(zip(outs, st, wants)):
        if i == 5:
            assert s_ == J.ERR_INTERNAL
        else:
            assert s_ == 0 and np.array_equal(o, w), i


def test_decode_files_streaming_engine(J, oracle_mod, ctxs):
    """The whole-file engine under load: many 1080p-class images over few and many host threads (ring wrap-around,
    ring growth when a larger image follows smaller ones, more threads than images), results bit-exact."""
    from jpeg_decoder_b200 import workload
    ctx = ctxs[("scalar", "auto")]
    small = [workload.synth_jpeg(320, 240, seed=40 + i, subsampling=2) for i in range(3)]
    big = [workload.synth_jpeg(1920, 1080, seed=50 + i, subsampling=2) for i in range(2)]
    prog = workload.synth_jpeg(640, 480, seed=60, subsampling=0, progressive=True)
    files = (small * 5 + big + [prog] + small + big * 9 + [prog] * 3 + [b"\xff\xd8\xff\xd9"] + small * 4)
    cache = {}
    for f in files:
        if f not in cache:
            try:
                cache[f] = oracle_mod.Decoder(f).decode()
            except oracle_mod.OracleError as e:
                cache[f] = -e.code
    for nthreads in (1, 3, 16, 200):
        outs, st, _ = J.decode_files(ctx, files, nthreads=nthreads)
        for i, (f, o, s_) in enumerate(zip(files, outs, st)):
            w = cache[f]
            if isinstance(w, int):
                assert s_ == w, (nthreads, i, s_, w)
            else:
                assert s_ == 0 and np.array_equal(o, w), (nthreads, i, s_)


def test_run_host_compaction_equals_dense_upload(J, oracle_mod):
    """b200jpg_batch_run_host gives the same pixels whether the dense coefficients are uploaded as they are
    (COMPACT_OFF) or compacted into sparse block streams by host threads first (COMPACT_ON): mixed geometries, a
    grey image, a progressive one, an image the planner rejects and one with a missing component."""
    from jpeg_decoder_b200 import workload
    jobs = [workload.synth_jpeg(320, 176, 1, 2), workload.synth_jpeg(96, 64, 2, 0), workload.synth_jpeg(161, 99, 3, 1),
            open(os.path.join(GOLDEN, "benches", "tower_grayscale.jpg"), "rb").read(), workload.synth_jpeg(640, 368, 4, 2, progressive=True),
            workload.synth_jpeg(1920, 1080, 5, 2)]
    decs, descs, wants = [], [], []
    for data in jobs * 9:
        d = J.Decoder(data)
        decs.append(d)
        descs.append(d.entropy_decode())
        wants.append(oracle_mod.Decoder(data).decode())
    bad = J.ImageDesc()
    for f, _ in J.ImageDesc._fields_:
        setattr(bad, f, getattr(descs[0], f))
    bad.color_transform = J.CT_CMYK          # 3 components + CMYK: Format error
    descs.insert(2, bad)
    wants.insert(2, J.ERR_FORMAT)
    hole = J.ImageDesc()
    for f, _ in J.ImageDesc._fields_:
        setattr(hole, f, getattr(descs[1], f))
    hole.coefs[1] = None                     # "not all components have data"
    descs.insert(7, hole)
    wants.insert(7, J.ERR_FORMAT)
    results = []
    for mode, threads in ((J.COMPACT_OFF, 0), (J.COMPACT_ON, 0), (J.COMPACT_ON, 1), (J.COMPACT_ON, 5)):
        ctx = J.Context(device=0, host_compact=mode, host_threads=threads)
        try:
            batch = J.Batch(ctx, descs)
            outs = [np.zeros(int(d.width) * int(d.height) * int(d.ncomp), dtype=np.uint8) for d in descs]
            try:
                st = batch.run_host(outs)
            except J.B200JpgError:
                st = list(batch.statuses)
            for i, (o, s_, w) in enumerate(zip(outs, st, wants)):
                if isinstance(w, int):
                    assert s_ == w, (mode, i, s_)
                else:
                    assert s_ == 0 and np.array_equal(o, w), (mode, threads, i, s_)
            batch.close()
        finally:
            ctx.close()
