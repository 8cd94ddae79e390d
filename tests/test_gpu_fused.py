"""GPU parity of the fused kernel KF (csrc/kf_fused.cu: dequantise + IDCT + upsample + colour in one pass, SURVEY
section 8 row f4) through the C ABI: bit-exact against the CPU oracle's whole hot path (start + append_row* +
get_result + compute_image, oracle/ref_image.c) and against the two-kernel route (K1 then K2) of the same library.
Geometries chosen to stress what is new in it: column strips with a recomputed chroma halo (widths > 1920), MCU rows
carried across items, CTA ranges that start in the middle of a column (many small images), ragged widths and heights,
blocks that take the exact slow path (|c*q| >= 2^19), 16-bit tables, more than four tables per batch."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAMPLING = {"420": [(2, 2), (1, 1), (1, 1)], "444": [(1, 1), (1, 1), (1, 1)]}
SIZES = [(1, 1), (2, 2), (3, 5), (8, 8), (15, 15), (16, 16), (17, 17), (33, 31), (64, 48), (160, 2), (2, 160), (16, 1), (32, 3), (48, 7),
         (240, 135), (641, 479), (1920, 40), (1921, 33), (1936, 17), (2000, 16), (3840, 35), (3857, 18), (4100, 20), (8200, 9)]


@pytest.fixture(scope="module")
def fctx(J):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    a, b = J.Context(device=0, fuse=J.FUSE_ON), J.Context(device=0, fuse=J.FUSE_OFF)
    yield a, b
    a.close()
    b.close()


def coefs_for(rng, nblocks, kind):
    if kind == "photo":
        c = (rng.standard_normal((nblocks, 64)) * 40).astype(np.int16)
        c[rng.random((nblocks, 64)) > 0.25] = 0
        c[:, 0] = rng.integers(-1024, 1024, nblocks)
    elif kind == "dense":
        c = rng.integers(-2048, 2048, (nblocks, 64)).astype(np.int16)
    else:  # "extreme": i16 extremes + DC-only columns: wrapping arithmetic and the zero-AC column shortcut
        c = rng.choice(np.array([-32768, 32767, 0, 0, 0, 1, -1, 2047, -2047], dtype=np.int16), (nblocks, 64))
        c[rng.random(nblocks) < 0.3, 8:] = 0
    return c.reshape(-1)


def ocomps_of(oracle_mod, comps):
    arr = (oracle_mod.Component * len(comps))()
    for i, c in enumerate(comps):
        for f, _ in oracle_mod.Component._fields_:
            setattr(arr[i], f, getattr(c, f))
    return arr


def run_images(J, oracle_mod, fctx, images):
    """images: list of (w, h, sampling name, qts, coefs).  Fused == two-kernel == oracle, image by image."""
    fused, split = fctx
    keep, descs, wants = [], [], []
    for (w, h, sname, qts, coefs) in images:
        comps, _ = J.make_components(w, h, SAMPLING[sname])
        descs.append(J.make_image_desc(w, h, comps, qts, coefs, J.CT_YCBCR, keep))
        wants.append(oracle_mod.hotpath_image(ocomps_of(oracle_mod, comps), qts, coefs, w, h, oracle_mod.CT_YCBCR))
    # 4:2:0 with a 1-pixel dimension is not "H2V2" for the reference (choose_upsampler, src/upsampler.rs:84-85) -> K1 + K2
    n_eligible = sum(1 for (w, h, sname, _, _) in images if sname == "444" or (w > 1 and h > 1))
    for ctx, expect_fused in ((fused, n_eligible), (split, None)):
        batch = J.Batch(ctx, descs)
        if expect_fused is not None:
            assert batch.info.n_fused == expect_fused
        outs = [np.zeros(w * h * 3, dtype=np.uint8) for (w, h, _, _, _) in images]
        l0 = ctx.launch_count
        assert batch.run_host(outs) == [0] * len(images)
        launches = ctx.launch_count - l0
        batch.close()
        for i, (got, want) in enumerate(zip(outs, wants)):
            assert np.array_equal(got, want), (images[i][:3], "fused" if expect_fused is not None else "k1+k2",
                                               int(np.abs(got.astype(int) - want.astype(int)).max()))
        if expect_fused == len(images):
            assert launches <= 2, launches   # one KF launch per sampling mode, no K1 / K2


def make_image(J, rng, w, h, sname, kind, qmax):
    comps, _ = J.make_components(w, h, SAMPLING[sname])
    qy = rng.integers(1, qmax + 1, 64).astype(np.uint16)
    qc = rng.integers(1, qmax + 1, 64).astype(np.uint16)
    coefs = [coefs_for(rng, c.block_w * c.block_h, kind) for c in comps]
    return (w, h, sname, [qy, qc, qc], coefs)


@pytest.mark.parametrize("sname", ["420", "444"])
@pytest.mark.parametrize("kind,qmax", [("photo", 255), ("dense", 255), ("extreme", 65535)])
def test_fused_geometries_bit_exact(J, oracle_mod, fctx, sname, kind, qmax):
    rng = np.random.default_rng(zlib.crc32(("kf-%s-%s" % (sname, kind)).encode()))
    for (w, h) in SIZES:
        run_images(J, oracle_mod, fctx, [make_image(J, rng, w, h, sname, kind, qmax)])


def test_fused_many_small_images_and_mixed_batch(J, oracle_mod, fctx):
    """More items than CTAs and columns shorter than a CTA's range: ranges start inside columns (the 4:2:0 warm-up
    item) and cross image boundaries; both sampling modes, 8- and 16-bit tables and six distinct tables in one batch."""
    rng = np.random.default_rng(77)
    images = []
    for i in range(420):
        sname = "420" if i % 3 else "444"
        w, h = [(64, 48), (80, 112), (33, 170), (200, 24)][i % 4]
        images.append(make_image(J, rng, w, h, sname, "photo", 255 if i % 5 else 4000))
    run_images(J, oracle_mod, fctx, images)


def test_fused_full_size(J, oracle_mod, fctx):
    """BASELINE geometries: 1920x1080 4:2:0 (67.5 MCU rows: the last one is half visible) and 3840x2160 4:4:4."""
    rng = np.random.default_rng(1080)
    run_images(J, oracle_mod, fctx, [make_image(J, rng, 1920, 1080, "420", "photo", 255) for _ in range(3)] +
               [make_image(J, rng, 3840, 2160, "444", "photo", 255)])


def test_fused_not_taken_when_not_eligible(J, fctx):
    """4:2:2, RGB transform, scaled IDCT: the plan keeps them on K1 + K2 (n_fused == 0) and a mixed range falls back."""
    fused, _ = fctx
    keep = []
    comps, _ = J.make_components(64, 64, [(2, 1), (1, 1), (1, 1)])
    q = [np.ones(64, np.uint16)] * 3
    d1 = J.make_image_desc(64, 64, comps, q, [np.zeros(c.block_w * c.block_h * 64, np.int16) for c in comps], J.CT_YCBCR, keep)
    comps2, _ = J.make_components(64, 64, [(1, 1)] * 3)
    d2 = J.make_image_desc(64, 64, comps2, q, [np.zeros(c.block_w * c.block_h * 64, np.int16) for c in comps2], J.CT_RGB, keep)
    comps3, _ = J.make_components(64, 64, [(1, 1)] * 3, dct_scale=4)
    d3 = J.make_image_desc(64, 64, comps3, q, [np.zeros(c.block_w * c.block_h * 64, np.int16) for c in comps3], J.CT_YCBCR, keep)
    d4 = J.make_image_desc(64, 64, comps2, q, [np.zeros(c.block_w * c.block_h * 64, np.int16) for c in comps2], J.CT_YCBCR, keep)
    b = J.Batch(fused, [d1, d2, d3])
    assert b.info.n_fused == 0
    b.close()
    b = J.Batch(fused, [d1, d4])
    assert b.info.n_fused == 1
    outs = [np.zeros(64 * 64 * 3, np.uint8) for _ in range(2)]
    assert b.run_host(outs) == [0, 0]
    assert outs[1].tolist() == [128] * (64 * 64 * 3)   # all-zero coefficients: every sample 128 -> grey
    b.close()
