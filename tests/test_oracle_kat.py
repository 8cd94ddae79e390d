"""Pins the oracle (and the constants the CUDA kernels hard-code) on the reference's own known-answer
tests: src/idct.rs:580-657 (three IDCT KATs), src/parser.rs:312-329 (geometry), src/idct.rs:30-203
(choose_idct_size).  CPU only."""
import re

import numpy as np
import pytest

COEFS = [-14, -39, 58, -2, 3, 3, 0, 1, 11, 27, 4, -3, 3, 0, 1, 0, -6, -13, -9, -1, -2, -1, 0, 0, -4, 0, -1, -2, 0, 0, 0, 0,
         3, 0, 0, 0, 0, 0, 0, 0, -3, -2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
QT = [8, 6, 5, 8, 12, 20, 26, 31, 6, 6, 7, 10, 13, 29, 30, 28, 7, 7, 8, 12, 20, 29, 35, 28, 7, 9, 11, 15, 26, 44, 40, 31,
      9, 11, 19, 28, 34, 55, 52, 39, 12, 18, 28, 32, 41, 52, 57, 46, 25, 32, 39, 44, 52, 61, 60, 51, 36, 46, 48, 49, 56, 50, 52, 50]
EXPECTED = [118, 92, 110, 83, 77, 93, 144, 198, 172, 116, 114, 87, 78, 93, 146, 191, 194, 107, 91, 76, 71, 93, 160, 198,
            196, 100, 80, 74, 67, 92, 174, 209, 182, 104, 88, 81, 68, 89, 178, 206, 105, 64, 59, 59, 63, 94, 183, 201,
            35, 27, 28, 37, 72, 121, 203, 204, 37, 45, 41, 47, 98, 154, 223, 208]
SATURATED = [0, 0, 0, 255, 255, 0, 0, 255, 0, 0, 215, 0, 0, 255, 255, 0, 255, 255, 255, 255, 255, 0, 0, 255, 0, 0, 255, 0,
             255, 0, 255, 255, 0, 0, 255, 255, 0, 255, 0, 0, 255, 255, 0, 255, 255, 255, 170, 0, 0, 255, 0, 0, 0, 0, 0, 255,
             255, 255, 0, 255, 0, 255, 0, 0]


def test_idct_kat_real_block(oracle_mod):
    """src/idct.rs:580-627: tolerance +-1 in the reference; the scalar restatement is exact."""
    out = oracle_mod.idct_block(COEFS, QT).reshape(-1)
    assert out.tolist() == EXPECTED
    out_s = oracle_mod.idct_block(COEFS, QT, arith=oracle_mod.ARITH_SSSE3).reshape(-1).astype(int)
    assert np.abs(out_s - np.array(EXPECTED)).max() <= 1


def test_idct_kat_all_zero(oracle_mod):
    """src/idct.rs:629-634"""
    for arith in (0, 1):
        assert oracle_mod.idct_block([0] * 64, [666] * 64, arith=arith).reshape(-1).tolist() == [128] * 64


def test_idct_kat_saturated(oracle_mod):
    """src/idct.rs:636-657: pins i32 wrapping of the scalar variant"""
    assert oracle_mod.idct_block([32767] * 64, [65535] * 64).reshape(-1).tolist() == SATURATED


def test_ssse3_emulation_matches_intrinsics(oracle_mod):
    """The portable 16-bit emulation of src/arch/ssse3.rs equals real <tmmintrin.h> code on random blocks."""
    rng = np.random.default_rng(1)
    n = 0
    for _ in range(3000):
        c = rng.integers(-1024, 1024, 64).astype(np.int16)
        if rng.random() < 0.3:
            c = rng.integers(-32768, 32768, 64).astype(np.int16)
        q = rng.integers(1, 256 if rng.random() < 0.8 else 65536, 64).astype(np.uint16)
        a = oracle_mod.idct_block(c, q, arith=oracle_mod.ARITH_SSSE3)
        b = oracle_mod.idct_block_ssse3_intrinsics(c, q)
        if b is None:
            return
        assert np.array_equal(a, b)
        n += 1
    assert n == 3000


def test_ssse3_colour_emulation_matches_intrinsics(oracle_mod):
    import ctypes as C
    rng = np.random.default_rng(2)
    for w in (8, 15, 16, 17, 64, 100):
        y, cb, cr = (rng.integers(0, 256, w).astype(np.uint8) for _ in range(3))
        out = np.zeros(3 * w, dtype=np.uint8)
        done = C.c_size_t()
        ok = oracle_mod.lib().orc_ycbcr_line_ssse3_intrin(y.ctypes.data, cb.ctypes.data, cr.ctypes.data, out.ctypes.data, w, C.byref(done))
        if not ok:
            return
        assert done.value == max(w // 8 - 1, 0) * 8
        comps, _ = oracle_mod.make_components(w, 1, [(1, 1)] * 3)
        img = oracle_mod.compute_image(comps, [np.pad(p, (0, comps[0].block_w * 8 * 8 - w)) for p in (y, cb, cr)], w, 1,
                                       oracle_mod.CT_YCBCR, arith=oracle_mod.ARITH_SSSE3)
        assert np.array_equal(img[:3 * done.value], out[:3 * done.value])


def test_geometry_kat(oracle_mod, J):
    """src/parser.rs:312-329"""
    for mod in (oracle_mod, J):
        comps, mcu = mod.make_components(800, 280, [(2, 2)])
        assert mcu == (50, 18)
        assert (comps[0].block_w, comps[0].block_h) == (100, 36)
        assert (comps[0].size_w, comps[0].size_h) == (800, 280)


def test_choose_idct_size(oracle_mod, J):
    """src/idct.rs:30-203 (selected rows of the table)"""
    cases = [((5472, 3648), (200, 200), 1), ((5472, 3648), (500, 500), 1), ((5472, 3648), (684, 456), 1),
             ((5472, 3648), (999, 456), 1), ((5472, 3648), (684, 999), 1), ((5472, 3648), (500, 333), 1),
             ((5472, 3648), (685, 999), 2), ((5472, 3648), (1000, 1000), 2), ((5472, 3648), (1400, 1400), 4),
             ((5472, 3648), (5472, 3648), 8), ((5472, 3648), (16384, 16384), 8), ((1, 1), (65535, 65535), 8),
             ((5472, 3648), (16384, 16384), 8)]
    for full, req, want in cases:
        assert oracle_mod.lib().orc_choose_idct_size(full[0], full[1], req[0], req[1]) == want
        assert J.lib().b200jpg_choose_idct_size(full[0], full[1], req[0], req[1]) == want


def test_kernel_constants_match_f32_evaluation():
    """The integer constants hard-coded in the .cu files equal the reference's f32 expressions
    stbi_f2f(x) = (x * 4096 + 0.5) as i32 (src/idct.rs:572-574) and (x * 2^20 + 0.5) as i32 (src/decoder.rs:1502-1504)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    k1 = open(os.path.join(root, "jpeg_decoder_b200", "csrc", "idct_core.cuh")).read()   # shared by K1 and the fused kernel
    k2 = open(os.path.join(root, "jpeg_decoder_b200", "csrc", "color_core.cuh")).read()  # shared by K2 and the fused kernel

    def f2f(x, bits):
        return int(np.float32(x) * np.float32(2 ** bits) + np.float32(0.5))

    want = {"F2F_0_5411961": f2f(0.5411961, 12), "F2F_N1_847759065": f2f(-1.847759065, 12), "F2F_0_765366865": f2f(0.765366865, 12),
            "F2F_1_175875602": f2f(1.175875602, 12), "F2F_0_298631336": f2f(0.298631336, 12), "F2F_2_053119869": f2f(2.053119869, 12),
            "F2F_3_072711026": f2f(3.072711026, 12), "F2F_1_501321110": f2f(1.501321110, 12), "F2F_N0_899976223": f2f(-0.899976223, 12),
            "F2F_N2_562915447": f2f(-2.562915447, 12), "F2F_N1_961570560": f2f(-1.961570560, 12), "F2F_N0_390180644": f2f(-0.390180644, 12)}
    for name, v in want.items():
        m = re.search(r"#define %s \(?\(?(?:unsigned\))?(-?\d+)u?\)?" % name, k1)
        assert m, name
        assert int(m.group(1)) == v, (name, m.group(1), v)
    want2 = {"C_R_CR": f2f(1.40200, 20), "C_G_CB": f2f(0.34414, 20), "C_G_CR": f2f(0.71414, 20), "C_B_CB": f2f(1.77200, 20)}
    for name, v in want2.items():
        m = re.search(r"#define %s (-?\d+)" % name, k2)
        assert m and int(m.group(1)) == v, (name, v)


def test_direct_form_pass_is_the_butterfly_mod_2_32():
    """idct_core.cuh evaluates the 1-D pass of src/idct.rs:378-447 in "direct form" (IDCT_1D_DIRECT: the odd half as a
    4x4 constant matrix seeded with the even half).  Everything between two shifts is Wrapping<i32>, so the regrouping
    must be an identity over Z/2^32: checked here on the CPU with exact integers, extremes included."""
    M = 1 << 32
    A, B, Cc, D = 2217, -7567, 3135, 4816
    E1, E2, E3, E4 = -3685, -10497, -8034, -1597
    G1, G3, G5, G7 = 6149, 12586, 8410, 1223

    def butterfly(s, xs):   # the reference's order of operations
        s0, s1, s2, s3, s4, s5, s6, s7 = s
        p1 = (s2 + s6) * A
        e2, e3 = p1 + s6 * B, p1 + s2 * Cc
        e0, e1 = ((s0 + s4) << 12) + xs, ((s0 - s4) << 12) + xs
        x0, x3, x1, x2 = e0 + e3, e0 - e3, e1 + e2, e1 - e2
        q3, q4, q1, q2 = s7 + s3, s5 + s1, s7 + s1, s5 + s3
        p5 = (q3 + q4) * D
        q1, q2, q3, q4 = p5 + q1 * E1, p5 + q2 * E2, q3 * E3, q4 * E4
        t3, t2, t1, t0 = s1 * G1 + q1 + q4, s3 * G3 + q2 + q3, s5 * G5 + q2 + q4, s7 * G7 + q1 + q3
        return [v % M for v in (x0 + t3, x1 + t2, x2 + t1, x3 + t0, x3 - t0, x2 - t1, x1 - t2, x0 - t3)]

    def direct(s, xs):      # IDCT_1D_DIRECT, every intermediate reduced mod 2^32 like the GPU's registers
        s0, s1, s2, s3, s4, s5, s6, s7 = s
        e3 = (s2 * ((A + Cc) % M) + s6 * (A % M)) % M
        e2 = (s2 * (A % M) + s6 * ((A + B) % M)) % M
        e0, e1 = ((((s0 + s4) % M) << 12) + xs) % M, ((((s0 - s4) % M) << 12) + xs) % M
        x0, x3, x1, x2 = (e0 + e3) % M, (e0 - e3) % M, (e1 + e2) % M, (e1 - e2) % M
        o0 = (x0 + s1 * ((G1 + D + E1 + E4) % M) + s3 * (D % M) + s5 * ((D + E4) % M) + s7 * ((D + E1) % M)) % M
        o1 = (x1 + s1 * (D % M) + s3 * ((G3 + D + E2 + E3) % M) + s5 * ((D + E2) % M) + s7 * ((D + E3) % M)) % M
        o2 = (x2 + s1 * ((D + E4) % M) + s3 * ((D + E2) % M) + s5 * ((G5 + D + E2 + E4) % M) + s7 * (D % M)) % M
        o3 = (x3 + s1 * ((D + E1) % M) + s3 * ((D + E3) % M) + s5 * (D % M) + s7 * ((G7 + D + E1 + E3) % M)) % M
        return [o0, o1, o2, o3, (2 * x3 - o3) % M, (2 * x2 - o2) % M, (2 * x1 - o1) % M, (2 * x0 - o0) % M]

    rng = np.random.default_rng(42)
    cases = [list(map(int, rng.integers(0, M, 8))) for _ in range(3000)]
    cases += [[int(v) % M for v in rng.choice([0, 1, -1, 2 ** 31 - 1, -2 ** 31, 32767 * 65535, -32768 * 65535], 8)] for _ in range(500)]
    for s in cases:
        for xs in (512 % M, (512 + 2 ** 31) % M, (65536 + (128 << 17)) % M):
            assert butterfly(s, xs) == direct(s, xs)


def test_ssse3_colour_path_never_saturates_for_8bit_inputs(oracle_mod):
    """color_core.cuh evaluates the SSSE3 colour path (src/arch/ssse3.rs:208-244) with plain adds: for 8-bit inputs no
    saturating add of that path can saturate.  Exhaustive over all 2^24 (y, cb, cr): the saturating form (the oracle's
    emulation, restated with numpy) and the plain form give the same bytes, and both equal the C oracle on a sample."""
    y = np.arange(256, dtype=np.int32)[:, None, None]
    cb = np.arange(256, dtype=np.int32)[None, :, None]
    cr = np.arange(256, dtype=np.int32)[None, None, :]

    def sat(v):
        return np.clip(v, -32768, 32767)

    def mulhrs(a, c):
        return (((a * c) >> 14) + 1) >> 1

    y6 = sat((y << 6) + 32)
    cb6, cr6 = sat((cb << 6) - 8192), sat((cr << 6) - 8192)
    r_s = sat(y6 + sat(mulhrs(cr6, 13173) + cr6)) >> 6
    g_s = sat(y6 - sat(mulhrs(cb6, 11276) + mulhrs(cr6, 23401))) >> 6
    b_s = sat(y6 + sat(mulhrs(cb6, 25297) + cb6)) >> 6
    cbm, crm = (cb - 128) * 64, (cr - 128) * 64
    yy = y * 64 + 32
    r_p = (yy + ((crm * 13173 + 16384) >> 15) + crm) >> 6
    g_p = (yy - (((cbm * 11276 + 16384) >> 15) + ((crm * 23401 + 16384) >> 15))) >> 6
    b_p = (yy + ((cbm * 25297 + 16384) >> 15) + cbm) >> 6
    assert np.array_equal(np.broadcast_to(r_s, (256, 256, 256)), np.broadcast_to(r_p, (256, 256, 256)))
    assert np.array_equal(g_s, g_p)
    assert np.array_equal(np.broadcast_to(b_s, (256, 256, 256)), np.broadcast_to(b_p, (256, 256, 256)))
    # and the numpy restatement is the oracle's: one row of 40 pixels through the real intrinsics
    import ctypes as C
    rng = np.random.default_rng(9)
    n = 40
    yy8, cb8, cr8 = (rng.integers(0, 256, n).astype(np.uint8) for _ in range(3))
    out = np.zeros(3 * n, np.uint8)
    done = C.c_size_t()
    if oracle_mod.lib().orc_ycbcr_line_ssse3_intrin(yy8.ctypes.data, cb8.ctypes.data, cr8.ctypes.data, out.ctypes.data, n, C.byref(done)):
        k = done.value
        assert k == 32
        want = np.stack([np.clip(r_s[yy8[:k], 0, cr8[:k]], 0, 255), np.clip(g_s[yy8[:k], cb8[:k], cr8[:k]], 0, 255),
                         np.clip(b_s[yy8[:k], cb8[:k], 0], 0, 255)], -1).astype(np.uint8)
        assert np.array_equal(out[:3 * k].reshape(k, 3), want)


@pytest.mark.parametrize("name", ["h2v2", "h2v1", "h1v2"])
def test_hand_derived_upsampling_vectors(oracle_mod, name):
    """The oracle against bytes worked out by hand from src/upsampler.rs (tests/hand_vectors.py)."""
    from hand_vectors import planes_for
    comps, planes, w, h, want = planes_for(oracle_mod.make_components, name, np.random.default_rng(1))
    for arith in (oracle_mod.ARITH_SCALAR, oracle_mod.ARITH_SSSE3):
        got = oracle_mod.compute_image(comps, planes, w, h, oracle_mod.CT_RGB, arith=arith).reshape(h, w, 3)
        assert np.array_equal(got[..., 1], want), (name, got[..., 1])
