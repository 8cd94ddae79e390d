"""CPU oracle for the JPEG block pipeline -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

ctypes binding of oracle/liboracle.so, a plain-C restatement of image-rs/jpeg-decoder v0.3.2
(/root/reference/src/{idct,upsampler,decoder,parser,huffman,marker}.rs, src/arch/ssse3.rs,
src/worker/immediate.rs).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this package; the product (jpeg_decoder_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liboracle.so")

ARITH_SCALAR, ARITH_SSSE3, ARITH_SSSE3_NATIVE = 0, 1, 2
CT_NONE, CT_UNKNOWN, CT_GRAYSCALE, CT_RGB, CT_YCBCR, CT_CMYK, CT_YCCK, CT_JCS_BG_YCC, CT_JCS_BG_RGB = range(9)
OK, ERR_FORMAT, ERR_UNSUPPORTED, ERR_IO, ERR_INTERNAL = range(5)


def build(force=False):
    """Compile oracle/*.c -> oracle/liboracle.so (gcc, a few seconds)."""
    srcs = [os.path.join(_DIR, f) for f in ("ref_idct.c", "ref_image.c", "ref_decoder.c", "oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-s", "-C", _DIR, "-B", "liboracle.so"])
    return _SO


class Component(C.Structure):
    """src/parser.rs:77-89"""
    _fields_ = [("identifier", C.c_uint8), ("h", C.c_uint8), ("v", C.c_uint8), ("tq", C.c_uint8),
                ("dct_scale", C.c_uint16), ("size_w", C.c_uint16), ("size_h", C.c_uint16),
                ("block_w", C.c_uint16), ("block_h", C.c_uint16)]

    def __repr__(self):
        return ("Component(id=%d, h=%d, v=%d, tq=%d, dct_scale=%d, size=%dx%d, block_size=%dx%d)" % (
            self.identifier, self.h, self.v, self.tq, self.dct_scale, self.size_w, self.size_h,
            self.block_w, self.block_h))


class ImageInfo(C.Structure):
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("pixel_format", C.c_int), ("coding_process", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, i16p, u16p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_uint16)
    L.orc_idct_block.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_idct_block.restype = None
    L.orc_idct8x8_ssse3_intrin.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_ycbcr_line_ssse3_intrin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.orc_choose_idct_size.argtypes = [C.c_uint16] * 4
    L.orc_update_component_sizes.argtypes = [C.c_uint16, C.c_uint16, C.POINTER(Component), C.c_int, u16p, u16p]
    L.orc_worker_new.restype = C.c_void_p
    L.orc_worker_new.argtypes = [C.c_int]
    L.orc_worker_free.argtypes = [C.c_void_p]
    L.orc_worker_free.restype = None
    L.orc_worker_start.argtypes = [C.c_void_p, C.c_int, C.POINTER(Component), C.c_void_p]
    L.orc_worker_append_row.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.orc_worker_get_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_compute_image.argtypes = [C.c_int, C.POINTER(Component), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                    C.c_uint16, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.orc_last_error.restype = C.c_char_p
    L.orc_hotpath_image.argtypes = [C.c_int, C.POINTER(Component), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.c_uint16, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t]
    L.orc_hotpath_batch.argtypes = [C.c_int, C.c_int, C.c_size_t, C.POINTER(Component), C.c_int, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.c_uint16, C.c_uint16, C.c_int, C.POINTER(C.c_void_p), C.c_size_t]
    L.orc_hotpath_image_mt.argtypes = [C.c_int, C.c_int, C.POINTER(Component), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.c_uint16, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t]
    L.orc_decoder_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.orc_decoder_set_threads.restype = None
    L.orc_decoder_set_taps.argtypes = [C.c_void_p, C.c_int]
    L.orc_decoder_set_taps.restype = None
    L.orc_decoder_new.restype = C.c_void_p
    L.orc_decoder_new.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    L.orc_decoder_free.argtypes = [C.c_void_p]
    L.orc_decoder_free.restype = None
    L.orc_decoder_set_color_transform.argtypes = [C.c_void_p, C.c_int]
    L.orc_decoder_set_color_transform.restype = None
    L.orc_decoder_set_max_decoding_buffer_size.argtypes = [C.c_void_p, C.c_size_t]
    L.orc_decoder_set_max_decoding_buffer_size.restype = None
    L.orc_decoder_read_info.argtypes = [C.c_void_p]
    L.orc_decoder_info.argtypes = [C.c_void_p, C.POINTER(ImageInfo)]
    L.orc_decoder_scale.argtypes = [C.c_void_p, C.c_uint16, C.c_uint16, u16p, u16p]
    L.orc_decoder_decode.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_decoder_error.argtypes = [C.c_void_p]
    L.orc_decoder_error.restype = C.c_char_p
    L.orc_decoder_color_transform.argtypes = [C.c_void_p]
    L.orc_decoder_ncomp.argtypes = [C.c_void_p]
    L.orc_decoder_component.argtypes = [C.c_void_p, C.c_int, C.POINTER(Component), C.c_void_p]
    for name in ("orc_decoder_coefficients", "orc_decoder_plane"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    for name in ("orc_decoder_icc_profile", "orc_decoder_exif", "orc_decoder_xmp"):
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    _lib = L
    return L


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__("oracle error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def idct_block(coefs, qt, scale=8, arith=ARITH_SCALAR):
    """dequantize_and_idct_block (src/idct.rs:205-239) -> (scale, scale) uint8."""
    c = np.ascontiguousarray(coefs, dtype=np.int16).reshape(64)
    q = np.ascontiguousarray(qt, dtype=np.uint16).reshape(64)
    out = np.zeros((8, 8), dtype=np.uint8)
    lib().orc_idct_block(arith, scale, _ptr(c), _ptr(q), 8, _ptr(out))
    return out[:scale, :scale].copy()


def idct_block_ssse3_intrinsics(coefs, qt):
    c = np.ascontiguousarray(coefs, dtype=np.int16).reshape(64)
    q = np.ascontiguousarray(qt, dtype=np.uint16).reshape(64)
    out = np.zeros((8, 8), dtype=np.uint8)
    ok = lib().orc_idct8x8_ssse3_intrin(_ptr(c), _ptr(q), 8, _ptr(out))
    return out if ok else None


def make_components(width, height, sampling, dct_scale=8, tqs=None, ids=None):
    """Build Component[] with update_component_sizes (src/parser.rs:292-310).
    sampling: [(H, V), ...].  Returns (ctypes array, (mcu_w, mcu_h))."""
    n = len(sampling)
    arr = (Component * n)()
    for i, (h, v) in enumerate(sampling):
        arr[i].identifier = (ids[i] if ids else i + 1)
        arr[i].h, arr[i].v = h, v
        arr[i].tq = (tqs[i] if tqs else (0 if i == 0 else 1))
        arr[i].dct_scale = dct_scale
    mw, mh = C.c_uint16(), C.c_uint16()
    rc = lib().orc_update_component_sizes(width, height, arr, n, C.byref(mw), C.byref(mh))
    if rc:
        raise OracleError(rc, "invalid dimensions")
    return arr, (mw.value, mh.value)


class Worker:
    """trait Worker (src/worker/mod.rs:24-35) backed by the immediate worker restatement."""

    def __init__(self, arith=ARITH_SCALAR):
        self._w = lib().orc_worker_new(arith)

    def __del__(self):
        if getattr(self, "_w", None):
            lib().orc_worker_free(self._w)
            self._w = None

    def start(self, index, component, qt):
        q = np.ascontiguousarray(qt, dtype=np.uint16).reshape(64)
        rc = lib().orc_worker_start(self._w, index, C.byref(component), _ptr(q))
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())

    def append_row(self, index, coefs):
        c = np.ascontiguousarray(coefs, dtype=np.int16).reshape(-1)
        rc = lib().orc_worker_append_row(self._w, index, _ptr(c), c.size)
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())

    def get_result(self, index):
        p, n = C.c_void_p(), C.c_size_t()
        rc = lib().orc_worker_get_result(self._w, index, C.byref(p), C.byref(n))
        if rc:
            raise OracleError(rc, lib().orc_last_error().decode())
        if not p.value:
            return np.zeros(0, dtype=np.uint8)
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()
        C.CDLL(None).free(p)
        return out


def idct_planes(components, qts, coefs, arith=ARITH_SCALAR):
    """start + append_row per MCU row + get_result for each component -> list of planes."""
    w = Worker(arith)
    planes = []
    for i, comp in enumerate(components):
        w.start(i, comp, qts[i])
        per_row = comp.block_w * comp.v * 64
        c = np.ascontiguousarray(coefs[i], dtype=np.int16).reshape(-1)
        assert c.size % per_row == 0
        for r in range(c.size // per_row):
            w.append_row(i, c[r * per_row:(r + 1) * per_row])
        planes.append(w.get_result(i))
    return planes


def compute_image(components, planes, out_w, out_h, color_transform, arith=ARITH_SCALAR):
    """compute_image (src/decoder.rs:1300-1336)."""
    n = len(planes)
    ps = [np.ascontiguousarray(p, dtype=np.uint8).reshape(-1) for p in planes]
    pp = (C.c_void_p * n)(*[p.ctypes.data for p in ps])
    pl = (C.c_size_t * n)(*[p.size for p in ps])
    cap = int(out_w) * int(out_h) * max(n, 1)
    out = np.zeros(cap, dtype=np.uint8)
    ol = C.c_size_t()
    rc = lib().orc_compute_image(arith, components, n, pp, pl, out_w, out_h, color_transform, _ptr(out), cap, C.byref(ol))
    if rc:
        raise OracleError(rc, lib().orc_last_error().decode())
    return out[:ol.value]


def hotpath_image(components, qts, coefs, out_w, out_h, color_transform, arith=ARITH_SCALAR, nthreads=1):
    n = len(components)
    qs = [np.ascontiguousarray(q, dtype=np.uint16).reshape(64) for q in qts]
    cs = [np.ascontiguousarray(c, dtype=np.int16).reshape(-1) for c in coefs]
    qp = (C.c_void_p * 4)(*([q.ctypes.data for q in qs] + [None] * (4 - n)))
    cp = (C.c_void_p * 4)(*([c.ctypes.data for c in cs] + [None] * (4 - n)))
    cap = int(out_w) * int(out_h) * n
    out = np.zeros(cap, dtype=np.uint8)
    rc = lib().orc_hotpath_image_mt(arith, nthreads, components, n, qp, cp, out_w, out_h, color_transform, _ptr(out), cap)
    if rc:
        raise OracleError(rc, lib().orc_last_error().decode())
    return out if n > 1 else out[:components[0].size_w * components[0].size_h]


def hotpath_batch(components, qts, coefs_per_image, out_w, out_h, color_transform, outs, nthreads, arith=ARITH_SCALAR):
    """Timed CPU baseline: n images of identical geometry over `nthreads` threads.
    coefs_per_image: list (len n) of lists (len ncomp) of int16 arrays; outs: list of uint8 arrays."""
    n = len(coefs_per_image)
    nc = len(components)
    qs = [np.ascontiguousarray(q, dtype=np.uint16).reshape(64) for q in qts]
    qp = (C.c_void_p * 4)(*([q.ctypes.data for q in qs] + [None] * (4 - nc)))
    cp = (C.c_void_p * (n * nc))(*[c.ctypes.data for img in coefs_per_image for c in img])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    rc = lib().orc_hotpath_batch(arith, nthreads, n, components, nc, qp, cp, out_w, out_h, color_transform, op, outs[0].size)
    if rc:
        raise OracleError(rc, lib().orc_last_error().decode())


class Decoder:
    """Decoder<R> (src/decoder.rs:101-295) over an in-memory file."""

    def __init__(self, data, arith=ARITH_SCALAR):
        self._buf = np.frombuffer(bytes(data), dtype=np.uint8)
        self._d = lib().orc_decoder_new(_ptr(self._buf), self._buf.size, arith)

    def __del__(self):
        if getattr(self, "_d", None):
            lib().orc_decoder_free(self._d)
            self._d = None

    def _check(self, rc):
        if rc:
            raise OracleError(rc, lib().orc_decoder_error(self._d).decode(errors="replace"))

    def set_color_transform(self, ct):
        lib().orc_decoder_set_color_transform(self._d, ct)

    def set_threads(self, n):
        """compute_image rows over n threads (the rayon build's colour stage); default 1"""
        lib().orc_decoder_set_threads(self._d, n)

    def set_taps(self, on):
        lib().orc_decoder_set_taps(self._d, int(bool(on)))

    def set_max_decoding_buffer_size(self, n):
        lib().orc_decoder_set_max_decoding_buffer_size(self._d, n)

    def read_info(self):
        self._check(lib().orc_decoder_read_info(self._d))

    def info(self):
        inf = ImageInfo()
        return inf if lib().orc_decoder_info(self._d, C.byref(inf)) else None

    def scale(self, w, h):
        ow, oh = C.c_uint16(), C.c_uint16()
        self._check(lib().orc_decoder_scale(self._d, w, h, C.byref(ow), C.byref(oh)))
        return ow.value, oh.value

    def decode(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(lib().orc_decoder_decode(self._d, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()

    def color_transform(self):
        return lib().orc_decoder_color_transform(self._d)

    def components(self):
        n = lib().orc_decoder_ncomp(self._d)
        comps = (Component * n)()
        qts = []
        for i in range(n):
            q = np.zeros(64, dtype=np.uint16)
            lib().orc_decoder_component(self._d, i, C.byref(comps[i]), _ptr(q))
            qts.append(q)
        return comps, qts

    def _blob(self, fn, *args):
        p, n = C.c_void_p(), C.c_size_t()
        if not fn(self._d, *args, C.byref(p), C.byref(n)):
            return None
        return p, n.value

    def coefficients(self, i):
        r = self._blob(lib().orc_decoder_coefficients, i)
        if r is None:
            return None
        return np.ctypeslib.as_array(C.cast(r[0], C.POINTER(C.c_int16)), shape=(r[1],)).copy()

    def plane(self, i):
        r = self._blob(lib().orc_decoder_plane, i)
        if r is None:
            return None
        return np.ctypeslib.as_array(C.cast(r[0], C.POINTER(C.c_uint8)), shape=(r[1],)).copy()

    def _bytes(self, fn):
        r = self._blob(fn)
        if r is None:
            return None
        return C.string_at(r[0], r[1])

    def icc_profile(self):
        return self._bytes(lib().orc_decoder_icc_profile)

    def exif_data(self):
        return self._bytes(lib().orc_decoder_exif)

    def xmp_data(self):
        return self._bytes(lib().orc_decoder_xmp)
