"""jpeg_decoder_b200 -- B200-native JPEG block pipeline behind the API of image-rs/jpeg-decoder.

Host-side mirror of the reference's interfaces for the hot path, over the C ABI of
include/b200jpg.h (see INTEGRATION.md for the Rust binding a maintainer would add):

  Decoder        jpeg_decoder::Decoder      (reference src/decoder.rs:101-295)
  Worker         trait Worker               (src/worker/mod.rs:24-35)
  compute_image  decoder::compute_image     (src/decoder.rs:1300-1336)
  decode_batch   new surface: n independent images, dense coefficients -> pixels

The reference is Rust and no Rust toolchain exists in this image, so the compiled host side is
C++ (csrc/host_decoder.cpp, csrc/pipeline.cu) and this module is a thin ctypes layer used by the
tests and bench.py.  Nothing here imports oracle/.
"""
import ctypes as C

import numpy as np

from ._native import (ARITH_SCALAR, ARITH_SSSE3, FUSE_AUTO, FUSE_OFF, FUSE_ON, FMT_RGB8_PLANAR, FMT_RGB_F32_NHWC, FMT_RGB_F32_NCHW, ENTROPY_AUTO, ENTROPY_DEVICE, ENTROPY_HOST, COMPACT_AUTO, COMPACT_OFF, COMPACT_ON, CP_DCT_PROGRESSIVE, CP_DCT_SEQUENTIAL, CP_LOSSLESS, CT_CMYK,
                      CT_GRAYSCALE, CT_JCS_BG_RGB, CT_JCS_BG_YCC, CT_NONE, CT_RGB, CT_UNKNOWN, CT_YCBCR, CT_YCCK,
                      ERR_FORMAT, ERR_INTERNAL, ERR_IO, ERR_UNSUPPORTED, KERNEL_AUTO, KERNEL_FAST, KERNEL_GENERIC, OK,
                      PF_CMYK32, PF_L8, PF_L16, PF_RGB24, SBS_INTERLEAVED, SBS_NATURAL, SBS_PLANAR, BatchInfo, Component, FileJob, ImageDesc,
                      ImageInfo, Options, SbsStream, lib)

__all__ = ["Context", "Worker", "Batch", "Decoder", "B200JpgError", "FileJob", "make_components", "make_image_desc",
           "compute_image", "decode_batch", "decode_batch_sbs", "expand_sbs", "sbs_from_dense", "decode_files", "decode_files_into", "read_info_files", "Component",
           "ImageDesc", "SbsStream"]


class B200JpgError(Exception):
    """Error::{Format, Unsupported, Io, Internal}, reference src/error.rs:37-48"""

    def __init__(self, code, msg):
        kind = {ERR_FORMAT: "Format", ERR_UNSUPPORTED: "Unsupported", ERR_IO: "Io", ERR_INTERNAL: "Internal"}.get(code, "?")
        super().__init__("%s(%d): %s" % (kind, code, msg))
        self.code = code
        self.msg = msg


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_components(width, height, sampling, dct_scale=8, tqs=None, ids=None):
    """Component[] with update_component_sizes (src/parser.rs:292-310). sampling = [(H, V), ...]."""
    n = len(sampling)
    arr = (Component * n)()
    for i, (h, v) in enumerate(sampling):
        arr[i].identifier = ids[i] if ids else i + 1
        arr[i].h, arr[i].v = h, v
        arr[i].tq = tqs[i] if tqs else (0 if i == 0 else 1)
        arr[i].dct_scale = dct_scale
    mw, mh = C.c_uint16(), C.c_uint16()
    rc = lib().b200jpg_update_component_sizes(width, height, arr, n, C.byref(mw), C.byref(mh))
    if rc:
        raise B200JpgError(rc, "invalid dimensions")
    return arr, (mw.value, mh.value)


def make_image_desc(width, height, components, qts, coefs, color_transform, keep):
    """ImageDesc over numpy arrays; `keep` (a list) receives references that must outlive the desc."""
    d = ImageDesc()
    d.width, d.height = width, height
    d.ncomp = len(components)
    d.color_transform = color_transform
    for i, c in enumerate(components):
        d.comps[i] = c
        q = np.ascontiguousarray(qts[i], dtype=np.uint16).reshape(64)
        keep.append(q)
        d.qt[i] = q.ctypes.data
        if coefs is not None and coefs[i] is not None:
            a = coefs[i]
            if not (isinstance(a, np.ndarray) and a.dtype == np.int16 and a.flags["C_CONTIGUOUS"]):
                a = np.ascontiguousarray(a, dtype=np.int16)
            keep.append(a)
            d.coefs[i] = a.ctypes.data
    return d


class Context:
    """b200jpg_ctx: one per device/stream."""

    def __init__(self, device=0, arith=ARITH_SCALAR, k1_kernel=KERNEL_AUTO, k2_kernel=KERNEL_AUTO, stream=None,
                 host_compact=COMPACT_AUTO, host_threads=0, entropy=0, fuse=0):
        opt = Options()
        lib().b200jpg_default_options(C.byref(opt))
        opt.device, opt.arith, opt.k1_kernel, opt.k2_kernel = device, arith, k1_kernel, k2_kernel
        opt.stream = stream
        opt.host_compact, opt.host_threads = host_compact, host_threads
        opt.entropy = entropy  # ENTROPY_AUTO / ENTROPY_HOST / ENTROPY_DEVICE
        opt.fuse = fuse        # FUSE_AUTO / FUSE_OFF / FUSE_ON
        h = C.c_void_p()
        rc = lib().b200jpg_create(C.byref(opt), C.byref(h))
        if rc:
            raise B200JpgError(rc, "b200jpg_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self._h = h
        self.arith = arith
        self.fuse = fuse

    def close(self):
        if getattr(self, "_h", None):
            lib().b200jpg_destroy(self._h)
            self._h = None

    __del__ = close

    def check(self, rc):
        if rc:
            raise B200JpgError(rc, lib().b200jpg_last_error(self._h).decode(errors="replace"))

    def synchronize(self):
        self.check(lib().b200jpg_synchronize(self._h))

    def set_fuse(self, fuse):
        """FUSE_AUTO: the fused kernel where it is the faster route (4:4:4); FUSE_ON: wherever it applies; FUSE_OFF: never."""
        lib().b200jpg_set_fuse(self._h, fuse)

    @property
    def launch_count(self):
        return lib().b200jpg_launch_count(self._h)

    @property
    def device_scan_counts(self):
        """(scans Huffman-decoded on the GPU, scans the GPU handed back to the host) by decode_files so far."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        lib().b200jpg_device_scan_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value


class Worker:
    """trait Worker { start, append_row, append_rows, get_result } (src/worker/mod.rs:24-35)."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(lib().b200jpg_worker_new(ctx._h, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().b200jpg_worker_free(self._h)
            self._h = None

    __del__ = close

    def start(self, index, component, qt):
        q = np.ascontiguousarray(qt, dtype=np.uint16).reshape(64)
        self.ctx.check(lib().b200jpg_worker_start(self._h, index, C.byref(component), _ptr(q)))

    def append_row(self, index, coefs):
        c = np.ascontiguousarray(coefs, dtype=np.int16).reshape(-1)
        self.ctx.check(lib().b200jpg_worker_append_row(self._h, index, _ptr(c), c.size))

    def append_rows(self, index, coefs, nrows):
        c = np.ascontiguousarray(coefs, dtype=np.int16).reshape(-1)
        self.ctx.check(lib().b200jpg_worker_append_rows(self._h, index, _ptr(c), c.size, nrows))

    def get_result(self, index):
        n = C.c_size_t()
        self.ctx.check(lib().b200jpg_worker_get_result(self._h, index, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self.ctx.check(lib().b200jpg_worker_get_result(self._h, index, _ptr(out), out.size, C.byref(n)))
        return out

    def compute_image(self, ncomp, out_w, out_h, color_transform):
        out = np.zeros(int(out_w) * int(out_h) * ncomp, dtype=np.uint8)
        n = C.c_size_t()
        self.ctx.check(lib().b200jpg_worker_compute_image(self._h, ncomp, out_w, out_h, color_transform, _ptr(out), out.size, C.byref(n)))
        return out[:n.value]


def compute_image(ctx, components, planes, out_w, out_h, color_transform):
    """decoder::compute_image (src/decoder.rs:1300-1336): host planes in, interleaved pixels out."""
    n = len(planes)
    ps = [np.ascontiguousarray(p, dtype=np.uint8).reshape(-1) for p in planes]
    pp = (C.c_void_p * max(n, 1))(*[p.ctypes.data if p.size else None for p in ps])
    pl = (C.c_size_t * max(n, 1))(*[p.size for p in ps])
    out = np.zeros(int(out_w) * int(out_h) * max(n, 1), dtype=np.uint8)
    ol = C.c_size_t()
    ctx.check(lib().b200jpg_compute_image(ctx._h, components, n, pp, pl, out_w, out_h, color_transform, _ptr(out), out.size, C.byref(ol)))
    return out[:ol.value]


class Batch:
    """b200jpg_batch: plan for n images (slab layout + device tables)."""

    def __init__(self, ctx, descs):
        self.ctx = ctx
        self.n = len(descs)
        self.descs = (ImageDesc * self.n)(*descs)
        self.statuses = (C.c_int * self.n)()
        h = C.c_void_p()
        ctx.check(lib().b200jpg_batch_create(ctx._h, self.descs, self.n, self.statuses, C.byref(h)))
        self._h = h
        self.info = BatchInfo()
        lib().b200jpg_batch_get_info(h, C.byref(self.info))

    def close(self):
        if getattr(self, "_h", None):
            lib().b200jpg_batch_free(self._h)
            self._h = None

    __del__ = close

    def layout(self, i):
        co, po = (C.c_size_t * 4)(), (C.c_size_t * 4)()
        oo, ol = C.c_size_t(), C.c_size_t()
        st = lib().b200jpg_batch_image_layout(self._h, i, co, po, C.byref(oo), C.byref(ol))
        return {"status": st, "coef_off": list(co), "plane_off": list(po), "out_off": oo.value, "out_len": ol.value}

    def run_device(self, d_coefs, d_planes, d_out, stages=3):
        """Enqueue K1 (bit 0) and/or K2 (bit 1) on the context's stream; raw device addresses."""
        self.ctx.check(lib().b200jpg_batch_run_device(self._h, d_coefs, d_planes, d_out, stages))

    def format_device(self, d_out, fmt, d_dst, scale=None, bias=None):
        """b200jpg_batch_format_device: interleaved RGB8 slab -> planar u8 / float NHWC / float NCHW (device pointers)."""
        sc = (C.c_float * 3)(*scale) if scale is not None else None
        bi = (C.c_float * 3)(*bias) if bias is not None else None
        st = (C.c_int * self.n)()
        self.ctx.check(lib().b200jpg_batch_format_device(self._h, d_out, fmt, d_dst, sc, bi, st))
        return list(st)

    def run_host(self, outs):
        """Host coefficients (from the descs) -> host pixels in `outs` (list of uint8 arrays or raw addresses)."""
        n = self.n
        addrs = [(o.ctypes.data if isinstance(o, np.ndarray) else int(o)) for o in outs]
        op = (C.c_void_p * n)(*addrs)
        caps = (C.c_size_t * n)(*[(o.size if isinstance(o, np.ndarray) else self.layout(i)["out_len"]) for i, o in enumerate(outs)])
        st = (C.c_int * n)()
        rc = lib().b200jpg_batch_run_host(self._h, self.descs, op, caps, st)
        self.statuses = st
        self.ctx.check(rc)
        return list(st)


def decode_batch(ctx, descs):
    """b200jpg_decode_batch: returns (list of uint8 arrays, list of status codes)."""
    n = len(descs)
    arr = (ImageDesc * n)(*descs)
    outs = [np.zeros(int(d.width) * int(d.height) * int(d.ncomp), dtype=np.uint8) for d in descs]
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    caps = (C.c_size_t * n)(*[o.size for o in outs])
    st = (C.c_int * n)()
    rc = lib().b200jpg_decode_batch(ctx._h, arr, n, op, caps, st)
    if rc and all(s == 0 for s in st):
        ctx.check(rc)
    return outs, list(st)


def decode_batch_sbs(ctx, descs, streams):
    """b200jpg_decode_batch_sbs: coefficients as sparse block streams [(uint8 array, order)] -> (pixels, statuses)."""
    n = len(descs)
    arr = (ImageDesc * n)(*descs)
    ss = (SbsStream * n)()
    for s, (buf, order) in zip(ss, streams):
        s.data, s.len, s.order = buf.ctypes.data, buf.size, order
    outs = [np.zeros(int(d.width) * int(d.height) * int(d.ncomp), dtype=np.uint8) for d in descs]
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    caps = (C.c_size_t * n)(*[o.size for o in outs])
    st = (C.c_int * n)()
    rc = lib().b200jpg_decode_batch_sbs(ctx._h, arr, ss, n, op, caps, st)
    if rc and all(s == 0 for s in st):
        ctx.check(rc)
    return outs, list(st)


def sbs_from_dense(desc):
    """b200jpg_sbs_from_dense: (uint8 stream, order) of an ImageDesc whose coefs point to dense coefficients."""
    nb = sum(int(desc.comps[c].block_w) * int(desc.comps[c].block_h) for c in range(desc.ncomp))
    buf = np.zeros(lib().b200jpg_sbs_worst_bytes(nb), dtype=np.uint8)
    s = SbsStream()
    rc = lib().b200jpg_sbs_from_dense(C.byref(desc), buf.ctypes.data, buf.size, C.byref(s))
    if rc:
        raise B200JpgError(rc, "b200jpg_sbs_from_dense failed")
    return buf[:s.len], s.order


def expand_sbs(ctx, desc, buf, order):
    """b200jpg_debug_expand_sbs: the dense coefficient arrays kernel K0 rebuilds from one stream."""
    s = SbsStream()
    s.data, s.len, s.order = buf.ctypes.data, buf.size, order
    dense = [np.zeros(int(desc.comps[c].block_w) * int(desc.comps[c].block_h) * 64, dtype=np.int16) for c in range(desc.ncomp)]
    ptrs = (C.c_void_p * 4)(*([d.ctypes.data for d in dense] + [None] * (4 - len(dense))))
    ctx.check(lib().b200jpg_debug_expand_sbs(ctx._h, C.byref(desc), C.byref(s), ptrs))
    return dense


def read_info_files(files, nthreads=0):
    """b200jpg_read_info_files: [(status, ImageInfo, out_len)] for a list of bytes objects."""
    n = len(files)
    bufs = [np.frombuffer(bytes(f), dtype=np.uint8) for f in files]
    jobs = (FileJob * n)()
    for j, b in zip(jobs, bufs):
        j.data, j.len = b.ctypes.data, b.size
    lib().b200jpg_read_info_files(jobs, n, nthreads)
    return [(j.status, j.info, j.out_len) for j in jobs]


def decode_files(ctx, files, nthreads=0, outs=None):
    """b200jpg_decode_files: list of bytes -> (list of uint8 arrays or None, list of status codes, list of ImageInfo)."""
    n = len(files)
    bufs = [np.frombuffer(bytes(f), dtype=np.uint8) for f in files]
    jobs = (FileJob * n)()
    for j, b in zip(jobs, bufs):
        j.data, j.len = b.ctypes.data, b.size
    lib().b200jpg_read_info_files(jobs, n, nthreads)
    if outs is None:
        outs = [np.zeros(j.out_len, dtype=np.uint8) if j.status == 0 else np.zeros(1, dtype=np.uint8) for j in jobs]
    for j, o in zip(jobs, outs):
        j.out, j.out_cap = o.ctypes.data, o.size
    rc = lib().b200jpg_decode_files(ctx._h, jobs, n, nthreads)
    if rc and all(j.status == 0 for j in jobs):
        ctx.check(rc)
    return [o[:j.out_len] if j.status == 0 else None for j, o in zip(jobs, outs)], [j.status for j in jobs], [j.info for j in jobs]


def decode_files_into(ctx, files, out_ptrs, out_caps, nthreads=0):
    """b200jpg_decode_files with caller-owned pixel buffers given as raw addresses (host or device memory of the
    context's GPU).  Returns (status codes, out_len per image)."""
    n = len(files)
    bufs = [np.frombuffer(bytes(f), dtype=np.uint8) for f in files]
    jobs = (FileJob * n)()
    for j, b, p, c in zip(jobs, bufs, out_ptrs, out_caps):
        j.data, j.len = b.ctypes.data, b.size
        j.out, j.out_cap = p, c
    rc = lib().b200jpg_decode_files(ctx._h, jobs, n, nthreads)
    if rc and all(j.status == 0 for j in jobs):
        ctx.check(rc)
    return [j.status for j in jobs], [j.out_len for j in jobs]


class Decoder:
    """jpeg_decoder::Decoder (src/decoder.rs:101-295) over an in-memory file."""

    def __init__(self, data, ctx=None):
        self.ctx = ctx
        self._buf = np.frombuffer(bytes(data), dtype=np.uint8)
        h = C.c_void_p()
        rc = lib().b200jpg_decoder_new(ctx._h if ctx else None, _ptr(self._buf), self._buf.size, C.byref(h))
        if rc:
            raise B200JpgError(rc, "b200jpg_decoder_new failed")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().b200jpg_decoder_free(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise B200JpgError(rc, lib().b200jpg_decoder_error(self._h).decode(errors="replace"))

    def set_color_transform(self, ct):
        lib().b200jpg_decoder_set_color_transform(self._h, ct)

    def set_max_decoding_buffer_size(self, n):
        lib().b200jpg_decoder_set_max_decoding_buffer_size(self._h, n)

    def read_info(self):
        self._check(lib().b200jpg_decoder_read_info(self._h))

    def info(self):
        inf = ImageInfo()
        return inf if lib().b200jpg_decoder_info(self._h, C.byref(inf)) else None

    def scale(self, w, h):
        ow, oh = C.c_uint16(), C.c_uint16()
        self._check(lib().b200jpg_decoder_scale(self._h, w, h, C.byref(ow), C.byref(oh)))
        return ow.value, oh.value

    def decode(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(lib().b200jpg_decoder_decode(self._h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()

    def entropy_decode(self):
        """Host half only: returns an ImageDesc whose pointers are owned by this decoder."""
        d = ImageDesc()
        self._check(lib().b200jpg_decoder_entropy_decode(self._h, C.byref(d)))
        return d

    def entropy_decode_sbs(self):
        """Host half only, coefficients as a sparse block stream: (ImageDesc, uint8 stream, order)."""
        nb = C.c_size_t()
        self._check(lib().b200jpg_decoder_total_blocks(self._h, C.byref(nb)))
        buf = np.zeros(lib().b200jpg_sbs_worst_bytes(nb.value), dtype=np.uint8)
        d, s = ImageDesc(), SbsStream()
        self._check(lib().b200jpg_decoder_entropy_decode_sbs(self._h, buf.ctypes.data, buf.size, C.byref(d), C.byref(s)))
        return d, buf[:s.len], s.order

    def coefficients(self, desc, i):
        c = desc.comps[i]
        n = int(c.block_w) * int(c.block_h) * 64
        return np.ctypeslib.as_array(C.cast(desc.coefs[i], C.POINTER(C.c_int16)), shape=(n,)).copy()

    def qtable(self, desc, i):
        return np.ctypeslib.as_array(C.cast(desc.qt[i], C.POINTER(C.c_uint16)), shape=(64,)).copy()

    def _blob(self, fn):
        p, n = C.c_void_p(), C.c_size_t()
        if not fn(self._h, C.byref(p), C.byref(n)):
            return None
        return C.string_at(p, n.value)

    def icc_profile(self):
        return self._blob(lib().b200jpg_decoder_icc_profile)

    def exif_data(self):
        return self._blob(lib().b200jpg_decoder_exif_data)

    def xmp_data(self):
        return self._blob(lib().b200jpg_decoder_xmp_data)
