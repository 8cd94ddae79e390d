"""ctypes binding of jpeg_decoder_b200/libb200jpg.so (the C ABI declared in include/b200jpg.h).

No fallback: if the library is missing it is built with nvcc; if that fails, import of the
native layer raises.  Device work raises B200JpgError(ERR_INTERNAL) when CUDA is unavailable."""
import ctypes as C
import os

from . import build as _build

OK, ERR_FORMAT, ERR_UNSUPPORTED, ERR_IO, ERR_INTERNAL = 0, -1, -2, -3, -4
CT_NONE, CT_UNKNOWN, CT_GRAYSCALE, CT_RGB, CT_YCBCR, CT_CMYK, CT_YCCK, CT_JCS_BG_YCC, CT_JCS_BG_RGB = range(9)
PF_L8, PF_L16, PF_RGB24, PF_CMYK32 = range(4)
CP_DCT_SEQUENTIAL, CP_DCT_PROGRESSIVE, CP_LOSSLESS = range(3)
ARITH_SCALAR, ARITH_SSSE3 = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FAST = 0, 1, 2


class Component(C.Structure):
    """parser::Component, reference src/parser.rs:77-89"""
    _fields_ = [("identifier", C.c_uint8), ("h", C.c_uint8), ("v", C.c_uint8), ("tq", C.c_uint8),
                ("dct_scale", C.c_uint16), ("size_w", C.c_uint16), ("size_h", C.c_uint16),
                ("block_w", C.c_uint16), ("block_h", C.c_uint16)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("arith", C.c_int), ("k1_kernel", C.c_int), ("k2_kernel", C.c_int),
                ("stream", C.c_void_p), ("host_compact", C.c_int), ("host_threads", C.c_int), ("entropy", C.c_int), ("fuse", C.c_int)]


class ImageDesc(C.Structure):
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("ncomp", C.c_uint8), ("color_transform", C.c_uint8),
                ("reserved", C.c_uint16), ("comps", Component * 4), ("qt", C.c_void_p * 4), ("coefs", C.c_void_p * 4)]


class BatchInfo(C.Structure):
    _fields_ = [("coef_bytes", C.c_size_t), ("plane_bytes", C.c_size_t), ("out_bytes", C.c_size_t),
                ("n_blocks", C.c_size_t), ("n_pixels", C.c_size_t), ("k1_algorithmic_bytes", C.c_size_t),
                ("k2_algorithmic_bytes", C.c_size_t), ("kf_algorithmic_bytes", C.c_size_t), ("n_fused", C.c_size_t)]


class ImageInfo(C.Structure):
    """ImageInfo, reference src/decoder.rs:64-74"""
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("pixel_format", C.c_int), ("coding_process", C.c_int)]


class FileJob(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("out", C.c_void_p), ("out_cap", C.c_size_t),
                ("info", ImageInfo), ("out_len", C.c_size_t), ("status", C.c_int)]


class SbsStream(C.Structure):
    """b200jpg_sbs_stream: one image's sparse block stream (csrc/sbs.h)"""
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("order", C.c_int)]


SBS_PLANAR, SBS_INTERLEAVED, SBS_NATURAL = 0, 1, 2
COMPACT_AUTO, COMPACT_OFF, COMPACT_ON = 0, 1, 2
ENTROPY_AUTO, ENTROPY_HOST, ENTROPY_DEVICE = 0, 1, 2
FUSE_AUTO, FUSE_OFF, FUSE_ON = 0, 1, 2
FMT_RGB8_PLANAR, FMT_RGB_F32_NHWC, FMT_RGB_F32_NCHW = 1, 2, 3

# every symbol include/b200jpg.h declares (tests/test_abi.py checks the two lists agree)
EXPORTS = {
    "b200jpg_default_options": (None, [C.POINTER(Options)]),
    "b200jpg_create": (C.c_int, [C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "b200jpg_destroy": (None, [C.c_void_p]),
    "b200jpg_last_error": (C.c_char_p, [C.c_void_p]),
    "b200jpg_version": (C.c_char_p, []),
    "b200jpg_launch_count": (C.c_uint64, [C.c_void_p]),
    "b200jpg_device_scan_counts": (None, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b200jpg_synchronize": (C.c_int, [C.c_void_p]),
    "b200jpg_set_fuse": (None, [C.c_void_p, C.c_int]),
    "b200jpg_update_component_sizes": (C.c_int, [C.c_uint16, C.c_uint16, C.POINTER(Component), C.c_int,
                                                 C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]),
    "b200jpg_choose_idct_size": (C.c_int, [C.c_uint16] * 4),
    "b200jpg_worker_new": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200jpg_worker_free": (None, [C.c_void_p]),
    "b200jpg_worker_start": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Component), C.c_void_p]),
    "b200jpg_worker_append_row": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "b200jpg_worker_append_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "b200jpg_worker_get_result": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200jpg_worker_compute_image": (C.c_int, [C.c_void_p, C.c_int, C.c_uint16, C.c_uint16, C.c_int, C.c_void_p,
                                               C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200jpg_compute_image": (C.c_int, [C.c_void_p, C.POINTER(Component), C.c_int, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t), C.c_uint16, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]),
    "b200jpg_batch_create": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "b200jpg_batch_free": (None, [C.c_void_p]),
    "b200jpg_batch_get_info": (C.c_int, [C.c_void_p, C.POINTER(BatchInfo)]),
    "b200jpg_batch_image_layout": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "b200jpg_batch_run_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "b200jpg_batch_format_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "b200jpg_batch_run_host": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_int)]),
    "b200jpg_decode_batch": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.c_size_t, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "b200jpg_debug_set_kernel_modes": (None, [C.c_int, C.c_int]),
    "b200jpg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "b200jpg_host_free": (None, [C.c_void_p]),
    "b200jpg_decoder_new": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "b200jpg_decoder_free": (None, [C.c_void_p]),
    "b200jpg_decoder_set_color_transform": (None, [C.c_void_p, C.c_int]),
    "b200jpg_decoder_set_max_decoding_buffer_size": (None, [C.c_void_p, C.c_size_t]),
    "b200jpg_decoder_read_info": (C.c_int, [C.c_void_p]),
    "b200jpg_decoder_info": (C.c_int, [C.c_void_p, C.POINTER(ImageInfo)]),
    "b200jpg_decoder_scale": (C.c_int, [C.c_void_p, C.c_uint16, C.c_uint16, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]),
    "b200jpg_decoder_decode": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200jpg_decoder_error": (C.c_char_p, [C.c_void_p]),
    "b200jpg_decoder_icc_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200jpg_decoder_exif_data": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200jpg_decoder_xmp_data": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200jpg_decoder_entropy_decode": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc)]),
    "b200jpg_sbs_worst_bytes": (C.c_size_t, [C.c_size_t]),
    "b200jpg_decoder_total_blocks": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "b200jpg_decoder_entropy_decode_sbs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(ImageDesc), C.POINTER(SbsStream)]),
    "b200jpg_sbs_from_dense": (C.c_int, [C.POINTER(ImageDesc), C.c_void_p, C.c_size_t, C.POINTER(SbsStream)]),
    "b200jpg_decode_batch_sbs": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.POINTER(SbsStream), C.c_size_t,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "b200jpg_debug_expand_sbs": (C.c_int, [C.c_void_p, C.POINTER(ImageDesc), C.POINTER(SbsStream), C.POINTER(C.c_void_p)]),
    "b200jpg_read_info_files": (C.c_int, [C.POINTER(FileJob), C.c_size_t, C.c_int]),
    "b200jpg_decode_files": (C.c_int, [C.c_void_p, C.POINTER(FileJob), C.c_size_t, C.c_int]),
}

_lib = None


def lib():
    """Loads (building first if needed) libb200jpg.so.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is None:
        so = _build.build()
        L = C.CDLL(so)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)  # AttributeError = missing export: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
