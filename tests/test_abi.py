"""The C-ABI library loads without a GPU and exports every symbol include/b200jpg.h declares; device entry
points fail loudly (ERR_INTERNAL) instead of falling back to the CPU.  CPU only."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200jpg.h")).read()
    return sorted(set(re.findall(r"B200JPG_API[^;(]*?\b(b200jpg_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(J):
    from jpeg_decoder_b200 import _native
    names = declared_symbols()
    assert len(names) >= 35
    L = J.lib()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_native.EXPORTS) == names  # the binding covers the header, nothing more, nothing less


def test_product_does_not_link_or_import_the_oracle():
    pkg = os.path.join(ROOT, "jpeg_decoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "orc_" not in text, f


def test_no_cpu_fallback_without_gpu(J):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(J.B200JpgError) as e:
        J.Context(device=0)
    assert e.value.code == J.ERR_INTERNAL
    data = open(os.path.join(ROOT, "tests", "golden", "reftest", "mozilla", "jpg-size-8x8.jpg"), "rb").read()
    d = J.Decoder(data)  # host half works without a context ...
    with pytest.raises(J.B200JpgError) as e2:
        d.decode()       # ... the worker path does not exist on the CPU
    assert e2.value.code == J.ERR_INTERNAL


def test_version_string(J):
    assert b"b200jpg" in J.lib().b200jpg_version()
