#!/usr/bin/env python
"""Uploads while the download stream saturates the link: copy engine (one 17 MB copy per group) against a kernel that reads the
page-locked source through its device mapping (zero copy), with 8 / 32 / 128 CTAs.  Prints download GB/s and upload GB/s."""
import ctypes
import json
import os

import torch

dev = torch.device("cuda", 0)
per = 1920 * 1080 * 3
G, NG = 27, 16
src = torch.empty(G * per, dtype=torch.uint8, device=dev)
dst = torch.empty(128 * per, dtype=torch.uint8, pin_memory=True)
UP = 17 << 20
up_h = torch.empty(UP * 2, dtype=torch.uint8, pin_memory=True)
up_d = torch.empty(UP, dtype=torch.uint8, device=dev)
s_out, s_in = torch.cuda.Stream(), torch.cuda.Stream()
zc = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprobe_zc.so"))
zc.probe_zc_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


filler = torch.randint(0, 255, (UP,), dtype=torch.uint8)
torch.set_num_threads(1)


def upload(mode, g):
    off = (g & 1) * UP
    if mode == "ce_dirty":   # the CPU has just written the source (like the submitter's gather): lines dirty in its caches
        up_h[off:off + UP].copy_(filler)
        mode = "ce"
    with torch.cuda.stream(s_in):
        if mode == "ce":
            up_d.copy_(up_h[off:off + UP], non_blocking=True)
        elif mode.startswith("zc"):
            rc = zc.probe_zc_copy(up_h.data_ptr() + off, up_d.data_ptr(), UP, int(mode[2:]), s_in.cuda_stream)
            assert rc == 0, rc


def run(mode, download=True):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_out):
        e0.record()
    with torch.cuda.stream(s_in):
        u0.record()
    for g in range(NG):
        if mode != "none":
            upload(mode, g)
        if download:
            with torch.cuda.stream(s_out):
                for k in range(G):
                    j = (g * G + k) % 128
                    dst[j * per:(j + 1) * per].copy_(src[k * per:(k + 1) * per], non_blocking=True)
    with torch.cuda.stream(s_out):
        e1.record()
    with torch.cuda.stream(s_in):
        u1.record()
    torch.cuda.synchronize()
    return (round(NG * G * per / e0.elapsed_time(e1) / 1e6, 2) if download else None, round(NG * UP / max(u0.elapsed_time(u1), 1e-3) / 1e6, 2) if mode != "none" else None)


res = {}
for mode in ("none", "ce", "ce_dirty", "zc128"):
    run(mode)
    res[mode + "+download"] = max(run(mode) for _ in range(2))
    if mode not in ("none",):
        run(mode, False)
        res[mode + "_alone"] = run(mode, False)[1]
print(json.dumps(res))
