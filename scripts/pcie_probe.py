#!/usr/bin/env python
"""What slows the pixel download of b200jpg_decode_files (46 of 57 GB/s while the download stream never idles)?
Pinned D2H bandwidth alone and next to: a trickle of uploads (10 % of the volume, like the JPEG bytes), HBM-bound kernels
covering the whole window, compute-bound kernels, and the library's own kernel chain (decode_files with device outputs
running in a second thread)."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

dev = torch.device("cuda", 0)
CH = 118 << 20  # one group's pixels
N = 24
src = torch.empty(CH, dtype=torch.uint8, device=dev)
dst = torch.empty(CH * 2, dtype=torch.uint8, pin_memory=True)
up_h = torch.empty(CH // 10, dtype=torch.uint8, pin_memory=True)
up_d = torch.empty(CH // 10, dtype=torch.uint8, device=dev)
big_a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
big_b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
ma = torch.randn(8192, 8192, device=dev)
mb = torch.randn(8192, 8192, device=dev)
s_out, s_in, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d=False, kernels=None, split=19):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_k):
        if kernels == "hbm":
            for _ in range(160):   # 160 x 0.35 ms: longer than the download window
                big_b.copy_(big_a)
        elif kernels == "fp32":
            for _ in range(12):    # 12 x ~ 18 ms of fp32 matmul
                torch.mm(ma, mb)
    with torch.cuda.stream(s_out):
        e0.record()
    for i in range(N):
        if h2d:
            with torch.cuda.stream(s_in):
                up_d.copy_(up_h, non_blocking=True)
        with torch.cuda.stream(s_out):
            per = CH // split
            for j in range(split):
                dst[(i & 1) * CH + j * per:(i & 1) * CH + (j + 1) * per].copy_(src[j * per:(j + 1) * per], non_blocking=True)
    with torch.cuda.stream(s_out):
        e1.record()
    e1.synchronize()
    gbs = N * CH / e0.elapsed_time(e1) / 1e6
    torch.cuda.synchronize()
    return gbs


res = {}
torch.backends.cuda.matmul.allow_tf32 = False
for name, kw in (("d2h_alone", {}), ("d2h_with_upload_trickle", dict(h2d=True)), ("d2h_with_hbm_kernels", dict(kernels="hbm")),
                 ("d2h_with_fp32_matmul", dict(kernels="fp32")), ("d2h_alone_again", {})):
    run(**kw)
    res[name] = round(max(run(**kw) for _ in range(2)), 2)

# next to the library's own chain
import jpeg_decoder_b200 as J  # noqa: E402
from jpeg_decoder_b200 import workload  # noqa: E402

cfg = workload.CONFIGS["cfg2"]
W, H = cfg["width"], cfg["height"]
n = 256
jpegs = [np.frombuffer(workload.config_jpeg("cfg2", k), dtype=np.uint8) for k in range(4)]
per = W * H * 3
out = torch.empty(n * per, dtype=torch.uint8, device=dev)
jobs = (J.FileJob * n)()
for j in range(n):
    jobs[j].data, jobs[j].len = jpegs[j % 4].ctypes.data, jpegs[j % 4].size
    jobs[j].out, jobs[j].out_cap = out.data_ptr() + j * per, per
ctx = J.Context(device=0, host_threads=8)
L = J.lib()
for _ in range(3):
    ctx.check(L.b200jpg_decode_files(ctx._h, jobs, n, 8))
stop = False
calls = [0]


def loop():
    while not stop:
        ctx.check(L.b200jpg_decode_files(ctx._h, jobs, n, 8))
        calls[0] += 1


t = threading.Thread(target=loop)
t.start()
time.sleep(0.2)
res["d2h_with_decode_files_device_outputs"] = round(max(run() for _ in range(3)), 2)
c0, t0 = calls[0], time.perf_counter()
time.sleep(0.5)
res["decode_files_gpx_meanwhile"] = round((calls[0] - c0) * n * W * H / (time.perf_counter() - t0) / 1e9, 1)
stop = True
t.join()
ctx.close()
print(json.dumps(res))
