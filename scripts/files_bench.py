#!/usr/bin/env python
"""b200jpg_decode_files throughput on one GPU (profiling helper): n synthetic cfg2 files per call, pixels to pinned host
memory or to device memory.  Run several copies side by side (CUDA_VISIBLE_DEVICES + taskset) to see what the host
side of a multi-GPU box sustains.
  python scripts/files_bench.py [--n 512] [--reps 6] [--threads 0] [--dev-out] [--tag x]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_decoder_b200 as J  # noqa: E402
from jpeg_decoder_b200 import workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--dev-out", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--config", default="cfg2")
a = ap.parse_args()
cfg = workload.CONFIGS[a.config]
W, H = cfg["width"], cfg["height"]
nthreads = a.threads or len(os.sched_getaffinity(0))
jpegs = [workload.config_jpeg(a.config, k) for k in range(4)]
bufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]
per = W * H * 3
dev = torch.device("cuda", 0)
out = torch.empty(a.n * per, dtype=torch.uint8, device=dev) if a.dev_out else torch.empty(a.n * per, dtype=torch.uint8, pin_memory=True)
jobs = (J.FileJob * a.n)()
for j in range(a.n):
    jobs[j].data, jobs[j].len = bufs[j % 4].ctypes.data, bufs[j % 4].size
    jobs[j].out, jobs[j].out_cap = out.data_ptr() + j * per, per
ctx = J.Context(device=0, host_threads=nthreads)
L = J.lib()
for _ in range(4):
    ctx.check(L.b200jpg_decode_files(ctx._h, jobs, a.n, nthreads))
torch.cuda.synchronize()
times = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    ctx.check(L.b200jpg_decode_files(ctx._h, jobs, a.n, nthreads))
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
assert all(jobs[j].status == 0 for j in range(a.n))
mp = a.n * W * H / 1e6
print(json.dumps({"tag": a.tag, "lib": os.environ.get("B200JPG_SO", "libb200jpg.so"), "n": a.n, "threads": nthreads, "dev_out": a.dev_out,
                  "mps_median": mp / float(np.median(times)), "mps_best": mp / min(times), "ms": [round(1e3 * t, 2) for t in times],
                  "scans": list(map(int, ctx.device_scan_counts)), "cpus": len(os.sched_getaffinity(0))}), flush=True)
ctx.close()
