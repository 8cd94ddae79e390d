"""Debug helper: per-file status of b200jpg_decode_files with device vs host entropy decoding over the fixtures."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import jpeg_decoder_b200 as J  # noqa: E402
from jpeg_decoder_b200 import workload  # noqa: E402

g = os.path.join(ROOT, "tests", "golden")
paths = sorted(glob.glob(os.path.join(g, "reftest", "**", "*.jpg"), recursive=True)) + sorted(glob.glob(os.path.join(g, "benches", "*.jpg")))
files = [open(p, "rb").read() for p in paths]
dev = J.Context(device=0, entropy=J.ENTROPY_DEVICE)
host = J.Context(device=0, entropy=J.ENTROPY_HOST)
for rep in range(2):
    o1, s1, _ = J.decode_files(dev, files, nthreads=4)
    o2, s2, _ = J.decode_files(host, files, nthreads=4)
    for p, a, b, x, y in zip(paths, s1, s2, o1, o2):
        same = (x is None and y is None) or (x is not None and y is not None and np.array_equal(x, y))
        if a != b or not same:
            print("DIFF", os.path.basename(p), "device", a, "host", b, "pixels equal", same, dev.last_error() if hasattr(dev, "last_error") else "")
    print("rep", rep, "scan counts", dev.device_scan_counts)
big = [workload.synth_jpeg(1920, 1080, seed=1234 + k, subsampling=2) for k in range(4)]
for n in (8, 33, 64, 96):
    o, s, _ = J.decode_files(dev, [big[i % 4] for i in range(n)], nthreads=16)
    print("n", n, "bad statuses", sum(1 for x in s if x), "scan counts", dev.device_scan_counts)
