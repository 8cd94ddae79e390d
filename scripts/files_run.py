"""Small whole-file run for profilers: n synthetic 1080p 4:2:0 JPEGs through b200jpg_decode_files (device entropy)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_decoder_b200 as J  # noqa: E402
from jpeg_decoder_b200 import workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = J.Context(device=0)
files = [workload.synth_jpeg(1920, 1080, seed=1234 + k, subsampling=2) for k in range(4)]
outs, st, _ = J.decode_files(ctx, [files[i % 4] for i in range(n)], nthreads=16)
assert st == [0] * n
print("scans", ctx.device_scan_counts, "launches", ctx.launch_count)
ctx.close()
