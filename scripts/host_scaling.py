#!/usr/bin/env python
"""Host-only: how does the Huffman feeder scale with threads on this box?  (profiling helper)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_decoder_b200 as J
from jpeg_decoder_b200 import workload
data = workload.synth_jpeg(1920, 1080, 1234, 2)
def work(m):
    for _ in range(m):
        d = J.Decoder(data); d.entropy_decode(); d.close()
work(2)
for n in (1, 2, 4, 8, 16, 32, 64, 128):
    if n > (os.cpu_count() or 1): break
    m = 8
    th = [threading.Thread(target=work, args=(m,)) for _ in range(n)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print("threads %3d: %.2f ms per image per thread, %.0f MP/s total" % (n, dt / m * 1e3, n * m * 1920 * 1080 / 1e6 / dt), flush=True)
