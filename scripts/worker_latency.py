#!/usr/bin/env python
"""The reference's trait Worker driven image after image through the C ABI (start / append_rows / get_result per component,
then compute_image on the device planes): latency per 1080p 4:2:0 image, against the oracle's pixels.
  python scripts/worker_latency.py [reps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_decoder_b200 as J  # noqa: E402
from jpeg_decoder_b200 import workload  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
u = workload.build_unique("cfg2", 1)[0]
ctx = J.Context(device=0)
w = J.Worker(ctx)


def one():
    for i, c in enumerate(u.components):
        w.start(i, c, u.qts[i])
        w.append_rows(i, u.coefs[i], c.block_h // c.v)
    # get_result(index, NULL) computes the plane and leaves it on the device for compute_image
    n = J.C.c_size_t()
    for i in range(u.ncomp):
        ctx.check(J.lib().b200jpg_worker_get_result(w._h, i, None, 0, J.C.byref(n)))
    return w.compute_image(u.ncomp, u.width, u.height, u.color_transform)


px = one()
import oracle  # noqa: E402
want = oracle.Decoder(u.jpeg).decode()
ok = bool(np.array_equal(px, want.reshape(-1)))
times = []
for _ in range(reps):
    t0 = time.perf_counter()
    one()
    times.append(1e3 * (time.perf_counter() - t0))
print(json.dumps({"worker_route_ms_per_1080p_image": {"median": round(float(np.median(times)), 3), "best": round(min(times), 3)},
                  "bit_exact_vs_oracle": ok, "reps": reps}))
w.close()
ctx.close()
