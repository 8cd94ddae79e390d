#!/usr/bin/env python
"""Times the bit-identical code-generation variants of K1 / K2 on a device-resident workload and checks
that every variant produces the same bytes.  Profiling helper (run under gpurun); its output is committed
under profiles/ as the evidence for the defaults chosen in k1_idct.cu / k2_color.cu."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_decoder_b200 as J  # noqa: E402
from jpeg_decoder_b200 import workload  # noqa: E402

B = int(os.environ.get("SWEEP_BATCH", "512"))
cfgname = os.environ.get("SWEEP_CONFIG", "cfg2")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
uniq = workload.build_unique(cfgname, 4)
keep, descs = [], []
for i in range(B):
    u = uniq[i % len(uniq)]
    descs.append(J.make_image_desc(u.width, u.height, u.components, u.qts, u.coefs, u.color_transform, keep))
ref_sum = None
for kernels in (("auto",) if (os.environ.get("SWEEP_ONLY_AUTO") or os.environ.get("SWEEP_SKIP_GENERIC")) else ("auto", "generic")):
    k = J.KERNEL_AUTO if kernels == "auto" else J.KERNEL_GENERIC
    ctx = J.Context(device=0, k1_kernel=k, k2_kernel=k, stream=stream.cuda_stream)
    batch = J.Batch(ctx, descs)
    info = batch.info
    d_coefs = torch.empty(info.coef_bytes, dtype=torch.uint8, device=dev)
    d_planes = torch.empty(info.plane_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(info.out_bytes, dtype=torch.uint8, device=dev)
    du = [[torch.from_numpy(c.view(np.uint8)).to(dev) for c in u.coefs] for u in uniq]
    for j in range(B):
        lay = batch.layout(j)
        for kk, src in enumerate(du[j % len(uniq)]):
            d_coefs[lay["coef_off"][kk]:lay["coef_off"][kk] + src.numel()].copy_(src)
    torch.cuda.synchronize()

    def timed(stages, steps=10):
        for _ in range(3):
            batch.run_device(d_coefs.data_ptr(), d_planes.data_ptr(), d_out.data_ptr(), stages)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            batch.run_device(d_coefs.data_ptr(), d_planes.data_ptr(), d_out.data_ptr(), stages)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    modes = [(5, 0), (8, 0), (5, 1), (0, 0)] if kernels == "auto" else [(-1, -1)]
    if os.environ.get("SWEEP_ONLY_AUTO"):
        modes = [(-1, -1)]  # whatever B200JPG_K1_MODE / B200JPG_K2_MODE say (profiling runs)
    for k1m, k2m in modes:
        J.lib().b200jpg_debug_set_kernel_modes(k1m, k2m)
        d_planes.zero_()
        d_out.zero_()
        ms1, ms2 = timed(1), timed(2)
        os.environ["B200JPG_FUSE"] = "0"
        ms12 = timed(3)
        batch.run_device(d_coefs.data_ptr(), d_planes.data_ptr(), d_out.data_ptr(), 3)
        torch.cuda.synchronize()
        csum = int(d_out.to(torch.int64).sum().item()) * 31 + int(d_planes.to(torch.int64).sum().item())
        ref_sum = csum if ref_sum is None else ref_sum
        # the fused kernel (planes never written): same pixels, so compare the pixel slab alone
        psum = int(d_out.to(torch.int64).sum().item())
        os.environ["B200JPG_FUSE"] = "1"
        d_out.zero_()
        msf = timed(3) if info.n_fused == B else None
        torch.cuda.synchronize()
        fsum = int(d_out.to(torch.int64).sum().item())
        del os.environ["B200JPG_FUSE"]
        mp = info.n_pixels / 1e6
        print(json.dumps({"config": cfgname, "kernels": kernels, "k1_mode": k1m, "k2_mode": k2m, "batch": B, "k1_ms": ms1, "k2_ms": ms2,
                          "k1_gbs": info.k1_algorithmic_bytes / ms1 / 1e6, "k2_gbs": info.k2_algorithmic_bytes / ms2 / 1e6,
                          "k1k2_ms": ms12, "k1k2_mps": mp / ms12 * 1e3, "fused_ms": msf, "fused_mps": mp / msf * 1e3 if msf else None,
                          "fused_gbs": info.kf_algorithmic_bytes / msf / 1e6 if msf else None, "fused_same_pixels": fsum == psum if msf else None,
                          "same_output_as_first": csum == ref_sum}), flush=True)
    J.lib().b200jpg_debug_set_kernel_modes(-1, -1)
    batch.close()
    ctx.close()
