#!/usr/bin/env python
"""Pinned D2H bandwidth as a function of the size of the destination region (IOMMU / DDIO effects on this box):
the library's pixel download marches through a 3.2 GB user buffer, the bench's PCIe probe reuses a small one."""
import json
import torch

dev = torch.device("cuda", 0)
per = 1920 * 1080 * 3
src = torch.empty(27 * per, dtype=torch.uint8, device=dev)
res = {}
for label, nimg in (("236MB", 38), ("800MB", 128), ("3.2GB", 512)):
    dst = torch.empty(nimg * per, dtype=torch.uint8, pin_memory=True)
    dst.zero_()
    best = 0.0
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        total = 0
        for j in range(512):
            k = j % nimg
            dst[k * per:(k + 1) * per].copy_(src[(j % 27) * per:(j % 27 + 1) * per], non_blocking=True)
            total += per
        e1.record()
        e1.synchronize()
        best = max(best, total / e0.elapsed_time(e1) / 1e6)
    res["d2h_dst_" + label] = round(best, 2)
    # strided order like the library's (jobs of a group are not neighbours)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(512):
        k = (j * 37) % nimg
        dst[k * per:(k + 1) * per].copy_(src[(j % 27) * per:(j % 27 + 1) * per], non_blocking=True)
    e1.record()
    e1.synchronize()
    res["d2h_dst_" + label + "_scattered"] = round(512 * per / e0.elapsed_time(e1) / 1e6, 2)
    del dst
print(json.dumps(res))
