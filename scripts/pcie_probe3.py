#!/usr/bin/env python
"""Pinned D2H bandwidth (27 x 6.2 MB per group, back to back) next to the uploads of the next groups, as a function of HOW the
17 MB of a group are uploaded: one copy, 27 copies of 631 KB, or those plus small table copies; and queued all at once
(like the library's submitter, which runs ahead) or one group per download."""
import json
import torch

dev = torch.device("cuda", 0)
per = 1920 * 1080 * 3
G = 27
NG = 16
src = torch.empty(G * per, dtype=torch.uint8, device=dev)
dst = torch.empty(128 * per, dtype=torch.uint8, pin_memory=True)
JB = 631372
up_h = torch.empty(G * JB * 4, dtype=torch.uint8, pin_memory=True)
up_d = torch.empty(G * JB, dtype=torch.uint8, device=dev)
s_out, s_in = torch.cuda.Stream(), torch.cuda.Stream()


filler = torch.randint(0, 255, (JB,), dtype=torch.uint8)


def upload(mode, g):
    if mode == "27dirty":   # the CPU has just written the sources (as the host threads do): their lines sit dirty in its caches
        for k in range(G):
            o = ((g * G + k) * 2654435761 % (3 * G)) * JB
            up_h[o:o + JB].copy_(filler)
        mode = "27"
    with torch.cuda.stream(s_in):
        if mode == "one":
            up_d.copy_(up_h[:G * JB], non_blocking=True)
        elif mode in ("27", "27+small"):
            for k in range(G):
                o = ((g * G + k) * 2654435761 % (3 * G)) * JB   # scattered sources, like the per-thread rings
                up_d[k * JB:(k + 1) * JB].copy_(up_h[o:o + JB], non_blocking=True)
            if mode == "27+small":
                for k in range(3):
                    up_d[k * 8192:k * 8192 + 5184].copy_(up_h[k * 8192:k * 8192 + 5184], non_blocking=True)


def run(mode, ahead):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_out):
        e0.record()
    with torch.cuda.stream(s_in):
        u0.record()
    if ahead and mode != "none":
        for g in range(NG):
            upload(mode, g)
    for g in range(NG):
        if not ahead and mode != "none":
            upload(mode, g)
        with torch.cuda.stream(s_out):
            for k in range(G):
                j = (g * G + k) % 128
                dst[j * per:(j + 1) * per].copy_(src[k * per:(k + 1) * per], non_blocking=True)
    with torch.cuda.stream(s_out):
        e1.record()
    with torch.cuda.stream(s_in):
        u1.record()
    torch.cuda.synchronize()
    return round(NG * G * per / e0.elapsed_time(e1) / 1e6, 2), round(NG * G * JB / max(u0.elapsed_time(u1), 1e-3) / 1e6, 2)


res = {}
for mode in ("none", "one", "27", "27+small", "27dirty"):
    for ahead in (True, False):
        run(mode, ahead)
        res["%s_%s" % (mode, "ahead" if ahead else "paced")] = max(run(mode, ahead) for _ in range(2))
print(json.dumps(res))
