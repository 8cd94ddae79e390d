#!/usr/bin/env python
"""bench.py -- megapixels/s of the JPEG block pipeline hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2]        # this repo's CUDA path
  python bench.py --impl reference ...                                        # CPU reference arm

One "step" = one pass of the hot path (dequant+IDCT -> planes -> upsample+colour -> pixels) over a
batch of synthetic images whose dense coefficients are already resident in HBM.  Prints ONE JSON
line (rank 0).  N>1 is launched by torchrun, one rank per GPU; images are sharded by index
(weak scaling: `batch` images per GPU), no data-path collective.

`--impl reference`: the reference is a Rust crate and cannot be built in this image (no cargo /
rustc), so the reference arm times the C restatement of the reference's CPU path (oracle/, kind
"port") on all host cores, on the same workload definition.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the config's)")
    ap.add_argument("--unique", type=int, default=8, help="distinct synthetic images (replicated to the batch)")
    ap.add_argument("--e2e-batch", type=int, default=256, help="images per end-to-end (host->host) step")
    ap.add_argument("--arith", default="scalar", choices=["scalar", "ssse3"])
    ap.add_argument("--k1", default="auto", choices=["auto", "generic"], help="K1 kernel variant (profiling)")
    ap.add_argument("--k2", default="auto", choices=["auto", "generic"], help="K2 kernel variant (profiling)")
    ap.add_argument("--host-compact", default="auto", choices=["auto", "on", "off"],
                    help="b200jpg_batch_run_host: compact the dense coefficients into sparse block streams on host threads")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled every few milliseconds through NVML while kernels run (the timed
    region is tens of milliseconds long: far too short for `nvidia-smi -lms`)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []      # (time, sm_mhz, reasons bitmask)
        self.windows = []      # (t0, t1) of the timed regions
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        self.error = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            try:
                self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            except Exception as e:  # noqa: BLE001
                self.error = "max clock: %r" % (e,)

            def reasons():
                for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
                    fn = getattr(pynvml, name, None)
                    if fn is not None:
                        try:
                            return int(fn(h))
                        except Exception:  # noqa: BLE001
                            continue
                return 0
            self.ok = True
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), reasons()))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)
            self.ok = False
        if not self.samples:
            self.run_smi()

    def run_smi(self):
        """Fallback: nvidia-smi in a loop (coarser: ~20 ms per query)."""
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0]
                f = [x.strip() for x in out.split(",")]
                mask = 0
                for b, v in zip(bits, f[2:6]):
                    if v.lower().startswith("active"):
                        mask |= b
                self.max_mhz = float(f[1])
                self.samples.append((time.perf_counter(), float(f[0]), mask))
            except Exception as e:  # noqa: BLE001
                self.error = repr(e)
                time.sleep(0.05)

    def stop(self):
        self.stop_flag = True

    def summary(self):
        inside = [s for s in self.samples if any(t0 <= s[0] <= t1 for t0, t1 in self.windows)]
        use = inside if inside else self.samples
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "error": self.error}
        mask = 0
        for s_ in use:
            mask |= s_[2]
        reasons = sorted(name for bit, name in self.REASONS.items() if mask & bit)
        return {"sm_mhz": float(np.median([s_[1] for s_ in use])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(use), "samples_inside_timed_regions": len(inside),
                "source": "NVML polled every 2 ms (nvidia-smi loop as fallback) while the CUDA-event timed regions run"}


def bind_to_gpu_numa_node(torch, index):
    """Pins this process to the CPUs of the GPU's NUMA node so that pinned buffers are allocated next to the
    GPU's PCIe root (host<->device copies then do not cross the socket interconnect)."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def pcie_probe(torch, dev, nbytes=1 << 30):
    """Measured pinned-memory copy bandwidth (GB/s): the roofline of the end-to-end (host->host) number."""
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return 3 * nbytes / (time.perf_counter() - t0) / 1e9
    run(True, True)
    return {"h2d": run(True, False), "d2h": run(False, True), "bidir_each": run(True, True)}


def cpu_port_throughput(unique, nthreads, target_seconds, arith):
    """Times oracle.hotpath_batch (C restatement of the reference CPU path) on a bounded sample."""
    import oracle
    u0 = unique[0]
    comps = (oracle.Component * u0.ncomp)()
    for i, c in enumerate(u0.components):
        for f, _ in oracle.Component._fields_:
            setattr(comps[i], f, getattr(c, f))
    n = max(nthreads * 2, 2)
    coefs = [unique[i % len(unique)].coefs for i in range(n)]
    outs = [np.zeros(u0.width * u0.height * u0.ncomp, dtype=np.uint8) for _ in range(n)]
    a = oracle.ARITH_SSSE3 if arith == "ssse3" else oracle.ARITH_SCALAR
    t0 = time.perf_counter()
    oracle.hotpath_batch(comps, u0.qts, coefs, u0.width, u0.height, u0.color_transform, outs, nthreads, a)
    pilot = time.perf_counter() - t0
    reps = max(1, min(200, int(target_seconds / max(pilot, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.hotpath_batch(comps, u0.qts, coefs, u0.width, u0.height, u0.color_transform, outs, nthreads, a)
    dt = time.perf_counter() - t0
    mp = reps * n * u0.width * u0.height / 1e6
    return mp / dt, "%d images x %d passes (%.1f s), dense coefficients -> RGB, one image per thread" % (n, reps, dt), outs[0]


def cpu_files_throughput(jpegs, width, height, nthreads, target_seconds, arith):
    """Whole files on the CPU: the oracle's restatement of Decoder::decode() (marker parsing, Huffman, IDCT, upsampling,
    colour), one image per host thread -- an outer par_iter over the reference's decoder."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    L = oracle.lib()
    a = oracle.ARITH_SSSE3 if arith == "ssse3" else oracle.ARITH_SCALAR
    bufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]

    def one(k):   # ctypes drops the GIL for the duration of each call
        b = bufs[k % len(bufs)]
        d = L.orc_decoder_new(b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, a)
        p, n = C.c_void_p(), C.c_size_t()
        rc = L.orc_decoder_decode(d, C.byref(p), C.byref(n))
        L.orc_decoder_free(d)
        return rc
    with ThreadPoolExecutor(nthreads) as ex:
        t0 = time.perf_counter()
        assert all(r == 0 for r in ex.map(one, range(nthreads)))
        pilot = time.perf_counter() - t0
        n = nthreads * max(1, min(64, int(target_seconds / max(pilot, 1e-3))))
        t0 = time.perf_counter()
        assert all(r == 0 for r in ex.map(one, range(n)))
        dt = time.perf_counter() - t0
    return n * width * height / 1e6 / dt, "%d files (%.1f s), JPEG bytes -> RGB, one image per thread" % (n, dt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from jpeg_decoder_b200 import workload
    cfg = workload.CONFIGS[args.config]
    unique = workload.build_unique(args.config, min(args.unique, 4))
    cores = os.cpu_count() or 1
    vals = []
    sample = ""
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_port_throughput(unique, cores, per_step / 2, args.arith)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sample, _ = cpu_port_throughput(unique, cores, per_step, args.arith)
        vals.append(v)
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    jpegs = [workload.synth_jpeg(cfg["width"], cfg["height"], cfg["seed"] + k, cfg["subsampling"]) for k in range(2)]
    fv, fsample = cpu_files_throughput(jpegs, cfg["width"], cfg["height"], cores, 5.0, args.arith)
    line = {
        "impl": "reference", "metric": "megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "width": cfg["width"], "height": cfg["height"], "arith": args.arith,
                   "note": "reference is Rust (no toolchain here): C restatement of its CPU hot path, one image per host thread"},
        "cpu_baseline": {"value": value, "unit": "MP/s", "cores": cores, "kind": "port", "sample": "per step: " + sample,
                         "files": {"value": fv, "unit": "MP/s", "sample": fsample}},
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON: everything else that writes to file descriptor 1 from here on -- NCCL's
    "NCCL version ..." banner, library chatter -- goes to stderr; emit() writes the result to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import jpeg_decoder_b200 as J
    from jpeg_decoder_b200 import workload

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"   # stdout carries ONE line, the JSON: no "NCCL version ..." banner
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload.CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    W, H = cfg["width"], cfg["height"]
    all_cpus = os.sched_getaffinity(0)
    numa = None if args.no_numa_bind else bind_to_gpu_numa_node(torch, local_rank)
    # control plane: rank 0 broadcasts the image -> GPU assignment (contiguous index ranges)
    table = workload.broadcast_assignment(B * world, world, dist if world > 1 else None, device=dev)
    lo, hi = int(table[rank, 0]), int(table[rank, 1])
    assert hi - lo == B
    unique = workload.build_unique(args.config, args.unique)
    U = len(unique)

    # an explicit (non-default) torch stream: the library enqueues on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    arith = J.ARITH_SSSE3 if args.arith == "ssse3" else J.ARITH_SCALAR
    kmap = {"auto": J.KERNEL_AUTO, "generic": J.KERNEL_GENERIC}
    assert stream.cuda_stream != 0
    cmap = {"auto": J.COMPACT_AUTO, "on": J.COMPACT_ON, "off": J.COMPACT_OFF}
    ctx = J.Context(device=local_rank, arith=arith, k1_kernel=kmap[args.k1], k2_kernel=kmap[args.k2], stream=stream.cuda_stream,
                    host_compact=cmap[args.host_compact])
    keep = []
    descs = []
    for i in range(lo, hi):
        u = unique[i % U]
        descs.append(J.make_image_desc(u.width, u.height, u.components, u.qts, u.coefs, u.color_transform, keep))
    batch = J.Batch(ctx, descs)
    info = batch.info
    d_coefs = torch.empty(info.coef_bytes, dtype=torch.uint8, device=dev)
    d_planes = torch.empty(info.plane_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(info.out_bytes, dtype=torch.uint8, device=dev)
    # upload each unique image once, replicate on the device
    dev_unique = [[torch.from_numpy(c.view(np.uint8)).to(dev) for c in u.coefs] for u in unique]
    for j in range(B):
        lay = batch.layout(j)
        for k, src in enumerate(dev_unique[(lo + j) % U]):
            d_coefs[lay["coef_off"][k]:lay["coef_off"][k] + src.numel()].copy_(src)
    torch.cuda.synchronize()

    def run(stages=3):
        batch.run_device(d_coefs.data_ptr(), d_planes.data_ptr(), d_out.data_ptr(), stages)

    def timed(stages, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            run(stages)
        e1.record(stream)
        torch.cuda.synchronize()
        sampler.windows.append((t0, time.perf_counter()))
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ctx.launch_count
    ms = timed(3, args.steps)
    launches = ctx.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        dist.barrier()
    else:
        ms_max = ms
    # per-kernel times for the roofline (same inputs, same stream, CUDA events)
    ms_k1 = timed(1, args.steps) / args.steps
    ms_k2 = timed(2, args.steps) / args.steps
    sampler.stop()
    clocks = sampler.summary()

    mp_per_step = world * B * W * H / 1e6
    value = mp_per_step * args.steps / (ms_max / 1e3)
    peak, peak_src = load_peaks()
    k1_gbs = info.k1_algorithmic_bytes / (ms_k1 * 1e-3) / 1e9
    k2_gbs = info.k2_algorithmic_bytes / (ms_k2 * 1e-3) / 1e9
    dominant = "k1_dequant_idct8x8" if ms_k1 >= ms_k2 else "k2_upsample_color"
    dom_gbs = k1_gbs if ms_k1 >= ms_k2 else k2_gbs
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    # same command (profiles/traffic.json, written by scripts/ncu_traffic.py); null if never captured
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get("%s:%d" % (args.config, B), {}).get(dominant)
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
        "traffic": traffic, "peak_source": peak_src,
        "kernels": {
            "k1_dequant_idct8x8": {"ms": ms_k1, "algorithmic_bytes": info.k1_algorithmic_bytes, "achieved": k1_gbs, "frac": k1_gbs / peak},
            "k2_upsample_color": {"ms": ms_k2, "algorithmic_bytes": info.k2_algorithmic_bytes, "achieved": k2_gbs, "frac": k2_gbs / peak},
        },
    }

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    Be = min(args.e2e_batch, B)
    u0 = unique[0]
    coef_per_img = u0.coef_bytes
    out_per_img = W * H * u0.ncomp
    h_in = torch.empty(Be * coef_per_img, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(Be * out_per_img, dtype=torch.uint8, pin_memory=True)
    h_in_np = h_in.numpy()
    e_descs, e_keep = [], []
    for j in range(Be):
        u = unique[(lo + j) % U]
        off = j * coef_per_img
        views = []
        for c in u.coefs:
            v = h_in_np[off:off + c.nbytes].view(np.int16)
            v[:] = c
            views.append(v)
            off += c.nbytes
        e_descs.append(J.make_image_desc(u.width, u.height, u.components, u.qts, views, u.color_transform, e_keep))
    e_batch = J.Batch(ctx, e_descs)
    outs = [h_out.data_ptr() + j * out_per_img for j in range(Be)]
    for _ in range(5):   # >= 4: host_compact=auto times two dense and two compacted runs before it settles
        e_batch.run_host(outs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        st = e_batch.run_host(outs)
    e_dt = time.perf_counter() - t0
    assert all(s == 0 for s in st)
    if world > 1:
        t = torch.tensor([e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.item())
    e2e_value = world * Be * W * H / 1e6 * e_steps / e_dt
    # the same call with each upload strategy forced (host_compact=auto, above, measures both itself and keeps the faster)
    forced = None
    if rank == 0 and world == 1 and args.host_compact == "auto":
        forced = {}
        for name, mode in (("dense_upload", J.COMPACT_OFF), ("host_compaction", J.COMPACT_ON)):
            ctx_f = J.Context(device=local_rank, arith=arith, k1_kernel=kmap[args.k1], k2_kernel=kmap[args.k2], host_compact=mode)
            f_batch = J.Batch(ctx_f, e_descs)
            for _ in range(2):
                f_batch.run_host(outs)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                f_batch.run_host(outs)
            forced[name] = {"value": Be * W * H / 1e6 * e_steps / (time.perf_counter() - t0), "unit": "MP/s"}
            f_batch.close()
            ctx_f.close()
    # cheap end-to-end sanity: the host result of image 0 equals the device-resident result
    ref0 = d_out[batch.layout(0)["out_off"]:batch.layout(0)["out_off"] + out_per_img].cpu().numpy()
    same = bool(np.array_equal(ref0, h_out.numpy()[:out_per_img]))

    # ---- whole files: JPEG bytes -> pixels (host Huffman on the NUMA-local cores + GPU worker path) ----
    files_e2e = None
    if world > 1:
        # every rank decodes its own shard of files on its own GPU; the host cores the ranks share are split between them
        ok, f_dt, f_reps, Bf = 1, 0.0, 3, 256
        try:
            nthreads = max(2, len(os.sched_getaffinity(0)) // world)
            jpegs = [workload.synth_jpeg(W, H, cfg["seed"] + k, cfg["subsampling"]) for k in range(min(U, 4))]
            f_out = torch.empty(Bf * out_per_img, dtype=torch.uint8, pin_memory=True).numpy()
            fbufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]
            jobs = (J.FileJob * Bf)()
            for j in range(Bf):
                jobs[j].data, jobs[j].len = fbufs[j % len(jpegs)].ctypes.data, fbufs[j % len(jpegs)].size
                jobs[j].out, jobs[j].out_cap = f_out[j * out_per_img:].ctypes.data, out_per_img
            ctx.check(J.lib().b200jpg_decode_files(ctx._h, jobs, Bf, nthreads))   # warm-up
        except Exception as e:   # a rank that fails still takes part in the collectives below
            ok = 0
            print("rank %d: files section failed: %r" % (rank, e), file=sys.stderr)
        dist.barrier()
        t0 = time.perf_counter()
        try:
            if ok:
                for _ in range(f_reps):
                    ctx.check(J.lib().b200jpg_decode_files(ctx._h, jobs, Bf, nthreads))
                ok = int(all(jobs[j].status == 0 for j in range(Bf)))
        except Exception as e:
            ok = 0
            print("rank %d: files section failed: %r" % (rank, e), file=sys.stderr)
        f_dt = time.perf_counter() - t0
        t = torch.tensor([f_dt, -float(ok)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank; -ok: any failure wins
        if rank == 0:
            if t[1].item() == -1.0:
                files_e2e = {"value": world * Bf * W * H / 1e6 * f_reps / float(t[0].item()), "unit": "MP/s", "images": world * Bf,
                             "host_threads_per_rank": nthreads,
                             "api": "b200jpg_decode_files on every rank's shard (JPEG bytes -> pinned host pixels, Huffman decoding on the GPUs)"}
            else:
                files_e2e = {"error": "a rank failed, see stderr"}
    if rank == 0 and world == 1:
        # host threads = the CPUs of the GPU's NUMA node (this process is bound to them): measured faster than using
        # both sockets (profiles/r01_files_trace.txt)
        nthreads = max(1, len(os.sched_getaffinity(0)))
        Bf = min(2048, max(256, 16 * nthreads))
        jpegs = [workload.synth_jpeg(W, H, cfg["seed"] + k, cfg["subsampling"]) for k in range(min(U, 4))]
        flist = [jpegs[j % len(jpegs)] for j in range(Bf)]
        f_out = torch.empty(Bf * out_per_img, dtype=torch.uint8, pin_memory=True).numpy()
        f_outs = [f_out[j * out_per_img:(j + 1) * out_per_img] for j in range(Bf)]
        # the job array is built outside the timed region; the timed call is the C entry point itself
        import ctypes as C
        fbufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]
        jobs = (J.FileJob * Bf)()
        for j in range(Bf):
            jobs[j].data, jobs[j].len = fbufs[j % len(jpegs)].ctypes.data, fbufs[j % len(jpegs)].size
            jobs[j].out, jobs[j].out_cap = f_outs[j].ctypes.data, f_outs[j].size
        def time_files(c):
            c.check(J.lib().b200jpg_decode_files(c._h, jobs, Bf, nthreads))   # warm-up (allocates the cached arenas)
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                c.check(J.lib().b200jpg_decode_files(c._h, jobs, Bf, nthreads))
            dt = (time.perf_counter() - t0) / reps
            assert all(jobs[j].status == 0 for j in range(Bf))
            assert bool(np.array_equal(f_outs[0], ref0))
            return Bf * W * H / 1e6 / dt, reps
        # Huffman decoding on the device (the default for complete baseline scans), then forced onto the host threads
        f_value, f_reps = time_files(ctx)
        scans = ctx.device_scan_counts
        f_out[:] = 0
        ctx_h = J.Context(device=local_rank, arith=arith, k1_kernel=kmap[args.k1], k2_kernel=kmap[args.k2], entropy=J.ENTROPY_HOST)
        fh_value, _ = time_files(ctx_h)

        def decode_latency_ms(c, reps=20):   # Decoder::new + decode() + drop of one 1080p file through the C ABI itself
            L = J.lib()
            src = fbufs[0]

            def once():
                h, px, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
                c.check(L.b200jpg_decoder_new(c._h, src.ctypes.data, src.size, C.byref(h)))
                rc = L.b200jpg_decoder_decode(h, C.byref(px), C.byref(n))
                first = bytes((C.c_uint8 * 64).from_address(px.value)) if rc == 0 else b""
                L.b200jpg_decoder_free(h)
                return rc, n.value, first
            assert once() == (0, out_per_img, bytes(ref0[:64]))
            t0 = time.perf_counter()
            for _ in range(reps):
                once()
            return 1e3 * (time.perf_counter() - t0) / reps
        single = {"device_entropy_ms": decode_latency_ms(ctx), "host_entropy_ms": decode_latency_ms(ctx_h),
                  "api": "b200jpg_decoder_new + b200jpg_decoder_decode (Decoder::decode) on one %dx%d file, latency per call" % (W, H)}
        ctx_h.close()
        # the same call with the pixel buffers in DEVICE memory (an on-GPU consumer, e.g. a training input pipeline): no D2H
        d_pix = torch.empty(Bf * out_per_img, dtype=torch.uint8, device=dev)
        for j in range(Bf):
            jobs[j].out = d_pix.data_ptr() + j * out_per_img
        for _ in range(4):   # warm-up: groups are larger in this mode, the engine's device buffers grow (by doubling) to fit them
            ctx.check(J.lib().b200jpg_decode_files(ctx._h, jobs, Bf, nthreads))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.check(J.lib().b200jpg_decode_files(ctx._h, jobs, Bf, nthreads))
        torch.cuda.synchronize()
        fd_value = Bf * W * H / 1e6 * 5 / (time.perf_counter() - t0)
        assert all(jobs[j].status == 0 for j in range(Bf))
        assert bool(np.array_equal(d_pix[:out_per_img].cpu().numpy(), ref0))
        for j in range(Bf):
            jobs[j].out = f_outs[j].ctypes.data
        del d_pix
        files_e2e = {"value": f_value, "unit": "MP/s", "images": Bf, "host_threads": nthreads,
                     "jpeg_bytes_per_image": int(np.mean([len(j) for j in jpegs])),
                     "calls": f_reps, "scans_decoded_on_device": int(scans[0]), "scans_handed_back_to_host": int(scans[1]),
                     "device_outputs": {"value": fd_value, "unit": "MP/s",
                                        "api": "same call, b200jpg_file_job.out in device memory: JPEG bytes over PCIe, pixels stay in HBM"},
                     "single_image": single,
                     "host_entropy": {"value": fh_value, "unit": "MP/s",
                                      "api": "same call with B200JPG_ENTROPY_HOST: Huffman on the host threads -> sparse block streams -> K0/K1/K2"},
                     "api": "b200jpg_decode_files (JPEG bytes -> pinned host pixels; host threads parse markers and copy the scan, "
                            "Huffman decoding + K1 + K2 on the GPU)"}
    pcie = pcie_probe(torch, dev) if rank == 0 else None
    os.sched_setaffinity(0, all_cpus)   # the CPU baseline uses every core
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample, cpu_img = cpu_port_throughput(unique, cores, args.cpu_seconds, args.arith)
        fv, fsample = cpu_files_throughput(jpegs, W, H, cores, 4.0, args.arith)
        cpu_baseline = {"value": v, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample,
                        "gpu_matches_cpu_bit_exact": bool(np.array_equal(cpu_img, ref0)),
                        "files": {"value": fv, "unit": "MP/s", "sample": fsample}}

    if rank == 0:
        line = {
            "metric": "megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "batch_per_gpu": B, "global_batch": B * world, "width": W, "height": H,
                       "unique_images": U, "arith": args.arith, "k1": args.k1, "k2": args.k2, "parallelism": "images sharded by index, %d rank(s)" % world,
                       "l2": "inputs (%.1f GB/GPU) far exceed the 126 MB L2; no flush needed" % (info.coef_bytes / 1e9)},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": Be * coef_per_img, "d2h_bytes_per_step": Be * out_per_img,
                    "images_per_step": Be, "steps": e_steps, "host_result_equals_device_result": same,
                    "api": "b200jpg_batch_run_host (pinned host dense coefficient buffers -> pinned host pixels)",
                    "host_compact": args.host_compact, "forced": forced,
                    "note": "h2d_bytes_per_step = the dense input the call is given; with host compaction the link carries the sparse streams",
                    "pcie_gbs_measured": pcie, "numa_bind": numa, "files": files_e2e,
                    "d2h_gbs": Be * out_per_img * e_steps / e_dt / 1e9},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    e_batch.close()
    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
