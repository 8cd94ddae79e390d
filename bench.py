#!/usr/bin/env python
"""bench.py -- megapixels/s of the JPEG block pipeline hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2]        # this repo's CUDA path
  python bench.py --impl reference ...                                        # CPU reference arm

One "step" = one pass of the hot path (dequantise + IDCT -> upsample + colour -> pixels) over a batch of synthetic
images whose dense coefficients are already resident in HBM.  Prints ONE JSON line (rank 0).  N > 1 is launched by
torchrun, one rank per GPU; images are sharded by index (weak scaling: `batch` images per GPU), no data-path
collective.

  value      device-timed MP/s of the step (one fused kernel, or K1 + K2 where it does not apply)
  roofline   the step's dominant kernel against the measured HBM peak; the two-kernel route beside it
  e2e        the same metric through the call a user of the reference makes -- JPEG files in host memory ->
             RGB pixels in (pinned) host memory, b200jpg_decode_files = an outer par_iter over Decoder::decode() --
             copies inside the timed region; the dense-coefficient worker boundary (b200jpg_batch_run_host) is
             reported beside it as e2e.worker_boundary
  configs    BASELINE.json configs 3 and 4 (4K 4:4:4 x1024, tower_progressive.jpg x512), each checked against the oracle

`--impl reference`: the reference is a Rust crate and cannot be built in this image (no cargo / rustc), so the
reference arm times the C restatement of the reference's CPU path (oracle/, kind "port") on all host cores in the
arithmetic an x86-64 build of the reference executes (SSSE3 intrinsics, src/arch/ssse3.rs), on the same workload;
`value` = dense coefficients -> RGB (the hot path), `e2e` = whole files -> RGB.  That arm never loads the product
library: its inputs come from the oracle's own decoder.
"""
import argparse
import ctypes as C
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
    ap.add_argument("--e2e-batch", type=int, default=256, help="images per worker-boundary (host->host) step")
    ap.add_argument("--arith", default="scalar", choices=["scalar", "ssse3"], help="GPU arithmetic variant (both are the reference's)")
    ap.add_argument("--cpu-arith", default="ssse3_native", choices=["ssse3_native", "scalar", "ssse3"],
                    help="arithmetic of the CPU arm: ssse3_native = what an x86-64 build of the reference runs")
    ap.add_argument("--k1", default="auto", choices=["auto", "generic"], help="K1 kernel variant (profiling)")
    ap.add_argument("--k2", default="auto", choices=["auto", "generic"], help="K2 kernel variant (profiling)")
    ap.add_argument("--fuse", default="auto", choices=["auto", "on", "off"],
                    help="route of the step: auto = the library's choice (fused kernel for 4:4:4, K1 + K2 for 4:2:0), on / off = force")
    ap.add_argument("--host-compact", default="auto", choices=["auto", "on", "off"],
                    help="b200jpg_batch_run_host: compact the dense coefficients into sparse block streams on host threads")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip cfg3 / cfg4 (they run at N=1 on the default config only)")
    ap.add_argument("--cfg3-batch", type=int, default=0, help="images of the cfg3 measurement (default: its 1024)")
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


def bind_cpus(torch, index, rank, world):
    """Pins this process to the CPUs of the GPU's NUMA node (pinned buffers are then allocated next to the GPU's PCIe
    root) and, when several ranks share the host, to this rank's own slice of them: N ranks x all-CPU thread pools on
    the same cores is what flattened the round-1 end-to-end curve."""
    info = {"node": None, "cpus": None, "slice": None}
    try:
        cpus = sorted(os.sched_getaffinity(0))
        try:
            pr = torch.cuda.get_device_properties(index)
            bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
            if node >= 0:
                ncpus = set()
                for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                    a, _, b = part.partition("-")
                    ncpus.update(range(int(a), int(b or a) + 1))
                local = sorted(ncpus & set(cpus))
                if local:
                    info["node"] = node
                    # ranks whose GPUs hang off the same node share that node's CPUs
                    same = [r for r in range(world) if _gpu_node(torch, r) == node] if world > 1 else [rank]
                    k = same.index(rank) if rank in same else 0
                    per = max(1, len(local) // max(1, len(same)))
                    mine = local[k * per:(k + 1) * per] or local
                    os.sched_setaffinity(0, mine)
                    info["cpus"], info["slice"] = len(mine), "node %d cpus %d-%d" % (node, mine[0], mine[-1])
                    return info
        except Exception:
            pass
        if world > 1:   # no NUMA information (single-node VM): a disjoint slice per rank
            per = max(1, len(cpus) // world)
            mine = cpus[rank * per:(rank + 1) * per] or cpus
            os.sched_setaffinity(0, mine)
            info["cpus"], info["slice"] = len(mine), "cpus %d-%d" % (mine[0], mine[-1])
    except Exception:
        pass
    return info


def _gpu_node(torch, index):
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception:
        return -1


def pcie_probe(torch, dev, dist=None, nbytes=1 << 29):
    """Measured pinned-memory copy bandwidth (GB/s per GPU): the roofline of the end-to-end numbers.  With several
    ranks every rank copies AT THE SAME TIME (barrier first), so the sum is what the box's host side sustains."""
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return 4 * nbytes / (time.perf_counter() - t0) / 1e9
    run(True, True)
    return {"h2d": run(True, False), "d2h": run(False, True), "bidir_each": run(True, True)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's CPU path), never the product library
# ---------------------------------------------------------------------------------------------
def cpu_arith(name):
    import oracle
    return {"ssse3_native": oracle.ARITH_SSSE3_NATIVE, "ssse3": oracle.ARITH_SSSE3, "scalar": oracle.ARITH_SCALAR}[name]


class OracleImage:
    """One image entropy-decoded by the ORACLE's decoder: geometry + dense coefficients."""

    def __init__(self, jpeg):
        import oracle
        d = oracle.Decoder(jpeg)
        px = d.decode()
        inf = d.info()
        self.width, self.height = inf.width, inf.height
        self.components, self.qts = d.components()
        self.ncomp = len(self.components)
        self.coefs = [d.coefficients(i) for i in range(self.ncomp)]
        self.color_transform = d.color_transform()
        self.jpeg = jpeg
        self.pixels_scalar = px


def cpu_hotpath_throughput(imgs, nthreads, target_seconds, arith):
    """oracle.hotpath_batch (start + append_row* + get_result + compute_image per image) on a bounded sample,
    one image per thread at a time: what an outer par_iter over the reference's decoder gives."""
    import oracle
    u0 = imgs[0]
    n = max(nthreads * 2, 2)
    coefs = [imgs[i % len(imgs)].coefs for i in range(n)]
    outs = [np.zeros(u0.width * u0.height * u0.ncomp, dtype=np.uint8) for _ in range(n)]
    a = cpu_arith(arith)
    t0 = time.perf_counter()
    oracle.hotpath_batch(u0.components, u0.qts, coefs, u0.width, u0.height, u0.color_transform, outs, nthreads, a)
    pilot = time.perf_counter() - t0
    reps = max(1, min(200, int(target_seconds / max(pilot, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.hotpath_batch(u0.components, u0.qts, coefs, u0.width, u0.height, u0.color_transform, outs, nthreads, a)
    dt = time.perf_counter() - t0
    mp = reps * n * u0.width * u0.height / 1e6
    return mp / dt, "%d images x %d passes (%.1f s), dense coefficients -> RGB, one image per thread, %s arithmetic" % (n, reps, dt, arith)


def cpu_files_throughput(jpegs, width, height, nthreads, target_seconds, arith):
    """Whole files on the CPU: the oracle's restatement of Decoder::decode() (marker parsing, Huffman, IDCT, upsampling,
    colour), one image per host thread -- an outer par_iter over the reference's decoder."""
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    L = oracle.lib()
    a = cpu_arith(arith)
    bufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]

    def one(k):   # ctypes drops the GIL for the duration of each call
        b = bufs[k % len(bufs)]
        d = L.orc_decoder_new(b.ctypes.data_as(C.POINTER(C.c_uint8)), b.size, a)
        L.orc_decoder_set_taps(d, 0)
        p, n = C.c_void_p(), C.c_size_t()
        rc = L.orc_decoder_decode(d, C.byref(p), C.byref(n))
        L.orc_decoder_free(d)
        return rc
    with ThreadPoolExecutor(nthreads) as ex:
        t0 = time.perf_counter()
        assert all(r == 0 for r in ex.map(one, range(nthreads)))
        pilot = time.perf_counter() - t0
        n = nthreads * max(1, min(512, int(target_seconds / max(pilot, 1e-3))))
        t0 = time.perf_counter()
        assert all(r == 0 for r in ex.map(one, range(n)))
        dt = time.perf_counter() - t0
    return n * width * height / 1e6 / dt, "%d files (%.1f s), JPEG bytes -> RGB, one image per thread, %s arithmetic" % (n, dt, arith)


def cpu_single_image_latency(img, nthreads, arith, reps=8):
    """one image the way the reference's rayon build decodes it -- Huffman + inline IDCT on the calling
    thread (src/worker/rayon.rs:121-132), then one task per output row over the pool (src/worker/rayon.rs:204-216)."""
    import oracle
    a = cpu_arith(arith)
    out = {}
    # hot path only (dense coefficients -> RGB)
    oracle.hotpath_image(img.components, img.qts, img.coefs, img.width, img.height, img.color_transform, a, nthreads)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.hotpath_image(img.components, img.qts, img.coefs, img.width, img.height, img.color_transform, a, nthreads)
    out["hot_path_ms"] = 1e3 * (time.perf_counter() - t0) / reps
    # whole file
    def once():
        d = oracle.Decoder(img.jpeg, a)
        d.set_taps(False)
        d.set_threads(nthreads)
        d.decode()
    once()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    out["whole_file_ms"] = 1e3 * (time.perf_counter() - t0) / reps
    out["threads"] = nthreads
    out["shape"] = "IDCT inline on one thread, upsample + colour rows over %d threads (rayon build)" % nthreads
    return out


def cpu_baseline_block(imgs, jpegs, width, height, cores, seconds, arith, with_scalar=True):
    """The cpu_baseline object: hot path + whole files, all cores, in the x86 arithmetic (and scalar beside it)."""
    v, sample = cpu_hotpath_throughput(imgs, cores, seconds, arith)
    fv, fsample = cpu_files_throughput(jpegs, width, height, cores, max(3.0, seconds / 3), arith)
    block = {"value": v, "unit": "MP/s", "cores": cores, "kind": "port", "arith": arith, "sample": sample,
             "files": {"value": fv, "unit": "MP/s", "sample": fsample},
             "single_image": cpu_single_image_latency(imgs[0], cores, arith),
             "note": "C restatement of the reference's CPU path (oracle/); %s = the arithmetic of src/arch/ssse3.rs through "
                     "<tmmintrin.h>, what an x86-64 build of the reference executes" % arith}
    if with_scalar and arith != "scalar":
        sv, ssample = cpu_hotpath_throughput(imgs, cores, max(2.0, seconds / 4), "scalar")
        block["scalar"] = {"value": sv, "unit": "MP/s", "sample": ssample,
                           "note": "the reference's platform_independent i32 path (src/idct.rs:260-370), the arithmetic the GPU arm computes in"}
    return block


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from jpeg_decoder_b200 import workload   # pure Python here: generators and config table, the library is never loaded
    cfg = workload.CONFIGS[args.config]
    jpegs = [workload.config_jpeg(args.config, k) for k in range(min(args.unique, 4))]
    imgs = [OracleImage(j) for j in jpegs]
    # the native intrinsics and the lane-by-lane emulation are the same function: checked on the bench images
    u0 = imgs[0]
    a_nat = oracle.hotpath_image(u0.components, u0.qts, u0.coefs, u0.width, u0.height, u0.color_transform, oracle.ARITH_SSSE3_NATIVE)
    a_emu = oracle.hotpath_image(u0.components, u0.qts, u0.coefs, u0.width, u0.height, u0.color_transform, oracle.ARITH_SSSE3)
    native_ok = bool(np.array_equal(a_nat, a_emu))
    cores = os.cpu_count() or 1
    vals = []
    sample = ""
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_hotpath_throughput(imgs, cores, per_step / 2, args.cpu_arith)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sample = cpu_hotpath_throughput(imgs, cores, per_step, args.cpu_arith)
        vals.append(v)
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    fv, fsample = cpu_files_throughput(jpegs, cfg["width"], cfg["height"], cores, 6.0, args.cpu_arith)
    sv, ssample = cpu_hotpath_throughput(imgs, cores, 3.0, "scalar")
    single = cpu_single_image_latency(imgs[0], cores, args.cpu_arith)
    line = {
        "impl": "reference", "metric": "megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16" if args.cpu_arith != "scalar" else "i32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "width": cfg["width"], "height": cfg["height"], "arith": args.cpu_arith,
                   "note": "reference is Rust (no toolchain here): C restatement of its CPU hot path in the x86 build's SSSE3 arithmetic "
                           "(real intrinsics), one image per host thread; inputs decoded by the oracle's own decoder"},
        "cpu_baseline": {"value": value, "unit": "MP/s", "cores": cores, "kind": "port", "arith": args.cpu_arith, "sample": "per step: " + sample,
                         "ssse3_native_equals_emulation": native_ok,
                         "scalar": {"value": sv, "unit": "MP/s", "sample": ssample},
                         "files": {"value": fv, "unit": "MP/s", "sample": fsample}, "single_image": single},
        "e2e": {"value": fv, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "whole files: JPEG bytes -> RGB (oracle's Decoder::decode, one image per thread); the hot-path-only number is `value`",
                "hot_path": {"value": value, "unit": "MP/s"}},
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


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class DeviceWorkload:
    """A batch of one config resident in HBM: plan + slabs, unique images uploaded once and replicated on the device."""

    def __init__(self, J, torch, ctx, dev, unique, lo, B, need_planes=True):
        self.J, self.torch, self.ctx, self.unique, self.lo, self.B = J, torch, ctx, unique, lo, B
        U = len(unique)
        keep, descs = [], []
        for i in range(lo, lo + B):
            u = unique[i % U]
            descs.append(J.make_image_desc(u.width, u.height, u.components, u.qts, u.coefs, u.color_transform, keep))
        self.keep = keep
        self.batch = J.Batch(ctx, descs)
        info = self.info = self.batch.info
        self.d_coefs = torch.empty(info.coef_bytes, dtype=torch.uint8, device=dev)
        self.d_planes = torch.empty(info.plane_bytes if need_planes else 256, dtype=torch.uint8, device=dev)
        self.d_out = torch.empty(info.out_bytes, dtype=torch.uint8, device=dev)
        dev_unique = [[torch.from_numpy(c.view(np.uint8)).to(dev) for c in u.coefs] for u in unique]
        for j in range(B):
            lay = self.batch.layout(j)
            for k, src in enumerate(dev_unique[(lo + j) % U]):
                self.d_coefs[lay["coef_off"][k]:lay["coef_off"][k] + src.numel()].copy_(src)
        torch.cuda.synchronize()
        self.fusable = info.n_fused == B

    def run(self, stages=3):
        self.batch.run_device(self.d_coefs.data_ptr(), self.d_planes.data_ptr(), self.d_out.data_ptr(), stages)

    def timed(self, stream, stages, steps, sampler=None):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            self.run(stages)
        e1.record(stream)
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.windows.append((t0, time.perf_counter()))
        return e0.elapsed_time(e1)

    def pixels(self, j):
        lay = self.batch.layout(j)
        return self.d_out[lay["out_off"]:lay["out_off"] + lay["out_len"]].cpu().numpy()

    def check_against_oracle(self, max_images=8):
        """The device result of the first image built from EACH distinct input equals the oracle's hot path (scalar
        arithmetic: bit-exact) -- and every replica of it carries the same checksum."""
        import oracle
        J, U = self.J, len(self.unique)
        checked, ok = 0, True
        sums = self.torch.stack([self.d_out[self.batch.layout(j)["out_off"]:self.batch.layout(j)["out_off"] + self.batch.layout(j)["out_len"]]
                                 .to(self.torch.int64).sum() for j in range(min(self.B, 4 * U))]).cpu().numpy()
        for k in range(min(U, max_images, self.B)):
            j = next(jj for jj in range(self.B) if (self.lo + jj) % U == (self.lo + k) % U)
            u = self.unique[(self.lo + j) % U]
            oc = (oracle.Component * u.ncomp)()
            for i, c in enumerate(u.components):
                for f, _ in oracle.Component._fields_:
                    setattr(oc[i], f, getattr(c, f))
            a = oracle.ARITH_SSSE3 if self.ctx.arith == J.ARITH_SSSE3 else oracle.ARITH_SCALAR
            want = oracle.hotpath_image(oc, u.qts, u.coefs, u.width, u.height, u.color_transform, a)
            ok = ok and bool(np.array_equal(self.pixels(j), want))
            checked += 1
        replicas_ok = all(int(sums[j]) == int(sums[j % U]) for j in range(len(sums)))
        return {"images_checked_vs_oracle": checked, "bit_exact": ok, "replica_checksums_agree": bool(replicas_ok)}

    def close(self):
        self.batch.close()
        del self.d_coefs, self.d_planes, self.d_out


def measure_config(J, torch, ctx, dev, stream, unique, lo, B, steps, peak, sampler=None, dist=None, world=1):
    """Device-timed hot path of one resident batch: the step (fused where it applies), K1 and K2 alone, K1 + K2."""
    need_planes = True
    wl = DeviceWorkload(J, torch, ctx, dev, unique, lo, B, need_planes)
    info = wl.info
    for _ in range(3):
        wl.run()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches0 = ctx.launch_count
    ms = wl.timed(stream, 3, steps, sampler)
    launches = ctx.launch_count - launches0
    fused = launches == steps          # one launch per step = the fused kernel, two = K1 then K2
    parity = wl.check_against_oracle()
    ms_max = ms
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        dist.barrier()
    # both routes on the same buffers: K1 and K2 alone and back to back, and the fused kernel where it applies
    ctx.set_fuse(J.FUSE_OFF)
    for _ in range(2):
        wl.run()
    ms_k1 = wl.timed(stream, 1, steps, sampler) / steps
    ms_k2 = wl.timed(stream, 2, steps, sampler) / steps
    ms_k12 = wl.timed(stream, 3, steps, sampler) / steps
    same_two_kernel = wl.check_against_oracle(2)["bit_exact"]
    ms_kf, same_kf = None, None
    if wl.fusable:
        ctx.set_fuse(J.FUSE_ON)
        wl.d_out.zero_()
        for _ in range(2):
            wl.run()
        ms_kf = wl.timed(stream, 3, steps, sampler) / steps
        same_kf = wl.check_against_oracle(2)["bit_exact"]
    ctx.set_fuse(ctx.fuse)
    k1_gbs = info.k1_algorithmic_bytes / (ms_k1 * 1e-3) / 1e9
    k2_gbs = info.k2_algorithmic_bytes / (ms_k2 * 1e-3) / 1e9
    k12_bytes = info.k1_algorithmic_bytes + info.k2_algorithmic_bytes
    kernels = {
        "k1_dequant_idct8x8": {"ms": ms_k1, "algorithmic_bytes": info.k1_algorithmic_bytes, "achieved": k1_gbs, "frac": k1_gbs / peak},
        "k2_upsample_color": {"ms": ms_k2, "algorithmic_bytes": info.k2_algorithmic_bytes, "achieved": k2_gbs, "frac": k2_gbs / peak},
        "k1_then_k2": {"ms": ms_k12, "algorithmic_bytes": k12_bytes, "achieved": k12_bytes / (ms_k12 * 1e-3) / 1e9,
                       "frac": k12_bytes / (ms_k12 * 1e-3) / 1e9 / peak, "bit_exact_vs_oracle": same_two_kernel},
    }
    if ms_kf is not None:
        kf_gbs = info.kf_algorithmic_bytes / (ms_kf * 1e-3) / 1e9
        kernels["kf_fused"] = {"ms": ms_kf, "algorithmic_bytes": info.kf_algorithmic_bytes, "achieved": kf_gbs, "frac": kf_gbs / peak,
                               "bit_exact_vs_oracle": same_kf,
                               "note": "planes staged in shared memory: coefficients in + pixels out are the only HBM bytes"}
    if fused:
        dominant, dom_gbs = "kf_fused", kernels["kf_fused"]["achieved"]
    else:
        dominant = "k1_dequant_idct8x8" if ms_k1 >= ms_k2 else "k2_upsample_color"
        dom_gbs = k1_gbs if ms_k1 >= ms_k2 else k2_gbs
    return {"wl": wl, "ms": ms, "ms_max": ms_max, "launches": launches, "parity": parity, "kernels": kernels, "dominant": dominant,
            "dom_gbs": dom_gbs, "fused": fused, "info": info}


def roofline_of(m, peak, peak_src, cfgname, B):
    traffic = None
    try:
        # DRAM bytes per image of the dominant kernel from the committed `ncu --set full` capture (scripts/ncu_traffic.py),
        # scaled to this launch's batch: per launch, like `achieved`
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        per_image = tr.get(cfgname, {}).get(m["dominant"])
        traffic = per_image * B if per_image is not None else None
    except Exception:
        pass
    return {"bound": "hbm", "kernel": m["dominant"], "achieved": m["dom_gbs"], "peak": peak, "unit": "GB/s", "frac": m["dom_gbs"] / peak,
            "traffic": traffic, "peak_source": peak_src, "step_route": "kf_fused" if m["fused"] else "k1_then_k2", "kernels": m["kernels"],
            "note": "both routes are timed on the same buffers; the library picks per sampling mode (fused for 4:4:4, K1 then K2 for 4:2:0). "
                    "The fused kernel moves fewer bytes (no plane write + read), so its byte fraction reads lower at equal MP/s: all kernels "
                    "are bound by integer issue, not by HBM (DESIGN.md section 4)"}


def files_jobs(J, jpegs, n, out_base, out_per_img):
    fbufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]
    jobs = (J.FileJob * n)()
    for j in range(n):
        jobs[j].data, jobs[j].len = fbufs[j % len(jpegs)].ctypes.data, fbufs[j % len(jpegs)].size
        jobs[j].out, jobs[j].out_cap = out_base + j * out_per_img, out_per_img
    return jobs, fbufs


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
    D = dist if world > 1 else None

    cfg = workload.CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    W, H = cfg["width"], cfg["height"]
    all_cpus = os.sched_getaffinity(0)
    numa = None if args.no_numa_bind else bind_cpus(torch, local_rank, rank, world)
    my_cpus = len(os.sched_getaffinity(0))
    # control plane: rank 0 broadcasts the image -> GPU assignment (contiguous index ranges)
    table = workload.broadcast_assignment(B * world, world, D, device=dev)
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
    fuse = {"auto": J.FUSE_AUTO, "on": J.FUSE_ON, "off": J.FUSE_OFF}[args.fuse]
    # several ranks share the host: every rank's thread pools stay inside its own CPU slice
    host_threads = my_cpus if world > 1 else 0
    ctx = J.Context(device=local_rank, arith=arith, k1_kernel=kmap[args.k1], k2_kernel=kmap[args.k2], stream=stream.cuda_stream,
                    host_compact=cmap[args.host_compact], host_threads=host_threads, fuse=fuse)
    peak, peak_src = load_peaks()

    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure_config(J, torch, ctx, dev, stream, unique, lo, B, args.steps, peak, sampler, D, world)
    sampler.stop()
    clocks = sampler.summary()
    wl, info = m["wl"], m["info"]
    mp_per_step = world * B * W * H / 1e6
    value = mp_per_step * args.steps / (m["ms_max"] / 1e3)
    roofline = roofline_of(m, peak, peak_src, args.config, B)
    u0 = unique[0]
    out_per_img = W * H * u0.ncomp
    ref0 = wl.pixels(0)

    # ---- worker boundary: dense coefficients in pinned host memory -> pinned host pixels (b200jpg_batch_run_host) ----
    Be = min(args.e2e_batch, B)
    coef_per_img = u0.coef_bytes
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
    wb_dt = time.perf_counter() - t0
    assert all(s == 0 for s in st)
    if world > 1:
        t = torch.tensor([wb_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wb_dt = float(t.item())
    wb_value = world * Be * W * H / 1e6 * e_steps / wb_dt
    same = bool(np.array_equal(ref0, h_out.numpy()[:out_per_img]))
    forced = None
    if rank == 0 and world == 1 and args.host_compact == "auto":
        forced = {}
        for name, mode in (("dense_upload", J.COMPACT_OFF), ("host_compaction", J.COMPACT_ON)):
            ctx_f = J.Context(device=local_rank, arith=arith, k1_kernel=kmap[args.k1], k2_kernel=kmap[args.k2], host_compact=mode, fuse=fuse)
            f_batch = J.Batch(ctx_f, e_descs)
            for _ in range(2):
                f_batch.run_host(outs)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                f_batch.run_host(outs)
            forced[name] = {"value": Be * W * H / 1e6 * e_steps / (time.perf_counter() - t0), "unit": "MP/s"}
            f_batch.close()
            ctx_f.close()
    worker_boundary = {"value": wb_value, "unit": "MP/s", "h2d_bytes_per_step": world * Be * coef_per_img, "d2h_bytes_per_step": world * Be * out_per_img,
                       "images_per_step": world * Be, "steps": e_steps, "host_result_equals_device_result": same,
                       "api": "b200jpg_batch_run_host (pinned host dense coefficient buffers -> pinned host pixels): the reference's trait Worker boundary, "
                              "3 B/px in + 3 B/px out over PCIe by construction",
                       "host_compact": args.host_compact, "forced": forced, "d2h_gbs_per_gpu": Be * out_per_img * e_steps / wb_dt / 1e9}
    e_batch.close()
    del h_in, h_in_np

    # ---- e2e: whole files, JPEG bytes in host memory -> RGB in pinned host memory (b200jpg_decode_files) ----
    nthreads = max(1, my_cpus)
    # 512 files per call and rank at every N (weak scaling: the per-GPU work does not change); a box that cannot page-lock
    # 3.2 GB per rank falls back to 256 on every rank together
    Bf = 512
    jpegs = [u.jpeg for u in unique[:4]]
    jpeg_bytes = int(np.mean([len(j) for j in jpegs]))
    try:
        f_out = torch.empty(Bf * out_per_img, dtype=torch.uint8, pin_memory=True)
        got = 1
    except RuntimeError:
        f_out, got = None, 0
    if world > 1:
        t = torch.tensor([got], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        got = int(t.item())
    if not got:
        f_out = None
        Bf = 256
        f_out = torch.empty(Bf * out_per_img, dtype=torch.uint8, pin_memory=True)
    f_np = f_out.numpy()
    jobs, fbufs = files_jobs(J, jpegs, Bf, f_out.data_ptr(), out_per_img)
    f_reps = 5

    def time_files(c, reps, sync=False):
        ok = 1
        try:
            c.check(J.lib().b200jpg_decode_files(c._h, jobs, Bf, nthreads))   # warm-up (allocates the cached arenas)
        except Exception as e:   # a rank that fails still takes part in the collectives below
            ok = 0
            print("rank %d: files section failed: %r" % (rank, e), file=sys.stderr)
        if sync:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        try:
            if ok:
                for _ in range(reps):
                    c.check(J.lib().b200jpg_decode_files(c._h, jobs, Bf, nthreads))
                if sync:
                    torch.cuda.synchronize()
                ok = int(all(jobs[j].status == 0 for j in range(Bf)))
        except Exception as e:
            ok = 0
            print("rank %d: files section failed: %r" % (rank, e), file=sys.stderr)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, -float(ok)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank; -ok: any failure wins
            dt, ok = float(t[0].item()), int(t[1].item() == -1.0)
        return (world * Bf * W * H / 1e6 * reps / dt) if ok else None

    f_value = time_files(ctx, f_reps)
    files_same = bool(np.array_equal(f_np[:out_per_img], ref0)) if args.arith == "scalar" else None
    scans = ctx.device_scan_counts
    # the same call with the pixel buffers in DEVICE memory (an on-GPU consumer, e.g. a training input pipeline): no D2H
    d_pix = torch.empty(Bf * out_per_img, dtype=torch.uint8, device=dev)
    for j in range(Bf):
        jobs[j].out = d_pix.data_ptr() + j * out_per_img
    for _ in range(3):   # warm-up: groups are larger in this mode, the engine's device buffers grow (by doubling) to fit them
        J.lib().b200jpg_decode_files(ctx._h, jobs, Bf, nthreads)
    fd_value = time_files(ctx, 12, sync=True)   # 12 calls of ~15 ms: a single slow call (a group that just missed a slot) moves a mean of 5 by 5 %
    dev_same = bool(np.array_equal(d_pix[:out_per_img].cpu().numpy(), ref0)) if args.arith == "scalar" else None
    for j in range(Bf):
        jobs[j].out = f_out.data_ptr() + j * out_per_img
    del d_pix
    files_e2e = {"images_per_step": world * Bf, "host_threads_per_rank": nthreads, "jpeg_bytes_per_image": jpeg_bytes, "calls": f_reps,
                 "scans_decoded_on_device": int(scans[0]), "scans_handed_back_to_host": int(scans[1]),
                 "host_pixels_equal_device_timed_result": files_same,
                 "device_outputs": {"value": fd_value, "unit": "MP/s", "pixels_equal": dev_same,
                                    "api": "same call, b200jpg_file_job.out in device memory: JPEG bytes over PCIe, pixels stay in HBM"}}
    if rank == 0 and world == 1:
        ctx_h = J.Context(device=local_rank, arith=arith, k1_kernel=kmap[args.k1], k2_kernel=kmap[args.k2], entropy=J.ENTROPY_HOST, fuse=fuse)
        fh_value = time_files(ctx_h, f_reps)

        def decode_latency_ms(c, reps=20):   # Decoder::new + decode() + drop of one file through the C ABI itself
            L = J.lib()
            src = fbufs[0]

            def once():
                h, px, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
                c.check(L.b200jpg_decoder_new(c._h, src.ctypes.data, src.size, C.byref(h)))
                rc = L.b200jpg_decoder_decode(h, C.byref(px), C.byref(n))
                first = bytes((C.c_uint8 * 64).from_address(px.value)) if rc == 0 else b""
                L.b200jpg_decoder_free(h)
                return rc, n.value, first
            assert once()[:2] == (0, out_per_img)
            t0 = time.perf_counter()
            for _ in range(reps):
                once()
            return 1e3 * (time.perf_counter() - t0) / reps
        files_e2e["single_image"] = {"device_entropy_ms": decode_latency_ms(ctx), "host_entropy_ms": decode_latency_ms(ctx_h),
                                     "api": "b200jpg_decoder_new + b200jpg_decoder_decode (Decoder::decode) on one %dx%d file, latency per call" % (W, H)}
        files_e2e["host_entropy"] = {"value": fh_value, "unit": "MP/s",
                                     "api": "same call with B200JPG_ENTROPY_HOST: Huffman on the host threads -> sparse block streams -> K0 + hot path"}
        ctx_h.close()
    pcie = pcie_probe(torch, dev, D)
    if world > 1:   # what the box's host side sustains with every GPU copying at once
        t = torch.tensor([pcie["h2d"], pcie["d2h"], pcie["bidir_each"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        pcie = {"concurrent_all_ranks_sum": {"h2d": float(t[0]), "d2h": float(t[1]), "bidir_each": float(t[2])}, "rank0": pcie}
    d2h_ceiling = (pcie["concurrent_all_ranks_sum"]["d2h"] if world > 1 else pcie["d2h"])
    e2e = {"value": f_value, "unit": "MP/s", "h2d_bytes_per_step": world * Bf * jpeg_bytes, "d2h_bytes_per_step": world * Bf * out_per_img,
           "api": "b200jpg_decode_files: JPEG files in host memory -> RGB in pinned host memory (marker parsing on host threads; Huffman decoding, "
                  "dequantise + IDCT, upsampling and colour on the GPU) -- Decoder::decode() over a batch, the call a user of the reference makes",
           "d2h_gbs": (f_value or 0) * 1e6 * u0.ncomp / 1e9,
           "fraction_of_measured_d2h_bandwidth": ((f_value or 0) * 1e6 * u0.ncomp / 1e9) / d2h_ceiling if d2h_ceiling else None,
           "pcie_gbs_measured": pcie, "cpu_binding": numa, "files": files_e2e, "worker_boundary": worker_boundary}

    # ---- BASELINE.json configs 3 and 4, and the CPU baseline: rank 0 of a single-GPU run of the default config ----
    configs = None
    cpu_baseline = None
    os.sched_setaffinity(0, all_cpus)   # the CPU baseline uses every core
    if rank == 0 and world == 1:
        wl.close()
        torch.cuda.empty_cache()
        if args.config == "cfg2" and not args.no_extra_configs:
            configs = {}
            for name in ("cfg3", "cfg4"):
                c2 = workload.CONFIGS[name]
                Bc = (args.cfg3_batch or c2["batch"]) if name == "cfg3" else c2["batch"]
                try:
                    t_h0 = time.perf_counter()
                    uq = workload.build_unique(name, 4)
                    host_ms = 1e3 * (time.perf_counter() - t_h0) / len(uq)
                    mc = measure_config(J, torch, ctx, dev, stream, uq, 0, Bc, max(3, min(args.steps, 10)), peak)
                    steps_c = max(3, min(args.steps, 10))
                    entry = {"workload": c2["desc"], "batch": Bc, "width": c2["width"], "height": c2["height"],
                             "value": Bc * c2["width"] * c2["height"] / 1e6 * steps_c / (mc["ms"] / 1e3), "unit": "MP/s",
                             "ms_per_step": mc["ms"] / steps_c, "fused": mc["fused"], "parity": mc["parity"],
                             "roofline": roofline_of(mc, peak, peak_src, name, Bc),
                             "device_bytes": {"coefficients": mc["info"].coef_bytes, "planes_two_kernel_route_only": mc["info"].plane_bytes, "pixels": mc["info"].out_bytes}}
                    mc["wl"].close()
                    torch.cuda.empty_cache()
                    if name == "cfg4":
                        # what the host does before the device sees anything: 10 scans accumulated into the coefficient buffers
                        t0 = time.perf_counter()
                        for _ in range(5):
                            workload.UniqueImage(uq[0].jpeg)
                        entry["host_accumulate_ms_per_image"] = 1e3 * (time.perf_counter() - t0) / 5
                        nt = max(1, len(all_cpus))
                        o4 = c2["width"] * c2["height"] * 3
                        f4 = torch.empty(Bc * o4, dtype=torch.uint8, pin_memory=True)
                        jobs4, keep4 = files_jobs(J, [uq[0].jpeg], Bc, f4.data_ptr(), o4)
                        ctx.check(J.lib().b200jpg_decode_files(ctx._h, jobs4, Bc, nt))
                        t0 = time.perf_counter()
                        for _ in range(3):
                            ctx.check(J.lib().b200jpg_decode_files(ctx._h, jobs4, Bc, nt))
                        dt = (time.perf_counter() - t0) / 3
                        import oracle
                        want4 = oracle.Decoder(uq[0].jpeg).decode()
                        entry["files"] = {"value": Bc * c2["width"] * c2["height"] / 1e6 / dt, "unit": "MP/s", "host_threads": nt,
                                          "wall_ms_per_batch": 1e3 * dt, "bit_exact_vs_oracle": bool(np.array_equal(f4.numpy()[:o4], want4)),
                                          "api": "b200jpg_decode_files x512: progressive scans are entropy-decoded and accumulated on the host threads "
                                                 "(src/decoder.rs:400-412, 1035-1048), the finished coefficients go to the GPU as sparse streams"}
                        del f4
                    if not args.no_cpu_baseline:   # the CPU arm on the same config (bounded samples), for a like-for-like ratio
                        oi = [OracleImage(u.jpeg) for u in uq[:2]]
                        ncores = os.cpu_count() or 1
                        cv, cs = cpu_hotpath_throughput(oi, ncores, 2.5, args.cpu_arith)
                        cf, cfs = cpu_files_throughput([u.jpeg for u in uq[:2]], c2["width"], c2["height"], ncores, 2.5, args.cpu_arith)
                        entry["cpu_baseline"] = {"value": cv, "unit": "MP/s", "cores": ncores, "kind": "port", "arith": args.cpu_arith, "sample": cs,
                                                 "files": {"value": cf, "unit": "MP/s", "sample": cfs}}
                    configs[name] = entry
                except Exception as e:   # an extra config must not take the headline down with it
                    configs[name] = {"error": repr(e)}
                    torch.cuda.empty_cache()
        if args.config == "cfg2" and not args.no_extra_configs:
            # the kernels behind everything else compute_image accepts, at 1080p x 256: one number each next to the headline
            variants = {}
            vsteps = 3
            for name, v in workload.VARIANTS.items():
                try:
                    uq = [workload.UniqueImage(workload.variant_jpeg(name, k), v.get("scale")) for k in range(2)]
                    mv = measure_config(J, torch, ctx, dev, stream, uq, 0, 256, vsteps, peak)
                    w_out, h_out = uq[0].width, uq[0].height
                    variants[name] = {"workload": v["desc"] + " x256", "value": 256 * w_out * h_out / 1e6 * vsteps / (mv["ms"] / 1e3), "unit": "MP/s (output pixels)",
                                      "parity": mv["parity"],
                                      "kernels": {k: {"ms": x["ms"], "frac": x["frac"]} for k, x in mv["kernels"].items()}}
                    mv["wl"].close()
                except Exception as e:
                    variants[name] = {"error": repr(e)}
                torch.cuda.empty_cache()
            try:   # the x86 build's arithmetic (int16 lanes) on the headline workload
                ctx3 = J.Context(device=local_rank, arith=J.ARITH_SSSE3, stream=stream.cuda_stream)
                m3 = measure_config(J, torch, ctx3, dev, stream, unique, 0, 256, vsteps, peak)
                variants["ssse3_arithmetic"] = {"workload": cfg["desc"] + " x256, arith = ssse3 (bit-exact vs the oracle's src/arch/ssse3.rs restatement)",
                                                "value": 256 * W * H / 1e6 * vsteps / (m3["ms"] / 1e3), "unit": "MP/s", "parity": m3["parity"],
                                                "kernels": {k: {"ms": x["ms"], "frac": x["frac"]} for k, x in m3["kernels"].items()}}
                m3["wl"].close()
                ctx3.close()
            except Exception as e:
                variants["ssse3_arithmetic"] = {"error": repr(e)}
            torch.cuda.empty_cache()
            configs["variants"] = variants
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            oimgs = [OracleImage(u.jpeg) for u in unique[:4]]
            cpu_baseline = cpu_baseline_block(oimgs, jpegs, W, H, cores, args.cpu_seconds, args.cpu_arith)
            # the two feeders agree: the product's host decoder and the oracle's produce the same coefficients
            cpu_baseline["oracle_and_product_decoders_agree"] = bool(all(np.array_equal(a, b) for a, b in zip(oimgs[0].coefs, unique[0].coefs)))
            cpu_baseline["gpu_matches_cpu_scalar_bit_exact"] = bool(np.array_equal(oimgs[0].pixels_scalar, ref0)) if args.arith == "scalar" else None
            import oracle
            px_x86 = oracle.Decoder(jpegs[0], oracle.ARITH_SSSE3_NATIVE).decode()
            dd = np.abs(px_x86.astype(np.int16) - ref0.astype(np.int16))
            cpu_baseline["gpu_vs_cpu_x86_arith_max_abs_diff"] = int(dd.max())
            cpu_baseline["gpu_vs_cpu_x86_arith_frac_over_1"] = float((dd > 1).mean())
    else:
        wl.close()

    if rank == 0:
        line = {
            "metric": "megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": m["ms_max"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i32" if args.arith == "scalar" else "i16", "data": "synthetic",
            "config": {"workload": cfg["desc"], "batch_per_gpu": B, "global_batch": B * world, "width": W, "height": H,
                       "unique_images": U, "arith": args.arith, "k1": args.k1, "k2": args.k2, "fused": m["fused"],
                       "parallelism": "images sharded by index, %d rank(s)" % world,
                       "l2": "inputs (%.1f GB/GPU) far exceed the 126 MB L2; no flush needed" % (info.coef_bytes / 1e9)},
            "parity": m["parity"],
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "configs": configs,
            "gpu_launches": int(m["launches"]),
            "clocks": clocks,
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
