"""Device entropy decoding on the GPU (csrc/entropy_dev.h, csrc/ke_entropy.cu) through b200jpg_decode_files.

The host decoder is the restatement of the reference's sequential Huffman loop (src/huffman.rs, src/decoder.rs:1086-1172)
and is itself checked against the oracle elsewhere; here the device path must give the same pixels and the same
per-image status as the host path AND as the oracle, on fixtures, BASELINE-shaped synthetic files and corrupted scans,
and it must really run (scan counters)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, bench_files, reftest_files

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(J):
    dev = J.Context(device=0, entropy=J.ENTROPY_DEVICE)
    host = J.Context(device=0, entropy=J.ENTROPY_HOST)
    yield dev, host
    dev.close()
    host.close()


def same(J, pair, files, nthreads=4):
    dev, host = pair
    o1, s1, _ = J.decode_files(dev, files, nthreads=nthreads)
    o2, s2, _ = J.decode_files(host, files, nthreads=nthreads)
    assert s1 == s2
    for i, (a, b) in enumerate(zip(o1, o2)):
        assert (a is None) == (b is None), i
        if a is not None:
            assert np.array_equal(a, b), i
    return o1, s1


def test_fixtures_device_equals_host_equals_oracle(J, oracle_mod, pair):
    dev, host = pair
    paths = reftest_files(include_disabled=True) + bench_files()
    files = [open(p, "rb").read() for p in paths]
    before = dev.device_scan_counts
    outs, st = same(J, pair, files)
    after = dev.device_scan_counts
    assert after[0] - before[0] >= 20, "the baseline fixtures did not take the device route"
    assert host.device_scan_counts == (0, 0)
    for p, data, o, s_ in zip(paths, files, outs, st):
        try:
            want = oracle_mod.Decoder(data).decode()
        except oracle_mod.OracleError as e:
            assert s_ == -e.code, p
            continue
        assert s_ == 0 and np.array_equal(o, want), p


@pytest.mark.parametrize("shape", [(1920, 1080, 2), (1920, 1080, 0), (640, 480, 1), (333, 217, 2), (8, 8, 2), (3840, 2160, 2)])
def test_synthetic_shapes(J, oracle_mod, pair, shape):
    from jpeg_decoder_b200 import workload
    w, h, ss = shape
    files = [workload.synth_jpeg(w, h, seed=900 + k, subsampling=ss) for k in range(3)]
    dev = pair[0]
    before = dev.device_scan_counts
    outs, st = same(J, pair, files * 2)
    after = dev.device_scan_counts
    assert after[0] - before[0] == 6 and after[1] == before[1], "expected every scan to be accepted by the device"
    assert st == [0] * 6
    for f, o in zip(files, outs):
        assert np.array_equal(o, oracle_mod.Decoder(f).decode())


def test_large_batch_many_threads(J, oracle_mod, pair):
    from jpeg_decoder_b200 import workload
    uniq = [workload.synth_jpeg(1920, 1080, seed=1234 + k, subsampling=2) for k in range(4)]
    wants = [oracle_mod.Decoder(f).decode() for f in uniq]
    files = [uniq[i % 4] for i in range(96)]
    dev = pair[0]
    for nthreads in (2, 16):
        outs, st, _ = J.decode_files(dev, files, nthreads=nthreads)
        assert st == [0] * 96
        for i, o in enumerate(outs):
            assert np.array_equal(o, wants[i % 4]), i


def test_corrupted_scans(J, oracle_mod, pair):
    """Random damage inside the entropy-coded segment: whatever the host path (= the reference's behaviour) reports
    or decodes, the device path reports or decodes the same -- by accepting only what it can prove and sending the
    rest back to the host loop."""
    from jpeg_decoder_b200 import workload
    rng = np.random.default_rng(11)
    base = [workload.synth_jpeg(333, 217, seed=5, subsampling=2), open(os.path.join(GOLDEN, "benches", "tower.jpg"), "rb").read(),
            open(os.path.join(GOLDEN, "reftest", "mozilla", "jpg-size-33x33.jpg"), "rb").read()]
    files = []
    for data in base:
        sos = data.rfind(b"\xff\xda")
        for t in range(60):
            b = bytearray(data)
            at = int(rng.integers(sos + 14, len(b) - 2))
            kind = t % 4
            if kind == 0:
                b[at] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                b[at] = int(rng.integers(0, 256))
            elif kind == 2:
                del b[at:min(len(b) - 2, at + 1 + int(rng.integers(0, 64)))]
            else:
                for q in range(at, min(len(b) - 2, at + 1 + int(rng.integers(0, 32)))):
                    b[q] = int(rng.integers(0, 256))
            files.append(bytes(b))
    dev = pair[0]
    before = dev.device_scan_counts
    outs, st = same(J, pair, files, nthreads=8)
    after = dev.device_scan_counts
    assert after[0] > before[0] and after[1] > before[1], "expected both accepted and handed-back scans"
    # and the oracle agrees on a sample
    for f, o, s_ in list(zip(files, outs, st))[::7]:
        try:
            want = oracle_mod.Decoder(f).decode()
        except oracle_mod.OracleError as e:
            assert s_ == -e.code
            continue
        assert s_ == 0 and np.array_equal(o, want)


def test_crashtest_files(J, pair):
    files = [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "crashtest", "**", "*.jpg"), recursive=True))]
    assert len(files) > 50
    same(J, pair, files)


def test_device_resident_outputs(J, oracle_mod, pair):
    """b200jpg_file_job.out may point to device memory: the pixels stay on the GPU (no D2H), same bytes."""
    import torch
    from jpeg_decoder_b200 import workload
    files = [workload.synth_jpeg(640, 480, seed=70 + k, subsampling=2) for k in range(5)] + [open(os.path.join(GOLDEN, "benches", "tower.jpg"), "rb").read()]
    wants = [oracle_mod.Decoder(f).decode() for f in files]
    d_outs = [torch.zeros(w.size, dtype=torch.uint8, device="cuda:0") for w in wants]
    for ctx in pair:
        for t in d_outs:
            t.zero_()
        st, lens = J.decode_files_into(ctx, files, [t.data_ptr() for t in d_outs], [t.numel() for t in d_outs], nthreads=3)
        torch.cuda.synchronize()
        assert st == [0] * len(files) and lens == [w.size for w in wants]
        for t, w in zip(d_outs, wants):
            assert np.array_equal(t.cpu().numpy(), w)


def test_restart_interval_scans(J, oracle_mod, pair):
    """DRI files: every restart interval goes to the kernels as a scan of its own; damaged ones behave like the host."""
    from test_entropy_emul import _dri_jpeg
    files = [_dri_jpeg(640, 480, 0, 80), _dri_jpeg(1920, 1080, 2, 120), _dri_jpeg(200, 200, 2, 7), _dri_jpeg(333, 217, 1, 21),
             _dri_jpeg(64, 64, 2, 1)]  # the last one has intervals too small to qualify: host
    dev = pair[0]
    before = dev.device_scan_counts
    outs, st = same(J, pair, files)
    after = dev.device_scan_counts
    assert after[0] - before[0] == 4 and after[1] == before[1]
    assert st == [0] * len(files)
    for f, o in zip(files, outs):
        assert np.array_equal(o, oracle_mod.Decoder(f).decode())
    rng = np.random.default_rng(5)
    damaged = []
    for data in files[:3]:
        sos = data.rfind(b"\xff\xda")
        for t in range(40):
            b = bytearray(data)
            at = int(rng.integers(sos + 14, len(b) - 2))
            if t % 3 == 0:
                b[at] ^= 1 << int(rng.integers(0, 8))
            elif t % 3 == 1:
                del b[at:min(len(b) - 2, at + 1 + int(rng.integers(0, 48)))]
            else:
                b[at:at] = bytes(int(rng.integers(0, 255)) for _ in range(int(rng.integers(1, 5))))
            damaged.append(bytes(b))
    same(J, pair, damaged, nthreads=8)


def test_concurrent_calls(J, oracle_mod):
    """b200jpg_decode_files from several host threads at once -- on one shared context (calls serialise on the context's
    engine) and on contexts of their own (two engines share the GPU), there with Decoder.decode() in between (a context
    with its own stream per host thread is the documented way to call the rest of the ABI concurrently)."""
    import threading
    from jpeg_decoder_b200 import workload
    files = [workload.synth_jpeg(640, 360, seed=300 + k, subsampling=2) for k in range(6)]
    files.append(workload.synth_jpeg(320, 200, seed=310, subsampling=0, progressive=True))
    wants = [oracle_mod.Decoder(f).decode() for f in files]
    shared = J.Context(device=0)
    own = [J.Context(device=0) for _ in range(2)]
    errors = []

    def worker(ctx, reps, seed, single):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(reps):
                order = rng.permutation(len(files))
                outs, st, _ = J.decode_files(ctx, [files[i] for i in order], nthreads=2)
                for i, o, s_ in zip(order, outs, st):
                    if s_ != 0 or not np.array_equal(o, wants[i]):
                        errors.append((seed, int(i), s_))
                i = int(rng.integers(0, len(files)))
                if single and not np.array_equal(J.Decoder(files[i], ctx).decode(), wants[i]):
                    errors.append((seed, i, "decode"))
        except Exception as e:  # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=worker, args=(shared, 6, 1, False)), threading.Thread(target=worker, args=(shared, 6, 2, False)),
               threading.Thread(target=worker, args=(own[0], 6, 3, True)), threading.Thread(target=worker, args=(own[1], 6, 4, True))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "deadlock"
    assert not errors, errors[:5]
    for c in own + [shared]:
        c.close()
