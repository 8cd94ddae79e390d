"""On-device output formats (b200jpg_batch_format_device): the interleaved RGB8 slab of a
device-resident run rearranged for a GPU-side consumer -- planar uint8, float32 NHWC, float32 NCHW -- against numpy
on the oracle's pixels.  uint8 and unscaled float are exact; the scaled float is one fp32 FMA per sample."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_output_formats(J, oracle_mod):
    import torch
    from jpeg_decoder_b200 import workload
    dev = torch.device("cuda", 0)
    ctx = J.Context(device=0)
    shapes = [(320, 176, 2), (161, 99, 0), (33, 17, 2), (2064, 9, 0)]
    uniq = [workload.UniqueImage(workload.synth_jpeg(w, h, seed=90 + k, subsampling=ss)) for k, (w, h, ss) in enumerate(shapes)]
    keep = []
    descs = [J.make_image_desc(u.width, u.height, u.components, u.qts, u.coefs, u.color_transform, keep) for u in uniq]
    batch = J.Batch(ctx, descs)
    info = batch.info
    d_coefs = torch.zeros(info.coef_bytes, dtype=torch.uint8, device=dev)
    d_planes = torch.zeros(info.plane_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(info.out_bytes, dtype=torch.uint8, device=dev)
    for j, u in enumerate(uniq):
        lay = batch.layout(j)
        for k, c in enumerate(u.coefs):
            t = torch.from_numpy(c.view(np.uint8)).to(dev)
            d_coefs[lay["coef_off"][k]:lay["coef_off"][k] + t.numel()].copy_(t)
    batch.run_device(d_coefs.data_ptr(), d_planes.data_ptr(), d_out.data_ptr(), 3)
    wants = [oracle_mod.Decoder(u.jpeg).decode().reshape(u.height, u.width, 3) for u in uniq]
    d_u8 = torch.zeros(info.out_bytes, dtype=torch.uint8, device=dev)
    d_f32 = torch.zeros(info.out_bytes, dtype=torch.float32, device=dev)
    assert batch.format_device(d_out.data_ptr(), J.FMT_RGB8_PLANAR, d_u8.data_ptr()) == [0] * len(uniq)
    ctx.synchronize()
    for j, (u, want) in enumerate(zip(uniq, wants)):
        lay = batch.layout(j)
        got = d_u8[lay["out_off"]:lay["out_off"] + lay["out_len"]].cpu().numpy().reshape(3, u.height, u.width)
        assert np.array_equal(got, want.transpose(2, 0, 1)), ("planar u8", j)
    scale, bias = [1 / 255.0, 1 / 128.0, 2.0], [0.0, -1.0, 0.5]
    for fmt, sc, bi in ((J.FMT_RGB_F32_NHWC, None, None), (J.FMT_RGB_F32_NCHW, None, None), (J.FMT_RGB_F32_NHWC, scale, bias), (J.FMT_RGB_F32_NCHW, scale, bias)):
        d_f32.zero_()
        batch.format_device(d_out.data_ptr(), fmt, d_f32.data_ptr(), sc, bi)
        ctx.synchronize()
        for j, (u, want) in enumerate(zip(uniq, wants)):
            lay = batch.layout(j)
            flat = d_f32[lay["out_off"]:lay["out_off"] + lay["out_len"]].cpu().numpy()
            ref = want.astype(np.float32)
            if sc is not None:
                ref = ref * np.array(sc, np.float32) + np.array(bi, np.float32)
            got = flat.reshape(u.height, u.width, 3) if fmt == J.FMT_RGB_F32_NHWC else flat.reshape(3, u.height, u.width).transpose(1, 2, 0)
            if sc is None:
                assert np.array_equal(got, ref), (fmt, j)
            else:
                assert np.allclose(got, ref, rtol=0, atol=1e-5), (fmt, j, float(np.abs(got - ref).max()))
    batch.close()
    ctx.close()
