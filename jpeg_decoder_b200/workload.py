"""Synthetic workloads of BASELINE.json and multi-GPU sharding helpers.

Images are generated in-process with PIL (libjpeg-turbo) exactly as SURVEY 8(d) specifies, entropy-
decoded by the product's own host decoder (csrc/host_decoder.cpp) into dense coefficient buffers --
the input format of the hot path -- and replicated to the batch size.  The device format is dense
int16, so kernel bytes are content-independent.
"""
import io
import os

import numpy as np

GOLDEN_BENCHES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "benches")

CONFIGS = {
    # BASELINE.json configs[1..3] (configs[0], tower.jpg on the CPU, is the reference arm's own case; configs[4] is
    # cfg2 at 8 x 1024 images, i.e. `bench.py --gpus 8`)
    "cfg2": dict(width=1920, height=1080, subsampling=2, seed=1234, batch=1024,
                 desc="1024 synthetic 1920x1080 baseline 4:2:0 JPEGs (q90)"),
    "cfg3": dict(width=3840, height=2160, subsampling=0, seed=5000, batch=1024,
                 desc="1024 synthetic 3840x2160 baseline 4:4:4 JPEGs (q90)"),
    "cfg4": dict(width=512, height=512, file="tower_progressive.jpg", batch=512,
                 desc="benches/tower_progressive.jpg x512 (progressive, 10 scans, 4:4:4; the host accumulates the scans "
                      "and feeds the finished coefficients)"),
    "tiny": dict(width=256, height=144, subsampling=2, seed=77, batch=8, desc="smoke-sized 4:2:0"),
}


def synth_pixels(width, height, seed):
    """Smooth plaid + 4x4-block noise + per-pixel noise, SURVEY 8(d) cfg2 recipe."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.empty((height, width, 3), dtype=np.float32)
    for c in range(3):
        px, py = rng.uniform(41, 131, size=2)
        amp = rng.uniform(80, 100)
        ph = rng.uniform(0, 6.28, size=2)
        img[..., c] = 127 + amp * 0.5 * (np.sin(xx * (6.2832 / px) + ph[0]) + np.cos(yy * (6.2832 / py) + ph[1]))
    blk = rng.normal(0, 12, size=((height + 3) // 4, (width + 3) // 4, 3)).astype(np.float32)
    img += np.repeat(np.repeat(blk, 4, axis=0), 4, axis=1)[:height, :width]
    img += rng.normal(0, 3, size=img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_jpeg(width, height, seed, subsampling=2, quality=90, progressive=False):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(synth_pixels(width, height, seed)).save(buf, "JPEG", quality=quality, subsampling=subsampling,
                                                             progressive=progressive)
    return buf.getvalue()


# Variants of the hot path that BASELINE.json does not name but compute_image accepts:
# each is measured by bench.py at 1080p so that the kernels behind them have a number next to the headline.
VARIANTS = {
    "ycbcr422": dict(width=1920, height=1080, mode="RGB", subsampling=1, seed=7100, desc="1920x1080 baseline 4:2:2 (H2V1 chroma)"),
    "gray": dict(width=1920, height=1080, mode="L", subsampling=None, seed=7200, desc="1920x1080 baseline greyscale"),
    "cmyk": dict(width=1920, height=1080, mode="CMYK", subsampling=None, seed=7300, desc="1920x1080 baseline CMYK (4 components, Adobe APP14)"),
    "scaled_half": dict(width=1920, height=1080, mode="RGB", subsampling=2, seed=7400, scale=(960, 540),
                        desc="1920x1080 4:2:0 decoded at 1/2 size (Decoder::scale -> 4x4 IDCT)"),
}


def variant_jpeg(name, k=0):
    from PIL import Image
    v = VARIANTS[name]
    px = synth_pixels(v["width"], v["height"], v["seed"] + k)
    if v["mode"] == "L":
        im = Image.fromarray(px[..., 0])
    elif v["mode"] == "CMYK":
        im = Image.fromarray(np.concatenate([px, 255 - px[..., :1]], axis=-1), "CMYK")
    else:
        im = Image.fromarray(px)
    buf = io.BytesIO()
    kw = {} if v["subsampling"] is None else {"subsampling": v["subsampling"]}
    im.save(buf, "JPEG", quality=90, **kw)
    return buf.getvalue()


def config_jpeg(cfg_name, k=0):
    """The k-th distinct JPEG of a config: the reference's bench file (same bytes for every k) or a synthetic one."""
    cfg = CONFIGS[cfg_name]
    if "file" in cfg:
        with open(os.path.join(GOLDEN_BENCHES, cfg["file"]), "rb") as f:
            return f.read()
    return synth_jpeg(cfg["width"], cfg["height"], cfg["seed"] + k, cfg["subsampling"])


class UniqueImage:
    """One entropy-decoded image: geometry + dense coefficients (numpy, host)."""

    def __init__(self, jpeg_bytes, scale=None):
        from . import Decoder
        dec = Decoder(jpeg_bytes)
        if scale is not None:   # Decoder::scale before decoding: the components come back with dct_scale 4 / 2 / 1
            dec.read_info()
            dec.scale(*scale)
        desc = dec.entropy_decode()
        self.width, self.height, self.ncomp = desc.width, desc.height, desc.ncomp
        self.color_transform = desc.color_transform
        self.components = [desc.comps[i] for i in range(desc.ncomp)]
        # copy out: the decoder owns the buffers
        self.coefs = [dec.coefficients(desc, i) for i in range(desc.ncomp)]
        self.qts = [dec.qtable(desc, i) for i in range(desc.ncomp)]
        self.jpeg_bytes = len(jpeg_bytes)
        self.jpeg = jpeg_bytes
        dec.close()

    @property
    def coef_bytes(self):
        return sum(c.nbytes for c in self.coefs)


def build_unique(cfg_name, n_unique, first_index=0):
    if "file" in CONFIGS[cfg_name]:
        n_unique = 1
    return [UniqueImage(config_jpeg(cfg_name, first_index + k)) for k in range(n_unique)]


def shard_range(n_items, rank, world_size):
    """Static contiguous split by image index: rank r gets [lo, hi)."""
    lo = (n_items * rank) // world_size
    hi = (n_items * (rank + 1)) // world_size
    return lo, hi


def broadcast_assignment(n_items, world_size, dist=None, device="cpu"):
    """Rank 0 computes the (lo, hi) table and broadcasts it (the only collective on the data path's
    control plane); returns an int64 array [world_size, 2].  Works with gloo (CPU) and nccl."""
    import torch
    table = torch.zeros((world_size, 2), dtype=torch.int64, device=device)
    if dist is None or not dist.is_initialized() or dist.get_rank() == 0:
        for r in range(world_size):
            lo, hi = shard_range(n_items, r, world_size)
            table[r, 0], table[r, 1] = lo, hi
    if dist is not None and dist.is_initialized():
        dist.broadcast(table, src=0)
    return table.cpu().numpy()


def gather_stats(local, dist=None, device="cpu"):
    """all_gather of a small per-rank float64 vector (pixels, device seconds, checksum, ...)."""
    import torch
    t = torch.tensor(local, dtype=torch.float64, device=device)
    if dist is None or not dist.is_initialized():
        return t.cpu().numpy()[None, :]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()
