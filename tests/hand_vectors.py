"""Hand-derived upsampling vectors: small planes pushed through the reference's formulas (src/upsampler.rs) BY HAND --
the arithmetic is written out in the comments -- so that the expected bytes depend on nobody's code.  Used on the CPU
against the oracle (tests/test_oracle_kat.py) and on the GPU against the kernels (tests/test_gpu_parity.py)."""
import numpy as np

# UpsamplerH2V2, src/upsampler.rs:191-228, chroma 2x2 = [[10, 50], [90, 130]] -> 4x4.
#   near/far rows (200-203): row 0 -> (0, 0); row 1 -> (0, 1); row 2 -> (1, 0); row 3 -> (1, 1)
#   t = 3 * near + far; out[0] = (t0 + 2) >> 2; out[1] = (3 t0 + t1 + 8) >> 4; out[2] = (3 t1 + t0 + 8) >> 4; out[3] = (t1 + 2) >> 2
#   row 0: t = (40, 200)   -> 42>>2, 328>>4, 648>>4, 202>>2      = 10, 20, 40, 50
#   row 1: t = (120, 280)  -> 122>>2, 648>>4, 968>>4, 282>>2     = 30, 40, 60, 70
#   row 2: t = (280, 440)  -> 282>>2, 1288>>4, 1608>>4, 442>>2   = 70, 80, 100, 110
#   row 3: t = (360, 520)  -> 362>>2, 1608>>4, 1928>>4, 522>>2   = 90, 100, 120, 130
H2V2_IN = np.array([[10, 50], [90, 130]], dtype=np.uint8)
H2V2_OUT = np.array([[10, 20, 40, 50], [30, 40, 60, 70], [70, 80, 100, 110], [90, 100, 120, 130]], dtype=np.uint8)

# UpsamplerH2V1, src/upsampler.rs:134-163, one row [10, 50, 90] -> 6 samples:
#   out[0] = in[0] = 10; out[1] = (3*10 + 50 + 2) >> 2 = 20; i = 1: sample = 152 -> (152 + 10) >> 2 = 40, (152 + 90) >> 2 = 60;
#   out[4] = (3*90 + 50 + 2) >> 2 = 80; out[5] = in[2] = 90
H2V1_IN = np.array([[10, 50, 90]], dtype=np.uint8)
H2V1_OUT = np.array([[10, 20, 40, 60, 80, 90]], dtype=np.uint8)

# UpsamplerH1V2, src/upsampler.rs:165-189, column [10, 90] -> 4 rows: near/far as above -> (3*near + far + 2) >> 2
#   row 0: (30 + 10 + 2) >> 2 = 10; row 1: (30 + 90 + 2) >> 2 = 30; row 2: (270 + 10 + 2) >> 2 = 70; row 3: (270 + 90 + 2) >> 2 = 90
H1V2_IN = np.array([[10], [90]], dtype=np.uint8)
H1V2_OUT = np.array([[10], [30], [70], [90]], dtype=np.uint8)

CASES = {"h2v2": ((2, 2), H2V2_IN, H2V2_OUT), "h2v1": ((2, 1), H2V1_IN, H2V1_OUT), "h1v2": ((1, 2), H1V2_IN, H1V2_OUT)}


def planes_for(make_components, name, rng):
    """Three-component image whose SECOND component is the hand vector (luma sampled (h, v), chroma (1, 1)); planes are
    full block grids filled with noise outside the valid samples, so that padding must not leak into the result."""
    (h, v), cin, cout = CASES[name]
    out_h, out_w = cout.shape
    comps, _ = make_components(out_w, out_h, [(h, v), (1, 1), (1, 1)])
    planes = []
    for k, c in enumerate(comps):
        p = rng.integers(0, 256, (c.block_h * 8, c.block_w * 8)).astype(np.uint8)
        if k == 1:
            assert (c.size_h, c.size_w) == cin.shape
            p[:cin.shape[0], :cin.shape[1]] = cin
        planes.append(p.reshape(-1))
    return comps, planes, out_w, out_h, cout
