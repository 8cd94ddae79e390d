"""Builds jpeg_decoder_b200/libb200jpg.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libb200jpg.so")
SOURCES = ["k0_expand.cu", "ke_entropy.cu", "k1_idct.cu", "k2_color.cu", "kf_fused.cu", "pipeline.cu", "sbs_pipeline.cpp", "sbs_api.cpp", "stream_engine.cpp", "host_decoder.cpp",
           "decoder_api.cpp", "files_api.cpp"]
HEADERS = ["device_types.h", "kernels.h", "host_decoder.h", "context.h", "sbs.h", "sbs_pipeline.h", "batch_internal.h", "ptx.cuh", "idct_core.cuh", "color_core.cuh", "ring_book.h", "stream_engine.h", "entropy_dev.h", "entropy_host.h", os.path.join("..", "..", "include", "b200jpg.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3,-pthread", "-shared"]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: an experiment build beside the product library (scripts/, profiling only), e.g.
    build(out="libb200jpg_sub512.so", defines=["B200JPG_ENT_SUB_BITS=512"]); selected at run time with B200JPG_SO."""
    if out is None and os.environ.get("B200JPG_SO"):   # a prebuilt experiment library (never built implicitly)
        return os.path.join(HERE, os.environ["B200JPG_SO"])
    target = os.path.join(HERE, out) if out else SO
    if out is None and not force and up_to_date():
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
