"""Build the reference's only native kernel — rope_2d (pi3/models/curope/kernels.cu + curope.cpp) — for sm_100a, FROM THE
SOURCES WHERE THEY LIE under /root/reference, into oracle/_ref/ref_curope.so (git-ignored, travels with gpurun).
Test/bench infrastructure only: scripts/ab_qkv_fused.py times it next to pi3b_gemm_qkv / pi3b_qknorm_rope on the same box
("beat the one sm_100 kernel the reference ships"). Same flags as the reference's setup.py
(-O3 --use_fast_math), our own recipe (torch.utils.cpp_extension.load), never the reference's build system.
Skips silently when /root/reference is absent (GPU box: uses the prebuilt file)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/pi3/models/curope"
OUT = os.path.join(ROOT, "oracle", "_ref")


def build() -> str:
    so = os.path.join(OUT, "ref_curope.so")
    if not os.path.isdir(SRC):
        return so if os.path.exists(so) else ""
    if os.path.exists(so):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, "build_curope")
    os.makedirs(bdir, exist_ok=True)
    load(name="ref_curope", sources=[os.path.join(SRC, "curope.cpp"), os.path.join(SRC, "kernels.cu")],
         extra_cuda_cflags=["-O3", "--use_fast_math", "-gencode", "arch=compute_100a,code=sm_100a"], extra_cflags=["-O3"],
         build_directory=bdir, is_python_module=False, verbose=False)
    import shutil
    shutil.copy(os.path.join(bdir, "ref_curope.so"), so)
    shutil.rmtree(bdir, ignore_errors=True)
    return so


if __name__ == "__main__":
    print(build() or "reference sources absent and no prebuilt library")
