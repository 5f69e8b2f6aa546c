"""Build the UNMODIFIED reference DFGNN CUDA extensions for sm_100a.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``dfgnn_b200/``) may
import anything produced here.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py`` (reference/cpu_baseline legs, and the side-by-side GPU
baseline) load ``oracle/_ref/*.so``.

The sources are compiled *where they lie* under ``/root/reference`` (read-only);
no reference source is copied into this repository.  Outputs go to
``oracle/_ref/`` only (git-ignored, but shipped to the GPU box by gpurun).

Source lists follow the reference's own ``setup.py:29-39`` (fused_gatconv) and
``setup.py:46-56`` (fused_gtconv); the only change is the ``-gencode`` target
(the reference ships sm_80/sm_90 SASS only, ``setup.py:41,58``) and ``-DNDEBUG``
(``fused_gatconv.cpp:176-184`` names an undeclared variable inside ``assert``).

Run:  python oracle/build_ref.py            (about 3-6 minutes on 8 cores)
"""
from __future__ import annotations

import os
import sys

REF_ROOT = os.environ.get("DFGNN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

GAT_SOURCES = [
    "fused_gatconv.cpp",
    "fused_gatconv_kernel.cu",
    "fused_gatconv_hyper.cu",
    "fused_gatconv_hyper_recompute.cu",
    "fused_gatconv_hyper_v2.cu",
    "fused_gatconv_softmax.cu",
    "fused_gatconv_hyper_ablation.cu",
    "fused_gatconv_tiling.cu",
    "fused_gatconv_softmax_gm.cu",
]
GT_SOURCES = [
    "fused_gtconv.cpp",
    "fused_gtconv_csr.cu",
    "fused_gtconv_hyper.cu",
    "fused_gtconv_tiling.cu",
    "fused_gtconv_hyper_ablation.cu",
    "fused_gtconv_softmax.cu",
    "fused_gtconv_softmax_gm.cu",
    "fused_gtconv_backward.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-DNDEBUG",
]


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "DFGNN", "src"))


def built() -> bool:
    return all(
        os.path.exists(os.path.join(OUT, name, name + ".so"))
        for name in ("fused_gtconv", "fused_gatconv")
    )


OPERATOR_FILES = ("fused_gtconv.py", "fused_gatconv.py")


def stage_operators() -> bool:
    """Place the reference's UNMODIFIED Python operator files (DFGNN/operators/*.py) next to the
    compiled extensions, under the git-ignored oracle/_ref/operators/, so that the GPU box (which
    has no /root/reference) can run them against the drop-in shim of integration/
    (tests/test_integration_gpu.py).  Byte-for-byte copies, never committed."""
    import shutil
    src_dir = os.path.join(REF_ROOT, "DFGNN", "operators")
    if not os.path.isdir(src_dir):
        return operators_staged()
    dst_dir = os.path.join(OUT, "operators")
    os.makedirs(dst_dir, exist_ok=True)
    for f in OPERATOR_FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(dst_dir, f))
    return operators_staged()


def operators_staged() -> bool:
    return all(os.path.exists(os.path.join(OUT, "operators", f)) for f in OPERATOR_FILES)


def build(verbose: bool = False) -> bool:
    """Compile both reference extensions.  Returns True when both .so exist."""
    if have_reference():
        stage_operators()
    if built():
        return True
    if not have_reference():
        return False
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    for name, srcs in (("fused_gtconv", GT_SOURCES), ("fused_gatconv", GAT_SOURCES)):
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        src_dir = os.path.join(REF_ROOT, "DFGNN", "src", name)
        load(
            name=name,
            sources=[os.path.join(src_dir, s) for s in srcs],
            extra_cflags=["-O2", "-DNDEBUG"],
            extra_cuda_cflags=NVCC_FLAGS,
            extra_ldflags=["-lcurand"],
            build_directory=bdir,
            with_cuda=True,
            is_python_module=False,
            verbose=verbose,
        )
    return built()


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
