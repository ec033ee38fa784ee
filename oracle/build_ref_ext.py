"""Build the reference's own CUDA extensions (system/ext/*) from the sources where they
lie under /root/reference into oracle/_ref/ext_build/ (git-ignored, shipped by gpurun).

TEST INFRASTRUCTURE ONLY. The resulting .so files are used by tests/ and bench.py's
reference legs to pin the oracle against the real reference kernels on a B200. Nothing
under nerf-fusion_b200/ may import them. No reference source is copied into the repo.

Usage: python oracle/build_ref_ext.py [name ...]   (names: indexing marching_cubes imgproc pcproc)
"""
import os
import sys
from pathlib import Path

REF = Path(os.environ.get("DFB_REFERENCE_ROOT", "/root/reference")) / "system" / "ext"
OUT = Path(__file__).resolve().parent / "_ref" / "ext_build"

SOURCES = {
    "indexing": ["indexing/indexing.cpp", "indexing/indexing.cu"],
    "marching_cubes": ["marching_cubes/mc.cpp", "marching_cubes/mc_interp_kernel.cu"],
    "imgproc": ["imgproc/imgproc.cu", "imgproc/imgproc.cpp", "imgproc/photometric.cu"],
    "pcproc": ["pcproc/pcproc.cpp", "pcproc/pcproc.cu", "pcproc/cuda_kdtree.cu"],
}


def build(names):
    if not REF.exists():
        print(f"reference not present at {REF}; nothing to build")
        return
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    for name in names:
        bdir = OUT / name
        bdir.mkdir(parents=True, exist_ok=True)
        if (bdir / f"ref_{name}.so").exists():
            print(f"[skip] {name} already built")
            continue
        print(f"[build] {name}", flush=True)
        load(name=f"ref_{name}", sources=[str(REF / s) for s in SOURCES[name]],
             build_directory=str(bdir), verbose=False, is_python_module=False)
        print(f"[done] {name}", flush=True)


if __name__ == "__main__":
    build(sys.argv[1:] or list(SOURCES))
