"""Stage the reference's Python hot path for the GPU box.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box, but the checker needs the reference's own `system/map.py`,
`system/tracker.py`, `network/*`, `utils/*` there (the timed `cuda_reference` leg of bench.py, the config-1 parity
run, the operator-level drop-in test).  This recipe copies those files, UNMODIFIED, from where they lie under
/root/reference into `oracle/_ref/pyref/` -- a git-ignored build output next to the reference-built `.so` files
(`oracle/_ref/ext_build/`), shipped by gpurun like them, never committed.  Nothing under nerf-fusion_b200/ reads it.

Usage: python oracle/stage_ref_py.py        (run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
from pathlib import Path

REF = Path(os.environ.get("DFB_REFERENCE_ROOT", "/root/reference"))
OUT = Path(__file__).resolve().parent / "_ref" / "pyref"

FILES = [
    "system/map.py", "system/tracker.py",
    "network/criterion.py", "network/di_decoder.py", "network/di_encoder.py", "network/utility.py",
    "utils/exp_util.py", "utils/motion_util.py", "utils/pt_util.py", "utils/vis_util.py",
    "dataset/production/__init__.py",
    "configs/fusion-lr-kt.yaml",
    "ckpt/default/hyper.json", "ckpt/default/model_300.pth.tar", "ckpt/default/encoder_300.pth.tar",
]


def stage():
    if not (REF / "system" / "map.py").exists():
        print(f"reference not present at {REF}; nothing staged")
        return False
    for rel in FILES:
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        if dst.exists():
            os.chmod(dst, 0o644)
        shutil.copyfile(REF / rel, dst)
    print(f"staged {len(FILES)} reference files under {OUT}")
    return True


if __name__ == "__main__":
    stage()
