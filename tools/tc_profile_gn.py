"""Cycle timeline of block 0 / group 0 in one fused Gauss-Newton evaluation launch (gn_eval_kernel) at the bench size.
Needs a library built with EXTRA=-DDFB_TC_PROFILE: DFB_LIB=nerf-fusion_b200/libdifusion_b200_prof.so python tools/tc_profile_gn.py"""
import sys, ctypes as C, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
dfb = importlib.import_module("nerf-fusion_b200")
raw = dfb._lib.load()._cdll
dev = "cuda:0"
frames, _raw, seq = bench.gen_frames(dfb, 6, dev, 0)
m, trk = bench.make_system(dfb, dev)
calib = dfb.FrameIntrinsic(*dfb.synth.ICL_CALIB)
first = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))
buf = (C.c_ulonglong * 256)(); n = C.c_int(0)
for i, (d, c) in enumerate(frames):
    if i == 5:
        torch.cuda.synchronize(); raw.dfb_debug_read_prof(buf, C.byref(n))
        trk.args.iter_config = [{"n": 1, "type": [["sdf"], ["rgb", int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 0]]}]
    bench.refresh(dfb, m, trk, i, d, c, calib, first)
torch.cuda.synchronize(); raw.dfb_debug_read_prof(buf, C.byref(n))
t = np.array(list(buf)[:n.value], dtype=np.int64)
print("marks", n.value)
per_tile = ["tile start", "lookup+stage"] + [f"fwd L{l} {x}" for l in range(4) for x in ["enter", "synced", "issued", "mma done"]] + ["heads done"] + \
           [f"bwd L{l} {x}" for l in (2, 1, 0) for x in ["enter", "synced", "issued", "mma done"]] + ["tile end"]
tail = ["tiles done", "pixels done", "all groups done", "sums out", "kernel end"]
names = ["kernel start", "prologue done"] + [f"t0 {x}" for x in per_tile] + [f"t1 {x}" for x in per_tile] + tail
t0 = t[0]; prev = t0
for i, v in enumerate(t[:len(names)]):
    nm = names[i]
    if ("fwd" in nm or "bwd" in nm or "heads" in nm or "lookup" in nm) and "-v" not in sys.argv:
        prev = v; continue
    print(f"{i:3d} {nm:28s} +{int(v - prev):6d}  @{int(v - t0):7d} cycles"); prev = v

sb = (C.c_ulonglong * 16)()
raw.dfb_debug_read_step_prof(sb)
lab = ["enter", "state loaded", "H,g scaled", "solved", "pose updated", "pose published", "state stored", "record written"]
for half, what in ((0, "step 0 (solve + pose update)"), (1, "step 1 (evaluation-only pass, ends the group)")):
    st = np.array(list(sb)[8 * half:8 * half + 8], dtype=np.int64)
    print(what + ":")
    prev = st[0]
    for i in range(1, 8):
        if st[i] < prev:            # phase not executed in this step
            continue
        print(f"   {lab[i]:16s} +{int(st[i] - prev):6d} cycles"); prev = st[i]
    print(f"   total            {int(prev - st[0]):7d} cycles")
