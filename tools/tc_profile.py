"""Cycle timeline of the tiles CTA 0 / group 0 processes in one tcgen05 sdf_hg launch at the bench size (~51k queries).
Needs a library built with EXTRA=-DDFB_TC_PROFILE: DFB_LIB=nerf-fusion_b200/libdifusion_b200_prof.so python tools/tc_profile.py"""
import sys, ctypes as C, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from util import GOLD, make_map, pkg
d = pkg(); lib = d._lib.load()
raw = lib._cdll
W = d.weights.load_npz(GOLD / "weights.npz"); G = dict(np.load(GOLD / "map_golden.npz"))
DEV = "cuda:0"
m = make_map(W)
Pw, Nw = torch.from_numpy(G["Pw"]).to(DEV), torch.from_numpy(G["Nw"]).to(DEV)
m.integrate_keyframe(Pw, Nw)
trk = d.SDFTracker(m, dict(iter_config=[], sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                           rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2)))
Pc = G["Pc"]
reps = int(np.ceil(51024 / Pc.shape[0]))
P = torch.from_numpy(np.tile(Pc, (reps, 1))[:51024]).to(DEV).contiguous()
last = d.Isometry.from_matrix(G["hg_last_R"], G["hg_last_t"]); delta = d.Isometry.from_matrix(G["hg_delta_R"], G["hg_delta_t"])
buf = (C.c_ulonglong * 256)(); n = C.c_int(0)
for it in range(3):
    trk.compute_sdf_Hg(0, last, delta, P)
    raw.dfb_debug_read_prof(buf, C.byref(n))
t = np.array(list(buf)[:n.value], dtype=np.int64)
print("queries", P.shape[0], "marks", n.value)
names = ["kernel start", "prologue done"]
layer = ["fence+sync", "issue", "mma wait", "epilogue+next"]
per_tile = ["tile start", "lookup+stage"] + [f"fwd L{l} {x}" for l in range(4) for x in ["enter", "synced", "issued", "mma done"]] + ["heads done"] + \
           [f"bwd L{l} {x}" for l in (2, 1, 0) for x in ["enter", "synced", "issued", "mma done"]] + ["tile end"]
n_tiles = (n.value - 4) // len(per_tile)
for k in range(n_tiles):
    names += [f"t{k} {x}" for x in per_tile]
names += ["tmem freed", "kernel end"]
t0 = t[0]
prev = t0
for i, v in enumerate(t):
    print(f"{i:3d} {names[i] if i < len(names) else '?':28s} +{int(v - prev):6d}  @{int(v - t0):7d} cycles")
    prev = v
