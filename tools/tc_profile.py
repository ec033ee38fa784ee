"""Cycle timeline of one tile of the tcgen05 sdf_hg kernel (needs a library built with EXTRA=-DDFB_TC_PROFILE)."""
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
P = torch.from_numpy(np.tile(G["Pc"], (4, 1))).to(DEV).contiguous()
last = d.Isometry.from_matrix(G["hg_last_R"], G["hg_last_t"]); delta = d.Isometry.from_matrix(G["hg_delta_R"], G["hg_delta_t"])
buf = (C.c_ulonglong * 256)(); n = C.c_int(0)
for it in range(3):
    trk.compute_sdf_Hg(0, last, delta, P)
    raw.dfb_debug_read_prof(buf, C.byref(n))
t = np.array(list(buf)[:n.value], dtype=np.int64)
print("marks", n.value)
dt = np.diff(t)
labels = ["fence+sync", "issue", "mma wait", "epilogue"]
for i, v in enumerate(dt[:64]):
    print(i, labels[i % 4] if i < 1000 else "", int(v))
