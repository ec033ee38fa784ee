"""Host-side view of a frame (no extra syncs): time from frame start to the Gauss-Newton call, inside it, and after it."""
import sys, time, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
frames, _raw, seq = bench.gen_frames(dfb, 45, dev, 0)
m, trk = bench.make_system(dfb, dev)
calib = dfb.FrameIntrinsic(*dfb.synth.ICL_CALIB)
first = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))
l2 = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
marks = {}
orig = trk._gauss_newton_native
def gn(*a, **k):
    marks['gn0'] = time.perf_counter(); r = orig(*a, **k); marks['gn1'] = time.perf_counter(); return r
trk._gauss_newton_native = gn
orig_fe = trk._frontend_graphed
def fe(*a, **k):
    marks['fe0'] = time.perf_counter(); r = orig_fe(*a, **k); marks['fe1'] = time.perf_counter(); return r
trk._frontend_graphed = fe
rows = []
for i, (d, c) in enumerate(frames):
    t0 = time.perf_counter()
    l2.zero_()
    bench.refresh(dfb, m, trk, i, d, c, calib, first)
    t1 = time.perf_counter()
    if i >= 5 and i % 20 != 0:
        rows.append([marks['fe0'] - t0, marks['fe1'] - marks['fe0'], marks['gn0'] - marks['fe1'], marks['gn1'] - marks['gn0'], t1 - marks['gn1'], t1 - t0])
r = np.median(np.array(rows), 0) * 1e6
print("median us/frame: pre-frontend %.0f | frontend enqueue %.0f | count read + prep until GN call %.0f | GN call %.0f | after GN %.0f | total %.0f" % tuple(r))
