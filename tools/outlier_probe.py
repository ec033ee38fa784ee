"""Where do the slow keyframes of the bench's first pass come from?  Host wall time of the pieces of integrate_keyframe."""
import importlib, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import bench
dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
calib = dfb.FrameIntrinsic(*dfb.synth.ICL_CALIB)
first_iso = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))
frames, _raw, seq = bench.gen_frames(dfb, 25, dev, seed=0)
l2 = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
T = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); T[name] = T.get(name, 0) + (time.perf_counter() - t) * 1e3; return r
    setattr(obj, name, g)
for rep in range(8):
    m, trk = bench.make_system(dfb, dev)
    wrap(m, "_inflate_latent_buffer"); wrap(m, "_workspace")
    for i, (d, c) in enumerate(frames):
        l2.zero_()
        tt0 = time.perf_counter()
        pose = trk.track_camera(c, d, calib, first_iso if i == 0 else None, depth_cut=(0.5, 5.0))
        tt1 = time.perf_counter()
        if i % 20 == 0:
            T.clear()
            pc, nrm = trk.last_processed_pc
            ta = time.perf_counter()
            a = pose @ pc; b = pose.rotation @ nrm
            print("rep", rep, "frame", i, "track host ms %.2f, pose @ cloud host ms %.2f" % ((tt1 - tt0) * 1e3, (time.perf_counter() - ta) * 1e3))
            t0 = time.perf_counter()
            m.integrate_keyframe(a, b, do_optimize=False)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            print("rep", rep, "frame", i, "integrate host ms %.2f (+sync %.2f)" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), {k: round(v, 2) for k, v in T.items()},
                  "cap", m.latent_vecs.size(0), "ws", m._ws.numel() >> 20, "MB")
