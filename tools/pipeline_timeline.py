"""Where a frame's GPU time goes with and without frame pipelining: CUDA events around the front-end graph replay (side
stream when prefetched) and around the pose solve, per frame, on the bench sequence.  python tools/pipeline_timeline.py [frames]"""
import importlib
import sys

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import torch

import bench

dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
frames, _raw, seq = bench.gen_frames(dfb, N, dev, 0)
calib = dfb.FrameIntrinsic(*dfb.synth.ICL_CALIB)
first = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))


def run(pipe):
    m, trk = bench.make_system(dfb, dev)
    ev = []
    replay, solve = trk._fe_replay, trk._gauss_newton_native

    def timed_replay(ent, rgb, depth):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = replay(ent, rgb, depth); b.record()
        ev.append(("fe", a, b))
        return out

    def timed_solve(*args, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = solve(*args, **kw); b.record()
        ev.append(("solve", a, b))
        return out
    trk._fe_replay, trk._gauss_newton_native = timed_replay, timed_solve
    base = torch.cuda.Event(enable_timing=True)
    marks = []
    for i, (d, c) in enumerate(frames):
        if i == 5:
            torch.cuda.synchronize(); ev.clear(); base.record()
        nxt = frames[i + 1] if (pipe and i + 1 < N) else None
        bench.refresh(dfb, m, trk, i, d, c, calib, first, next_frame=nxt)
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    torch.cuda.synchronize()
    fe = [(base.elapsed_time(a), base.elapsed_time(b)) for k, a, b in ev if k == "fe"]
    so = [(base.elapsed_time(a), base.elapsed_time(b)) for k, a, b in ev if k == "solve"]
    total = base.elapsed_time(marks[-1]) / (N - 5)
    print(f"pipeline={pipe}: {total * 1e3:.0f} us/frame; front end {np.median([b - a for a, b in fe]) * 1e3:.0f} us (median), "
          f"solve {np.median([b - a for a, b in so]) * 1e3:.0f} us (median), evals/frame {trk.n_sdf_evals / N:.1f} sdf {trk.n_rgb_evals / N:.1f} rgb")
    for k in range(3, 8):
        print(f"   frame {k}: front end [{fe[k][0] * 1e3:7.0f} .. {fe[k][1] * 1e3:7.0f}]  solve [{so[k][0] * 1e3:7.0f} .. {so[k][1] * 1e3:7.0f}] us")


for _ in range(2):
    run(False)
    run(True)
