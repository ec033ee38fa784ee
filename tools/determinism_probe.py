"""Is a tracked sequence reproducible run to run?  Same frames, fresh map + tracker each time; prints the largest pose
difference per frame between pairs of runs (non-pipelined twice, pipelined twice, one against the other)."""
import importlib, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
N = 12
frames, _raw, seq = bench.gen_frames(dfb, N, dev, 0)
calib = dfb.FrameIntrinsic(*dfb.synth.ICL_CALIB)
first = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))


def run(pipe):
    m, trk = bench.make_system(dfb, dev)
    out = []
    for i, (d, c) in enumerate(frames):
        nxt = frames[i + 1] if (pipe and i + 1 < N) else None
        p = bench.refresh(dfb, m, trk, i, d, c, calib, first, next_frame=nxt)
        out.append(np.concatenate([p.q.rotation_matrix.reshape(-1), p.t]))
    torch.cuda.synchronize()
    return np.array(out), (trk.n_sdf_evals, trk.n_rgb_evals)


a, ea = run(False); b, eb = run(False); c, ec = run(True); d, ed = run(True)
f = lambda x, y: " ".join("%.0e" % v for v in np.abs(x - y).max(1))
print("evals", ea, eb, ec, ed)
print("plain vs plain:    ", f(a, b))
print("pipelined vs same: ", f(c, d))
print("plain vs pipelined:", f(a, c))
