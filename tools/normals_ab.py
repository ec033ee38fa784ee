"""A/B of the two estimate_normals kernels on a bench frame: run once with DFB_NORMALS_V1=1 and once without; the second run
compares its normals bit for bit with the file the first one left and prints both timings (CUDA events, 50 launches)."""
import importlib, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
seq = dfb.synth.SyntheticSequence(n_frames=3, device=dev, seed=0)
calib = dfb.synth.ICL_CALIB
out = {}
for f in range(3):
    depth, _ = seq.frame(f)
    depth[(depth < 0.5) | (depth > 5.0)] = float("nan")
    sub = depth[::2, ::2].contiguous()
    pc = dfb.ext.unproject_depth(sub, calib[0] * 0.5, calib[1] * 0.5, calib[2] * 0.5, calib[3] * 0.5)
    pc = torch.cat([pc, torch.zeros_like(pc[..., :1])], -1).reshape(-1, 4)
    pc = pc[~torch.isnan(pc[:, 0])].contiguous()
    pc = pc[dfb.ext.remove_radius_outlier(pc, 16, 0.05)].contiguous()
    nrm = dfb.ext.estimate_normals(pc, 16, 0.1, [0.0, 0.0, 0.0])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        dfb.ext.estimate_normals(pc, 16, 0.1, [0.0, 0.0, 0.0])
    b.record(); torch.cuda.synchronize()
    print(f"frame {f}: {pc.size(0)} points, estimate_normals (grid build + kNN + PCA) {a.elapsed_time(b) / 50 * 1e3:.1f} us, "
          f"NaN rows {int(torch.isnan(nrm[:, 0]).sum())}, V1={os.environ.get('DFB_NORMALS_V1')}")
    out[f"n{f}"] = nrm.cpu().numpy()
path = "gpurun_out/normals_ab.npz"
if os.path.exists(path):
    ref = np.load(path)
    for k in out:
        same = np.array_equal(ref[k].view(np.uint32), out[k].view(np.uint32))
        print(k, "bit-identical to the other kernel:", same, "" if same else f"max abs diff {np.nanmax(np.abs(ref[k] - out[k]))}, rows differing {(ref[k].view(np.uint32) != out[k].view(np.uint32)).any(1).sum()}")
else:
    np.savez(path, **out)
