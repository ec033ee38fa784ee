"""Where does a frame go?  Wall-clock per phase with synchronisation (diagnostic, not a benchmark)."""
import sys, time, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
frames, _raw, seq = bench.gen_frames(dfb, 25, dev, 0)
m, trk = bench.make_system(dfb, dev)
calib = dfb.FrameIntrinsic(*dfb.synth.ICL_CALIB)
first = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))
T = {}
def tic(): torch.cuda.synchronize(); return time.perf_counter()
def add(k, t0): torch.cuda.synchronize(); T[k] = T.get(k, 0) + time.perf_counter() - t0
orig_sdf, orig_rgb = trk.compute_sdf_Hg, trk.compute_rgb_Hg
def sdf(*a, **k):
    t0 = time.perf_counter(); r = orig_sdf(*a, **k); T['gn_sdf'] = T.get('gn_sdf', 0) + time.perf_counter() - t0; T['n_sdf'] = T.get('n_sdf', 0) + 1; return r
def rgb(*a, **k):
    t0 = time.perf_counter(); r = orig_rgb(*a, **k); T['gn_rgb'] = T.get('gn_rgb', 0) + time.perf_counter() - t0; T['n_rgb'] = T.get('n_rgb', 0) + 1; return r
trk.compute_sdf_Hg, trk.compute_rgb_Hg = sdf, rgb
for i, (d, c) in enumerate(frames):
    if i == 5: T.clear()
    t0 = tic()
    d = torch.where((d < 0.5) | (d > 5.0), torch.full_like(d, float('nan')), d); add('depth_cut', t0)
    t0 = tic()
    Is, Ds, Gs, out_p, out_n, cnt = trk._frontend_graphed(c.contiguous(), d.contiguous(), calib)
    mcount = int(cnt.item()); pc, nrm = out_p[:mcount].clone(), out_n[:mcount].clone(); add('frontend (graph: pyramid + preprocess)', t0)
    trk.last_processed_pc = [pc, nrm]
    t0 = tic()
    if i == 0: pose = first
    else: pose = trk.gauss_newton(trk.all_pd_pose[-1].dot(dfb.Isometry()), Is, Ds, Gs, pc, calib)
    add('gauss_newton', t0)
    trk.last_intensity, trk.last_depth = Is, Ds; trk.all_pd_pose.append(pose)
    if i % 20 == 0:
        t0 = tic(); m.integrate_keyframe(pose @ pc, pose.rotation @ nrm); add('integrate', t0)
n = len(frames) - 5
print({k: (round(1e3 * v / n, 3) if not k.startswith('n_') else round(v / n, 1)) for k, v in T.items()}, "ms/frame")
# finer: preprocessing steps
d, c = frames[10]
d = torch.where((d < 0.5) | (d > 5.0), torch.full_like(d, float('nan')), d)
ext = dfb.ext
t0 = tic(); sub = torch.nn.functional.interpolate(d[None, None], scale_factor=0.5, mode="nearest", recompute_scale_factor=False)[0, 0].contiguous(); add('p_sub', t0)
t0 = tic(); pc = ext.unproject_depth(sub, 240.6, 240.0, 159.75, 119.75); add('p_unproject', t0)
t0 = tic(); pc4 = torch.cat([pc, torch.zeros((240, 320, 1), device=dev)], -1).reshape(-1, 4); pc4 = pc4[~torch.isnan(pc4[..., 0])].contiguous(); add('p_compact', t0)
t0 = tic(); mk = ext.remove_radius_outlier(pc4, 16, 0.05); add('p_outlier', t0)
t0 = tic(); pc4 = pc4[mk].contiguous(); add('p_compact2', t0)
t0 = tic(); nr = ext.estimate_normals(pc4, 16, 0.1, [0, 0, 0]); add('p_normals', t0)
t0 = tic(); ok = ~torch.isnan(nr[..., 0]); nr = nr[ok].contiguous(); p3 = pc4[ok, :3].contiguous(); add('p_compact3', t0)
t0 = tic(); a, b = ext.point_box_filter(p3, nr, 0.02); add('p_boxfilter', t0)
print({k: round(1e3 * v, 3) for k, v in T.items() if k.startswith('p_')}, "ms (single call, sync'd)")
