"""Per-kernel micro-benchmarks on one B200 for BASELINE.json configs 3 (decoder sweep) and 4 (meshing) plus the
HBM-bound stages: CUDA-event timing, warm-up, inputs larger than L2 or L2 flushed between iterations, achieved
TFLOP/s or GB/s against MEASURED_PEAKS.json.  Writes a markdown table (profiles/).  Not the headline bench."""
import argparse, importlib, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from util import GOLD, MAPPING, make_map, ns, pkg

d = pkg(); lib = d._lib.load(); DEV = "cuda:0"
PK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush = torch.empty(192 << 20, dtype=torch.uint8, device=DEV)
rows = []


def timeit(fn, iters=10, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts))


def add(name, t, flop=None, nbytes=None, units=None, note=""):
    r = {"kernel": name, "time_us": round(t * 1e6, 1), "note": note}
    if flop:
        r["TFLOP/s"] = round(flop / t / 1e12, 2); r["frac"] = round(flop / t / 1e12 / PK["bf16_tflops"], 4); r["bound"] = "tensor"
    if nbytes:
        r["GB/s"] = round(nbytes / t / 1e9, 1); r["frac"] = round(nbytes / t / 1e9 / PK["hbm_gbs"], 4); r["bound"] = "hbm"
    if units:
        r["units/s"] = f"{units[0] / t:.3e} {units[1]}"
    rows.append(r); print(r, flush=True)


W = d.weights.load_npz(GOLD / "weights.npz")
# ---- synthetic 200k-voxel map (configs 3 and 4): a thick shell of voxels, latents ~ N(0, 0.1^2), obs_count = 100
args = dict(MAPPING); args.update(bound_min=[-6.4, -6.4, -6.4], bound_max=[6.4, 6.4, 6.4])
m = d.DenseIndexedMap(W, ns(args), 29, torch.device(DEV))
n = m.n_xyz[0]
g = torch.stack(torch.meshgrid(*[torch.arange(n, device=DEV)] * 3, indexing="ij"), -1).reshape(-1, 3)
rad = ((g.float() + 0.5 - n / 2) ** 2).sum(1).sqrt()
ids = torch.nonzero((rad > 48) & (rad < 56.5)).squeeze(-1)[:200000]
V = ids.numel()
m.cold_vars["latent_vecs"] = (torch.randn(V, 29, device=DEV) * 0.1); m.cold_vars["latent_vecs_pos"] = ids.clone()
m.cold_vars["voxel_obs_count"] = torch.full((V,), 100.0, device=DEV); m.cold_vars["voxel_optimized"] = torch.zeros(V, dtype=torch.bool, device=DEV)
m.indexer[ids] = torch.arange(V, device=DEV); m.cold_vars["n_occupied"] = V; m._reserve(V, V)        # adopt the hand-built tensors
print("synthetic map voxels:", V)
trk = d.SDFTracker(m, dict(iter_config=[], sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                           rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2)))
I = d.Isometry()
# ---- config 3: decoder sweep
for e in (16, 18, 20, 22, 24):
    M = 1 << e
    pick = ids[torch.randint(0, V, (M,), device=DEV)]
    xyz = (m._unlinearize_id(pick).float() + torch.rand(M, 3, device=DEV) * 0.98 + 0.01) * 0.1 + m.bound_min
    xyz = xyz.contiguous()
    for eng in (1, 0) if e <= 20 else (1,):
        lib.dfb_set_decoder_engine(eng)
        tag = "tcgen05" if eng else "fp32"
        t = timeit(lambda: m._get_sdf_raw(xyz, None, None), iters=5, do_flush=e < 22)
        add(f"get_sdf fwd 2^{e} [{tag}]", t, flop=M * 98816, units=(M, "queries"))
        t = timeit(lambda: d._lib.check(m.lib.dfb_sdf_hg(__import__('ctypes').byref(m._params), d.ext._p(xyz), M, d._lib.fptr([1,0,0,0,1,0,0,0,1,0,0,0]*2+[1,0,0,0,1,0,0,0,1]),
                                                          d.ext._p(m.indexer), d.ext._p(m.latent_vecs), d.ext._p(m.voxel_obs_count), d.ext._p(m.decoder_blob), 1, 5.0, 1,
                                                          d.ext._p(trk._hg_dev), d.ext._stream())), iters=5, do_flush=e < 22)
        add(f"sdf_hg fwd+bwd+JtJ 2^{e} [{tag}]", t, flop=M * 182528, units=(M, "queries"))
lib.dfb_set_decoder_engine(1)
# ---- config 4: meshing at r = 8 (200k voxels) and r = 16 (subset: cubes are 262 kB/voxel)
for r, nv in ((4, V), (8, V), (16, 40000)):
    occ = torch.arange(nv, device=DEV)
    R = 2 * r
    t = timeit(lambda: m.decode_cubes(occ, r), iters=3, warm=1, do_flush=False)
    add(f"decode_cubes r={r} ({nv} voxels)", t, flop=nv * r ** 3 * 98816, units=(nv, "voxels"), note="low-res pass only counted; refine band adds exact re-decodes")
    cs, cd = m.decode_cubes(occ, r)
    mapping = torch.full((V,), -1, dtype=torch.int32, device=DEV); mapping[:nv] = torch.arange(nv, dtype=torch.int32, device=DEV)
    blocks = ids[:nv].contiguous()
    out = {}
    def mc():
        out["t"] = d.ext.marching_cubes_interp(m.indexer.view(m.n_xyz), blocks, mapping, cs, cd, int(6e7), m.n_xyz, 2000.0)
    t = timeit(mc, iters=3, warm=1, do_flush=True)
    T = out["t"][0].shape[0]
    add(f"marching_cubes r={r} ({nv} voxels, {T} tris)", t, nbytes=nv * R ** 3 * 8 + nv * 216 + T * 56, units=(nv, "voxels"), note=f"{T / t:.3e} triangles/s (includes the count read-back)")
    del cs, cd
# ---- stage 1 / 2 on a 640x480 frame
seq = d.synth.SyntheticSequence(n_frames=1, device=DEV)
depth, rgb = seq.frame(0)
depth[(depth < 0.5) | (depth > 5.0)] = float("nan")
sub = depth[::2, ::2].contiguous()
t = timeit(lambda: d.ext.unproject_depth(depth, 481.2, 480.0, 319.5, 239.5)); add("unproject 640x480", t, nbytes=640 * 480 * 16)
pc = d.ext.unproject_depth(sub, 240.6, 240.0, 159.75, 119.75)
pc4 = torch.cat([pc, torch.zeros_like(pc[..., :1])], -1).reshape(-1, 4); pc4 = pc4[~torch.isnan(pc4[:, 0])].contiguous()
N = pc4.shape[0]
t = timeit(lambda: d.ext.remove_radius_outlier(pc4, 16, 0.05)); add(f"remove_radius_outlier ({N} pts; grid build + count)", t, nbytes=N * 36, units=(N, "points"))
t = timeit(lambda: d.ext.estimate_normals(pc4, 16, 0.1, [0, 0, 0])); add(f"estimate_normals ({N} pts)", t, nbytes=N * 48, units=(N, "points"))
nr = d.ext.estimate_normals(pc4, 16, 0.1, [0, 0, 0]); ok = ~torch.isnan(nr[:, 0]); p3 = pc4[ok, :3].contiguous(); nr = nr[ok].contiguous()
t = timeit(lambda: d.ext.point_box_filter(p3, nr, 0.02)); add(f"point_box_filter ({p3.shape[0]} pts)", t, nbytes=p3.shape[0] * 48, units=(p3.shape[0], "points"))
Iimg = rgb.mean(-1).contiguous()
t = timeit(lambda: d.ext.gradient_xy(Iimg)); add("gradient_xy 640x480", t, nbytes=640 * 480 * 12)
G = d.ext.gradient_xy(Iimg)
K = [481.2, 480.0, 319.5, 239.5]
t = timeit(lambda: d.ext.rgb_hg(Iimg, depth, Iimg, depth, G, K, [1, 0, 0, 0, 1, 0, 0, 0, 1], [0.001, 0, 0], 0.0, 0.2, 0, 0.01, True)); add("rgb_hg 640x480 (residual+J+reduce)", t, nbytes=640 * 480 * 28, units=(640 * 480, "pixels"))
# encoder engines on 2^20 samples
xs = torch.cat([torch.rand(1 << 20, 3, device=DEV) - 0.5, torch.randn(1 << 20, 3, device=DEV)], 1).contiguous()
eblob = torch.from_numpy(d.weights.pack_encoder(W)).to(DEV)
for eng in (0, 1):
    lib.dfb_set_encoder_engine(eng)
    t = timeit(lambda: d.ext.encoder_forward(xs, eblob), iters=5)
    add(f"encoder_forward 2^20 samples [{'tcgen05' if eng else 'fp32'}]", t, flop=(1 << 20) * 52096, units=(1 << 20, "samples"))
lib.dfb_set_encoder_engine(0)
# integrate (one keyframe, ~50k points) on a fresh default map
P, Nn = d.ext.point_box_filter(p3, nr, 0.02)
m2 = make_map(W)
R0 = torch.from_numpy(d.synth.quat_to_R(d.synth.FIRST_TQ[3:])).float().to(DEV); t0 = torch.tensor(d.synth.FIRST_TQ[:3], device=DEV)
Pw, Nw = (P @ R0.T + t0).contiguous(), (Nn @ R0.T).contiguous()
def integ():
    mm = make_map(W); mm.integrate_keyframe(Pw, Nw)
for eng in (0, 1):
    lib.dfb_set_encoder_engine(eng)
    maps = [make_map(W) for _ in range(8)]
    it = iter(maps)
    ti = timeit(lambda: next(it).integrate_keyframe(Pw, Nw), iters=4, warm=2)
    add(f"integrate_keyframe ({Pw.shape[0]} pts, first keyframe of a fresh map, incl. 1 host read + buffer growth) [{'tcgen05' if eng else 'fp32'} encoder]",
        ti, nbytes=Pw.shape[0] * 370, units=(Pw.shape[0], "points"))
    mm = maps[0]
    ti = timeit(lambda: mm.integrate_keyframe(Pw, Nw), iters=6, warm=2)
    add(f"integrate_keyframe ({Pw.shape[0]} pts, repeated on the same map) [{'tcgen05' if eng else 'fp32'} encoder]",
        ti, nbytes=Pw.shape[0] * 370, units=(Pw.shape[0], "points"))
lib.dfb_set_encoder_engine(1)
out = ROOT / (sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/kernel_table.md")     # copied to profiles/rNN_kernel_table.md
keys = ["kernel", "time_us", "bound", "TFLOP/s", "GB/s", "frac", "units/s", "note"]
with open(out, "w") as f:
    f.write(f"# Per-kernel micro-benchmarks (1x B200, CUDA events, median; peaks: HBM {PK['hbm_gbs']} GB/s, bf16 burst {PK['bf16_tflops']} TFLOP/s, measured)\n\n")
    f.write("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
    for r in rows:
        f.write("| " + " | ".join(str(r.get(k, "")) for k in keys) + " |\n")
print("wrote", out)
