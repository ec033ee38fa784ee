"""BASELINE config 3: decoder SDF / Jacobian query sweep 2^16 .. 2^24 on one B200 (CUDA events, median of 5, warm-up 3,
L2 flushed below 2^22, inputs larger than L2 above).  Prints one JSON line per size; `--out` writes a markdown table."""
import argparse, ctypes, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from util import GOLD, MAPPING, ns, pkg

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--max", type=int, default=24)
ap.add_argument("--fp32", action="store_true", help="also time the FP32 CUDA-core engine up to 2^20")
a = ap.parse_args()
d = pkg(); lib = d._lib.load(); DEV = "cuda:0"
PK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"bf16_tflops": 1590.0}
flush = torch.empty(192 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=5, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts))


W = d.weights.load_npz(GOLD / "weights.npz")
args = dict(MAPPING); args.update(bound_min=[-6.4, -6.4, -6.4], bound_max=[6.4, 6.4, 6.4])
m = d.DenseIndexedMap(W, ns(args), 29, torch.device(DEV))
n = m.n_xyz[0]
g = torch.stack(torch.meshgrid(*[torch.arange(n, device=DEV)] * 3, indexing="ij"), -1).reshape(-1, 3)
rad = ((g.float() + 0.5 - n / 2) ** 2).sum(1).sqrt()
ids = torch.nonzero((rad > 48) & (rad < 56.5)).squeeze(-1)[:200000]
V = ids.numel()
m.cold_vars["latent_vecs"] = (torch.randn(V, 29, device=DEV) * 0.1); m.cold_vars["latent_vecs_pos"] = ids.clone()
m.cold_vars["voxel_obs_count"] = torch.full((V,), 100.0, device=DEV); m.cold_vars["voxel_optimized"] = torch.zeros(V, dtype=torch.bool, device=DEV)
m.indexer[ids] = torch.arange(V, device=DEV); m.cold_vars["n_occupied"] = V; m._reserve(V, V)
hg = torch.zeros(80, dtype=torch.float64, device=DEV)
pose = d._lib.fptr([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0] * 2 + [1, 0, 0, 0, 1, 0, 0, 0, 1])
rows = []
for e in range(16, a.max + 1, 2):
    M = 1 << e
    pick = ids[torch.randint(0, V, (M,), device=DEV)]
    xyz = ((m._unlinearize_id(pick).float() + torch.rand(M, 3, device=DEV) * 0.98 + 0.01) * 0.1 + m.bound_min).contiguous()
    for eng in ((1, 0) if (a.fp32 and e <= 20) else (1,)):
        lib.dfb_set_decoder_engine(eng)
        t_f = timeit(lambda: m._get_sdf_raw(xyz, None, None), do_flush=e < 22)
        t_b = timeit(lambda: d._lib.check(m.lib.dfb_sdf_hg(ctypes.byref(m._params), d.ext._p(xyz), M, pose, d.ext._p(m.indexer), d.ext._p(m.latent_vecs),
                                                           d.ext._p(m.voxel_obs_count), d.ext._p(m.decoder_blob), 1, 5.0, 1, d.ext._p(hg), d.ext._stream())),
                     do_flush=e < 22)
        r = {"queries": f"2^{e}", "engine": "tcgen05 fp16x3" if eng else "fp32", "fwd_us": round(t_f * 1e6, 1), "fwd_Gq/s": round(M / t_f / 1e9, 3),
             "fwd_TFLOP/s": round(M * 98816 / t_f / 1e12, 1), "fwdbwd_us": round(t_b * 1e6, 1), "fwdbwd_Gq/s": round(M / t_b / 1e9, 3),
             "fwdbwd_TFLOP/s": round(M * 182528 / t_b / 1e12, 1), "frac_of_bf16_burst": round(M * 182528 / t_b / 1e12 / PK["bf16_tflops"], 4)}
        rows.append(r); print(json.dumps(r), flush=True)
lib.dfb_set_decoder_engine(1)
if a.out:
    keys = list(rows[0].keys())
    with open(a.out, "w") as f:
        f.write(f"# Decoder query sweep (BASELINE config 3), 1x B200, CUDA events, median of 5; algorithmic FLOPs 98816 (fwd) / 182528 (fwd+bwd) per query; "
                f"peak = measured bf16 burst {PK['bf16_tflops']} TFLOP/s\n\n| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
        for r in rows:
            f.write("| " + " | ".join(str(r[k]) for k in keys) + " |\n")
