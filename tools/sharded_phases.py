"""Where a sharded keyframe's time goes on one GPU: CUDA events around the five phases (world 1, 1 M points, config-5 scene),
plus the encoder alone on the same number of samples (explicit rows, no scatter) for comparison."""
import importlib, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import numpy as np, torch
from util import GOLD, MAPPING, ns
from sharded_bench import scene_points
dfb = importlib.import_module("nerf-fusion_b200")
dev = "cuda:0"
W = dfb.weights.load_npz(GOLD / "weights.npz")
cfg = dict(MAPPING); cfg.update(bound_min=[-20.0, -2.0, -20.0], bound_max=[20.0, 8.0, 20.0], voxel_size=0.05)
fab = dfb.sharded.LocalFabric.create(W, ns(cfg), dev, 1, 1_000_000, 1 << 20)
m = fab.maps[0]
tot = np.zeros(5); n_s = 0
for k in range(6):
    P, N = scene_points(1_000_000, 100 + k, (-19.0 + 9.5 * (k % 4), -19.0), (-10.0 + 9.5 * (k % 4), -10.0), dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda.synchronize()
    for ph in range(5):
        ev[ph].record(); m.phase(ph + 1, P, N)
    ev[5].record(); torch.cuda.synchronize()
    st = m.read_stats()
    if k >= 2:
        tot += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(5)]); n_s = st["samples_in"]
tot /= 4
x = torch.cat([torch.rand(n_s, 3, device=dev) - 0.5, torch.randn(n_s, 3, device=dev)], 1).contiguous()
blob = torch.from_numpy(dfb.weights.pack_encoder(W)).to(dev)
for _ in range(3):
    dfb.ext.encoder_forward(x, blob)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    dfb.ext.encoder_forward(x, blob)
b.record(); torch.cuda.synchronize()
enc_ms = a.elapsed_time(b) / 5
print(json.dumps({"points": 1_000_000, "samples": int(n_s), "phase_ms": {f"phase{i + 1}": round(float(t), 3) for i, t in enumerate(tot)},
                  "keyframe_ms": round(float(tot.sum()), 3), "encoder_explicit_ms_same_samples": round(enc_ms, 3),
                  "encoder_Gsamples_per_s_explicit": round(n_s / enc_ms / 1e6, 3), "encoder_TFLOPs_algorithmic": round(n_s * 52096 / enc_ms / 1e9, 1)}))
