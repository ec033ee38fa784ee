"""-m gpu: pin the C-ABI ops against the REFERENCE'S OWN CUDA extensions (system/ext/*), built from /root/reference by
oracle/build_ref_ext.py into oracle/_ref/ext_build (git-ignored, shipped to the GPU box).  Skipped when absent."""
import importlib.util
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from util import ROOT, pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BUILD = ROOT / "oracle" / "_ref" / "ext_build"


def _load(name):
    so = BUILD / name / f"ref_{name}.so"
    if not so.exists():
        pytest.skip(f"{so} not built (reference extensions unavailable)")
    spec = importlib.util.spec_from_file_location(f"ref_{name}", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cloud():
    from oracle import ops
    synth = pkg().synth
    seq = synth.SyntheticSequence(n_frames=1, H=240, W=320)
    d, rgb = seq.frame(0)
    d[(d < 0.5) | (d > 5.0)] = float("nan")
    calib = tuple(c * 0.5 for c in synth.ICL_CALIB)
    return d, rgb, calib


def test_imgproc_vs_reference_cuda():
    ref = _load("imgproc")
    ext = pkg().ext
    d, rgb, calib = _cloud()
    dg = d.to(DEV)
    a = ext.unproject_depth(dg, *calib); b = ref.unproject_depth(dg, *calib)
    torch.cuda.synchronize()
    ok = ~torch.isnan(b[..., 0])
    assert torch.equal(torch.isnan(a[..., 0]), torch.isnan(b[..., 0]))
    assert torch.equal(a[ok], b[ok])                                          # bit-exact
    I = rgb.mean(-1).contiguous().to(DEV)
    ga, gb = ext.gradient_xy(I), ref.gradient_xy(I)
    torch.cuda.synchronize()
    assert torch.equal(torch.isnan(ga), torch.isnan(gb))
    assert torch.equal(ga[1:-1, 1:-1], gb[1:-1, 1:-1])
    synth = pkg().synth
    seq = synth.SyntheticSequence(n_frames=2, H=240, W=320)
    d1, c1 = seq.frame(1); d1[(d1 < 0.5) | (d1 > 5.0)] = float("nan")
    I1 = c1.mean(-1).contiguous().to(DEV); d1 = d1.to(DEV)
    g1 = ext.gradient_xy(I1)
    K = np.array([[calib[0], 0, calib[2]], [0, calib[1], calib[3]], [0, 0, 1.0]])
    from oracle.tracker_oracle import Pose
    dp = Pose.from_twist(np.array([0.004, -0.002, 0.003, 0.001, -0.002, 0.0015]))
    krk = (K @ dp.R @ np.linalg.inv(K)).flatten().tolist(); kt = (K @ dp.t).flatten().tolist()
    fa, Ja = ext.rgb_odometry(I, dg, I1, d1, g1, list(calib), krk, kt, 0.0, 0.2, True)
    fb, Jb = ref.rgb_odometry(I, dg, I1, d1, g1, list(calib), krk, kt, 0.0, 0.2, True)
    torch.cuda.synchronize()
    assert torch.equal(torch.isnan(fa), torch.isnan(fb))
    m = ~torch.isnan(fb)
    same = m & (fa == fb)
    frac = 1.0 - same.sum().item() / m.sum().item()
    print("rgb_odometry: fraction of valid pixels whose residual is not bit-identical:", frac)
    assert frac < 2e-3                         # warp-target rounding boundaries under different FMA contraction
    assert torch.allclose(Ja[same], Jb[same], rtol=1e-5, atol=1e-6)


def test_pcproc_vs_reference_cuda():
    ref = _load("pcproc")
    ext = pkg().ext
    d, _, calib = _cloud()
    pc = ext.unproject_depth(d.to(DEV), *calib)
    pc = torch.cat([pc, torch.zeros_like(pc[..., :1])], -1).reshape(-1, 4)
    pc = pc[~torch.isnan(pc[:, 0])].contiguous()
    ma = ext.remove_radius_outlier(pc, 16, 0.05); mb = ref.remove_radius_outlier(pc, 16, 0.05)
    torch.cuda.synchronize()
    assert (ma != mb).float().mean().item() < 1e-4
    pc2 = pc[mb].contiguous()
    na = ext.estimate_normals(pc2, 16, 0.1, [0.0, 0.0, 0.0]); nb = ref.estimate_normals(pc2, 16, 0.1, [0.0, 0.0, 0.0])
    torch.cuda.synchronize()
    assert (torch.isnan(na[:, 0]) != torch.isnan(nb[:, 0])).float().mean().item() < 1e-3
    ok = ~torch.isnan(na[:, 0]) & ~torch.isnan(nb[:, 0])
    err = (na[ok] - nb[ok]).abs().max(1).values
    assert torch.quantile(err, 0.99).item() < 2e-3
    assert err.median().item() < 1e-5


def test_indexing_vs_reference_cuda():
    ref = _load("indexing")
    ext = pkg().ext
    g = torch.Generator().manual_seed(0)
    vals = torch.randn(30000, 29, generator=g).to(DEV)
    idx = torch.randint(0, 900, (30000,), generator=g).to(DEV)
    sa, ca = ext.groupby_sum(vals, idx, 900); sb, cb = ref.groupby_sum(vals, idx, 900)
    torch.cuda.synchronize()
    assert torch.equal(ca, cb)
    assert torch.allclose(sa, sb, rtol=1e-4, atol=1e-4)


def test_marching_cubes_vs_reference_cuda(weights):
    ref = _load("marching_cubes")
    from util import GOLD, make_map, match_rows
    G = dict(np.load(GOLD / "map_golden.npz"))
    m = make_map(weights)
    Pw, Nw = torch.from_numpy(G["Pw"]).to(DEV), torch.from_numpy(G["Nw"]).to(DEV)
    m.integrate_keyframe(Pw, Nw)
    upd = m.mesh_cache.updated_vec_id
    focused = m.latent_vecs_pos[upd].contiguous()
    occ = m.indexer[m._expand_flatten_id(focused)]
    occ = occ[m.voxel_obs_count[occ] > 16.0]
    mapping = torch.full((int(occ.max().item()) + 1,), -1, device=DEV, dtype=torch.int)
    mapping[occ] = torch.arange(0, occ.numel(), device=DEV, dtype=torch.int)
    for r in (4, 8):
        cs, cd = m.decode_cubes(occ, r)
        ta, ia, sa = pkg().ext.marching_cubes_interp(m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(4e6), m.n_xyz, 0.15)
        tb, ib, sb = ref.marching_cubes_sparse_interp(m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(4e6), m.n_xyz, 0.15)
        torch.cuda.synchronize()
        assert ta.shape[0] == tb.shape[0] and ta.shape[0] > 100
        perm, _ = match_rows(ta.cpu().numpy(), tb.cpu().numpy(), 1e-5)
        assert np.array_equal(ia.cpu().numpy(), ib.cpu().numpy()[perm])
        assert np.abs(sa.cpu().numpy() - sb.cpu().numpy()[perm]).max() < 1e-5
