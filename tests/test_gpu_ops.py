"""-m gpu: operator-level parity of the C-ABI ops (called through nerf-fusion_b200.ext) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import nets, ops
from util import GOLD, pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _engines(use_engine):
    """Every test of this module runs under both engine configurations (tests/conftest.py): the default tcgen05 engines that
    bench.py measures, and the FP32 CUDA-core pair."""
    yield use_engine


def _depth_frame(H=240, W=320, seed=0):
    synth = pkg().synth
    seq = synth.SyntheticSequence(n_frames=1, H=H, W=W, seed=seed)
    d, rgb = seq.frame(0)
    d = d.clone(); d[(d < 0.5) | (d > 5.0)] = float("nan")
    return d, rgb, tuple(c * H / 480.0 for c in synth.ICL_CALIB)


def _cloud(n_max=None):
    d, _, calib = _depth_frame()
    pc = ops.unproject_depth(d.numpy(), *calib).reshape(-1, 3)
    pc = pc[~np.isnan(pc[:, 0])]
    if n_max:
        pc = pc[:n_max]
    return np.concatenate([pc, np.zeros((pc.shape[0], 1), np.float32)], 1)


def test_unproject_depth_bit_exact():
    ext = pkg().ext
    d, _, calib = _depth_frame()
    got = ext.unproject_depth(d.to(DEV), *calib).cpu().numpy()
    ref = ops.unproject_depth(d.numpy(), *calib)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    # empty input
    assert ext.unproject_depth(torch.zeros((0, 8), device=DEV), 1, 1, 0, 0).shape == (0, 8, 3)


def test_unproject_rejects_cpu_and_noncontiguous():
    ext = pkg().ext
    with pytest.raises(RuntimeError):
        ext.unproject_depth(torch.zeros(4, 4), 1, 1, 0, 0)
    with pytest.raises(RuntimeError):
        ext.unproject_depth(torch.zeros(4, 8, device=DEV)[:, ::2], 1, 1, 0, 0)


def test_remove_radius_outlier():
    ext = pkg().ext
    pc4 = _cloud()
    got = ext.remove_radius_outlier(torch.from_numpy(pc4).to(DEV), 16, 0.05).cpu().numpy()
    dist, _ = ops.knn(pc4, 16, max_radius=0.05)
    ref = dist[:, 15] < np.float32(0.05) * np.float32(0.05)
    r2 = np.float32(0.05) * np.float32(0.05)
    border = np.abs(dist[:, 15] - r2) <= 4e-7 * r2            # FMA-contraction ulp band
    assert np.array_equal(got[~border], ref[~border])
    assert border.mean() < 1e-3
    assert 0.05 < got.mean() < 1.0
    # ragged / tiny inputs
    assert ext.remove_radius_outlier(torch.zeros((0, 4), device=DEV), 16, 0.05).numel() == 0
    few = torch.from_numpy(pc4[:5]).to(DEV)
    assert not ext.remove_radius_outlier(few, 16, 0.05).any()


def test_estimate_normals():
    ext = pkg().ext
    pc4 = _cloud()
    keep = ops.remove_radius_outlier(pc4, 16, 0.05)
    pc4 = pc4[keep]
    got = ext.estimate_normals(torch.from_numpy(pc4).to(DEV), 16, 0.1, [0.0, 0.0, 0.0]).cpu().numpy()
    ref = ops.estimate_normals(pc4, 16, 0.1, [0.0, 0.0, 0.0])
    nan_g, nan_r = np.isnan(got[:, 0]), np.isnan(ref[:, 0])
    assert (nan_g != nan_r).mean() < 1e-3
    ok = ~nan_g & ~nan_r
    err = np.abs(got[ok] - ref[ok]).max(1)
    # ties in the 16-NN set / eigen-solver conditioning: allow a tiny tail, demand tight agreement elsewhere
    assert np.quantile(err, 0.99) < 2e-3, np.quantile(err, [0.5, 0.9, 0.99, 1.0])
    assert np.median(err) < 1e-5
    assert abs(np.linalg.norm(got[ok], axis=1).mean() - 1.0) < 1e-3
    assert ((got[ok] * pc4[ok, :3]).sum(1) <= 1e-6).all()       # oriented towards the camera at the origin


def test_point_box_filter_bit_exact():
    ext = pkg().ext
    pc4 = _cloud()
    rng = np.random.RandomState(0)
    nrm = rng.randn(pc4.shape[0], 3).astype(np.float32)
    gp, gn = ext.point_box_filter(torch.from_numpy(pc4[:, :3].copy()).to(DEV), torch.from_numpy(nrm).to(DEV), 0.02)
    rp, rn, _ = ops.point_box_filter(pc4[:, :3], nrm, 0.02)
    assert gp.shape[0] == rp.shape[0]
    assert np.array_equal(gp.cpu().numpy(), rp)
    assert np.array_equal(gn.cpu().numpy(), rn)
    gp2, _ = ext.point_box_filter(torch.from_numpy(pc4[:, :3].copy()).to(DEV), torch.from_numpy(nrm).to(DEV), 0.02, 1)
    rp2, _, _ = ops.point_box_filter(pc4[:, :3], nrm, 0.02, divide="recip")
    assert np.array_equal(gp2.cpu().numpy(), rp2)


def test_scatter_mean_and_groupby_sum():
    ext = pkg().ext
    rng = np.random.RandomState(3)
    n, C = 20000, 700
    idx = rng.randint(0, C, size=n).astype(np.int64)
    idx[0] = C - 1
    src = rng.randn(n, 3).astype(np.float32)
    got = ext.scatter_mean(torch.from_numpy(src).to(DEV), torch.from_numpy(idx).to(DEV), dim=0).cpu().numpy()
    assert np.array_equal(got, ops.scatter_mean(src, idx))
    src7 = rng.randn(n, 7).astype(np.float32)
    got7 = ext.scatter_mean(torch.from_numpy(src7).to(DEV), torch.from_numpy(idx).to(DEV)).cpu().numpy()
    assert np.array_equal(got7, ops.scatter_mean(src7, idx))
    vals = rng.randn(n, 29).astype(np.float32)
    s, c = ext.groupby_sum(torch.from_numpy(vals).to(DEV), torch.from_numpy(idx).to(DEV), C)
    rs, rc = ops.groupby_sum(vals, idx, C)
    assert np.array_equal(c.cpu().numpy(), rc)
    np.testing.assert_allclose(s.cpu().numpy(), rs, rtol=1e-4, atol=1e-4)
    s0, c0 = ext.groupby_sum(torch.zeros((0, 29), device=DEV), torch.zeros((0,), dtype=torch.long, device=DEV), 5)
    assert s0.abs().sum() == 0 and c0.sum() == 0


def test_gradient_xy_and_rgb_odometry():
    ext = pkg().ext
    synth = pkg().synth
    seq = synth.SyntheticSequence(n_frames=2, H=240, W=320)
    (d0, c0), (d1, c1) = seq.frame(0), seq.frame(1)
    for d in (d0, d1):
        d[(d < 0.5) | (d > 5.0)] = float("nan")
    I0, I1 = c0.mean(-1).contiguous(), c1.mean(-1).contiguous()
    g = ext.gradient_xy(I1.to(DEV)).cpu().numpy()
    gr = ops.gradient_xy(I1.numpy())
    assert np.array_equal(np.isnan(g), np.isnan(gr))
    np.testing.assert_allclose(g[1:-1, 1:-1], gr[1:-1, 1:-1], rtol=0, atol=1e-7)
    calib = [c * 0.5 for c in synth.ICL_CALIB]
    K = np.array([[calib[0], 0, calib[2]], [0, calib[1], calib[3]], [0, 0, 1.0]])
    from oracle.tracker_oracle import Pose
    dp = Pose.from_twist(np.array([0.004, -0.002, 0.003, 0.001, -0.002, 0.0015]))
    krk = (K @ dp.R @ np.linalg.inv(K)).flatten().tolist(); kt = (K @ dp.t).flatten().tolist()
    f, J = ext.rgb_odometry(I0.to(DEV), d0.to(DEV), I1.to(DEV), d1.to(DEV), torch.from_numpy(gr).to(DEV), calib, krk, kt, 0.0, 0.2, True)
    fr, Jr = ops.rgb_odometry(I0.numpy(), d0.numpy(), I1.numpy(), d1.numpy(), gr, calib, krk, kt, 0.0, 0.2, True)
    f = f.cpu().numpy(); J = J.cpu().numpy()
    both = ~np.isnan(f) & ~np.isnan(fr)
    assert (np.isnan(f) != np.isnan(fr)).mean() < 2e-3          # rounding-boundary pixels (see oracle/ops.py)
    same = both & (np.abs(f - fr) < 1e-6)
    assert same.sum() > 0.995 * both.sum()
    np.testing.assert_allclose(J[same], Jr[same], rtol=2e-4, atol=1e-4)
    # fused reduction == reduction of the per-pixel outputs
    out = ext.rgb_hg(I0.to(DEV), d0.to(DEV), I1.to(DEV), d1.to(DEV), torch.from_numpy(gr).to(DEV), calib, krk, kt, 0.0, 0.2, 0, 0.01, True)
    out = out.cpu().numpy()
    m = ~np.isnan(f)
    Jm = -J[m].astype(np.float64); fm = f[m].astype(np.float64)
    # per-thread partial sums are FP32: the error scales with the sum of magnitudes, not with the (cancelling) sum; and
    # the two kernels inline the per-pixel routine separately (different FMA contraction), so up to `flips` pixels on a
    # validity boundary (depth-delta / image-border tests) may be counted by one and not the other
    flips = 4
    aJ, af = np.abs(Jm), np.abs(fm)
    assert (np.abs(out[:36].reshape(6, 6) - Jm.T @ Jm) <= 2e-6 * (aJ.T @ aJ) + flips * (aJ[:, :, None] * aJ[:, None, :]).max(0)).all()
    assert (np.abs(out[36:42] - Jm.T @ fm) <= 2e-6 * (aJ.T @ af) + flips * (aJ * af[:, None]).max(0)).all()
    assert abs(out[42] - (fm * fm).sum()) <= 1e-5 * (fm * fm).sum() + flips * (fm * fm).max()
    assert abs(out[43] - m.sum()) <= flips


def test_encoder_and_decoder_forward(weights):
    d = pkg()
    rng = np.random.RandomState(5)
    x = np.concatenate([rng.rand(5000, 3) - 0.5, rng.randn(5000, 3)], 1).astype(np.float32)
    blob = torch.from_numpy(d.weights.pack_encoder(weights)).to(DEV)
    got = d.ext.encoder_forward(torch.from_numpy(x).to(DEV), blob).cpu().numpy()
    ref = nets.encoder_forward(weights, torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5)
    xq = np.concatenate([rng.randn(3001, 29) * 0.1, rng.rand(3001, 3) - 0.5], 1).astype(np.float32)
    dblob = torch.from_numpy(d.weights.pack_decoder(weights)).to(DEV)
    sdf, std = d.ext.decoder_forward(torch.from_numpy(xq).to(DEV), dblob)
    rs, rd = nets.decoder_forward(weights, torch.from_numpy(xq))
    np.testing.assert_allclose(sdf.cpu().numpy(), rs.numpy(), atol=2e-6)       # SDF tolerance: 1e-4 m = 1e-3 voxel units
    np.testing.assert_allclose(std.cpu().numpy(), rd.numpy(), atol=2e-6)
    s0, _ = d.ext.decoder_forward(torch.zeros((0, 32), device=DEV), dblob)
    assert s0.numel() == 0


def test_ingest_frame_and_reader(tmp_path):
    """dfb_ingest_frame against the reference reader's outputs (IEEE mode = torch CPU, bit-exact), the reciprocal mode against
    the oracle (what torch CUDA computes for tensor / scalar), and the whole ICLNUIMSequence reader on a PNG sequence."""
    from test_dataset_cpu import _write_sequence
    from oracle import ops as O
    d = pkg()
    D = dict(np.load(GOLD / "dataset_golden.npz"))
    for i in range(int(D["n_frames"])):
        d16 = torch.from_numpy(D[f"f{i}_depth_u16"].view(np.int16)).view(torch.uint16).to(DEV)
        c8 = torch.from_numpy(D[f"f{i}_bgr_u8"]).to(DEV)
        dep, rgb = d.ext.ingest_frame(d16, c8, 5000.0, None, bgr=True, div_mode=0)
        assert np.array_equal(dep.cpu().numpy(), D[f"f{i}_depth"]) and np.array_equal(rgb.cpu().numpy(), D[f"f{i}_rgb"])
        dep1, rgb1 = d.ext.ingest_frame(d16, c8, 5000.0, (0.5, 5.0), bgr=True, div_mode=1)
        od, oc = O.ingest_frame(D[f"f{i}_depth_u16"], D[f"f{i}_bgr_u8"], 5000.0, (0.5, 5.0), bgr=True, recip=True)
        assert np.array_equal(np.nan_to_num(dep1.cpu().numpy(), nan=-1.0), np.nan_to_num(od, nan=-1.0))
        assert np.array_equal(rgb1.cpu().numpy(), oc)
        only_d, none = d.ext.ingest_frame(d16, None)
        assert none is None and only_d.shape == d16.shape
    pytest.importorskip("cv2")
    _write_sequence(D, tmp_path)
    seq = d.dataset.ICLNUIMSequence(str(tmp_path), 0, -1, D["first_tq"].tolist(), load_gt=True, device=DEV)
    assert len(seq) == int(D["n_frames"])
    for i, fd in enumerate(seq):
        od, oc = O.ingest_frame(D[f"f{i}_depth_u16"], D[f"f{i}_bgr_u8"], 5000.0, None, bgr=True, recip=True)
        assert np.array_equal(fd.depth.cpu().numpy(), od) and np.array_equal(fd.rgb.cpu().numpy(), oc)
        assert np.abs(fd.depth.cpu().numpy() - D[f"f{i}_depth"]).max() < 1e-6          # reciprocal vs IEEE: 1 ulp
        assert np.abs(fd.gt_pose.t - D[f"f{i}_gt_t"]).max() < 1e-12 and fd.calib.dscale == 5000.0


def test_transform_points_equals_torch():
    """Isometry @ cloud through the small kernel == the reference's fp32 `other @ R^T + t` (motion_util.py:323-328) to an ulp
    or two (cuBLAS may order the three products differently); rotation-only and empty inputs."""
    d = pkg()
    rng = np.random.RandomState(3)
    iso = d.Isometry(q=d.Quaternion(array=[0.3, -0.5, 0.7, 0.2]), t=np.array([1.5, -2.0, 0.25]))
    x = torch.from_numpy((rng.rand(5001, 3) * 6 - 3).astype(np.float32)).to(DEV)
    R, t = iso.torch_matrices(DEV)
    ref = (x.cpu() @ R.cpu().t() + t.cpu().unsqueeze(0)).numpy()
    out = (iso @ x).cpu().numpy()
    assert np.abs(out - ref).max() <= 4e-7 * np.abs(ref).max()
    rot = (iso.rotation @ x).cpu().numpy()
    assert np.abs(rot - (x.cpu() @ R.cpu().t()).numpy()).max() <= 4e-7 * 6
    assert (iso @ x[:0]).shape == (0, 3)


def test_latent_adam_step_vs_torch(weights):
    """dfb_latent_adam_step (latent optimiser, map.py:81-113): gradient of the Gaussian log-likelihood w.r.t. the 29 latent inputs
    + torch.optim.Adam's update, against torch autograd on the oracle networks (CPU) with the same loss and regulariser."""
    from util import make_map
    d = pkg()
    m = make_map(weights)
    torch.manual_seed(0)
    U, n = 50, 4000
    lat = torch.randn(U, 29) * 0.1
    inv = torch.randint(0, U, (n,))
    xyz = torch.rand(n, 3) - 0.5
    gt = torch.randn(n) * 0.05
    lam = 1e-4

    def torch_run(iters):
        x = lat.clone().requires_grad_(True)
        opt = torch.optim.Adam([x], lr=1e-2)
        g1 = None
        for _ in range(iters):
            opt.zero_grad()
            s, sd = nets.decoder_forward(weights, torch.cat([x[inv], xyz], 1))
            ll = -torch.distributions.Normal(loc=torch.clamp(s.squeeze(-1), -0.2, 0.2), scale=sd.squeeze(-1)).log_prob(torch.clamp(gt, -0.2, 0.2))
            (ll.sum() / n + lam * torch.sum(torch.norm(x, dim=1)) / n).backward()
            g1 = x.grad.clone() if g1 is None else g1
            opt.step()
        return x.detach(), g1
    m.args.code_regularization = True; m.args.code_reg_lambda = lam
    # one step is exact to rounding; later steps divide by sqrt(v) of gradients around 1e-6, which amplifies last-bit differences
    for iters, tol in ((1, 1e-6), (3, 5e-3)):
        ref, g1 = torch_run(iters)
        m.args.optim_n_iters = iters
        out = m.optimize_latents(lat.to(DEV), inv.to(DEV), gt.to(DEV), xyz.to(DEV)).cpu()
        err = (out - ref).abs().max().item()
        print(f"{iters} Adam step(s): max |ours - torch| {err:.2e}, moved {(out - lat).abs().max().item():.3e}")
        assert err < tol
    # the gradient itself: first moment after one step with lr = 0 is 0.1 * grad
    lat_d, inv_d, xyz_d, gt_d = lat.to(DEV).contiguous(), inv.to(DEV), xyz.to(DEV).contiguous(), gt.to(DEV)
    grad, m1, m2 = torch.zeros_like(lat_d), torch.zeros_like(lat_d), torch.zeros_like(lat_d)
    p = d.ext._p
    d._lib.check(m.lib.dfb_latent_adam_step(p(lat_d), U, p(inv_d), p(xyz_d), p(gt_d), n, p(m.decoder_blob), p(grad), p(m1), p(m2), 1, 0.0,
                                            lam / n, d.ext._stream()))
    g = (m1 / 0.1).cpu()
    assert (g - g1).abs().max() <= 1e-5 * g1.abs().max()
    assert float(grad.abs().sum()) == 0.0                          # left zero for the next iteration


def test_frame_images_equals_torch():
    """dfb_frame_images (depth clipping, intensity, 3-level pyramid, gradients: two launches) against the torch CUDA kernels
    the reference calls (tracker.py:42-57, 84; main.py:56-57): bit-identical, also for sizes that do not halve evenly."""
    d = pkg()
    F = torch.nn.functional
    g = torch.Generator(device="cpu").manual_seed(11)
    for (H, W) in ((480, 640), (240, 320), (479, 641), (62, 90)):
        rgb = (torch.randint(0, 256, (H, W, 3), generator=g).float() / 255.0).to(DEV)
        depth = (torch.rand((H, W), generator=g) * 6.0).to(DEV)
        depth[torch.rand((H, W), generator=g).to(DEV) < 0.1] = float("nan")
        for cut in (None, (0.5, 5.0)):
            Is, Ds, Gs = d.ext.frame_images(rgb, depth, cut)
            dd = depth if cut is None else torch.where((depth < cut[0]) | (depth > cut[1]), torch.full_like(depth, float("nan")), depth)
            i0 = torch.mean(rgb, dim=-1).view(1, 1, H, W)
            d0 = dd.view(1, 1, H, W)
            i1 = F.interpolate(i0, (H // 2, W // 2), mode="bilinear", align_corners=True)
            d1 = F.interpolate(d0, (H // 2, W // 2), mode="nearest")
            i2 = F.interpolate(i1, (H // 2 // 2, W // 2 // 2), mode="bilinear", align_corners=True)
            d2 = F.interpolate(d1, (H // 2 // 2, W // 2 // 2), mode="nearest")
            ref_I = [t[0, 0].contiguous() for t in (i0, i1, i2)]
            ref_D = [t[0, 0].contiguous() for t in (d0, d1, d2)]
            for a, b in zip(Is + Ds, ref_I + ref_D):
                assert a.shape == b.shape
                assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)), (H, W, cut)
            for a, b in zip(Gs, ref_I):
                assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(d.ext.gradient_xy(b), nan=-1.0))
