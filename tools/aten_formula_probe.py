"""Which rounding sequence do torch's CUDA kernels use for the three image ops of the tracker front end
(tracker.py:42-57, 84: torch.mean over the colour axis, bilinear align_corners=True half-size resampling)?
Candidates are evaluated with single IEEE operations (separate torch ops; fma emulated through float64, exact up to
double rounding) and compared bit for bit with the ATen result.  The fused front-end kernel (preprocess.cu
frame_images_kernel) is written with the sequence that matches.  Run on a GPU box."""
import itertools
import torch

dev = "cuda:0"
g = torch.Generator(device="cpu").manual_seed(5)
rgb = (torch.randint(0, 256, (480, 640, 3), generator=g).float() / 255.0).to(dev)


def fma(a, b, c):
    return (a.double() * b.double() + c.double()).float()


ref = torch.mean(rgb, dim=-1)
r, gg, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
third = torch.tensor(float(307200) / 921600, dtype=torch.float32, device=dev)
sums = {"(r+g)+b": (r + gg) + b, "(r+b)+g": (r + b) + gg, "r+(g+b)": r + (gg + b)}
for name, s in sums.items():
    for pname, v in (("*factor", s * third), ("/3", s / 3.0)):
        print(f"mean {name}{pname}: mismatches {(v != ref).sum().item()} of {ref.numel()}")

F = torch.nn.functional
for (H, W) in ((480, 640), (240, 320), (120, 160), (479, 641)):
    I = torch.rand((H, W), generator=g).to(dev)
    h2, w2 = H // 2, W // 2
    out = F.interpolate(I.view(1, 1, H, W), (h2, w2), mode="bilinear", align_corners=True)[0, 0]
    rh = torch.tensor((H - 1) / (h2 - 1), dtype=torch.float64).float().item() if h2 > 1 else 0.0
    rw = torch.tensor((W - 1) / (w2 - 1), dtype=torch.float64).float().item() if w2 > 1 else 0.0
    # float32 scale exactly as static_cast<float>(in - 1) / (out - 1): a float division
    rh = (torch.tensor(float(H - 1), dtype=torch.float32) / torch.tensor(float(h2 - 1), dtype=torch.float32)).to(dev)
    rw = (torch.tensor(float(W - 1), dtype=torch.float32) / torch.tensor(float(w2 - 1), dtype=torch.float32)).to(dev)
    hh = torch.arange(h2, device=dev, dtype=torch.float32) * rh
    ww = torch.arange(w2, device=dev, dtype=torch.float32) * rw
    h1 = hh.to(torch.int64); w1 = ww.to(torch.int64)
    h1p = (h1 < H - 1).to(torch.int64); w1p = (w1 < W - 1).to(torch.int64)
    l1h = (hh - h1.float())[:, None]; l0h = 1.0 - l1h
    l1w = (ww - w1.float())[None, :]; l0w = 1.0 - l1w
    a = I[h1][:, w1]; bq = I[h1][:, w1 + w1p]; c = I[h1 + h1p][:, w1]; d = I[h1 + h1p][:, w1 + w1p]

    def comb(x, p, y, q, mode):      # x*p + y*q under three contraction choices
        if mode == 0:
            return x * p + y * q
        if mode == 1:
            return fma(x, p, y * q)
        return fma(y, q, x * p)

    for mi, mo in itertools.product(range(3), range(3)):
        top = comb(l0w, a, l1w, bq, mi); bot = comb(l0w, c, l1w, d, mi)
        v = comb(l0h, top, l1h, bot, mo)
        bad = (v != out).sum().item()
        print(f"bilinear {H}x{W} inner {mi} outer {mo}: mismatches {bad} of {out.numel()}  max abs {float((v - out).abs().max()):.2e}")
    for (Hn, Wn) in ((H, W),):
        Dn = F.interpolate(I.view(1, 1, H, W), (h2, w2), mode="nearest")[0, 0]
        sh = torch.tensor(float(H), dtype=torch.float32) / torch.tensor(float(h2), dtype=torch.float32)
        sw = torch.tensor(float(W), dtype=torch.float32) / torch.tensor(float(w2), dtype=torch.float32)
        ih = torch.clamp(torch.floor(torch.arange(h2, dtype=torch.float32) * sh).long(), max=H - 1).to(dev)
        iw = torch.clamp(torch.floor(torch.arange(w2, dtype=torch.float32) * sw).long(), max=W - 1).to(dev)
        print(f"nearest {H}x{W}: mismatches {(I[ih][:, iw] != Dn).sum().item()}")
