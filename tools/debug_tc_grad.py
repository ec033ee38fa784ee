import sys, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from util import GOLD, make_map, pkg
d = pkg(); lib = d._lib.load()
W = d.weights.load_npz(GOLD / "weights.npz")
G = dict(np.load(GOLD / "map_golden.npz"))
DEV = "cuda:0"
m = make_map(W)
Pw, Nw = torch.from_numpy(G["Pw"]).to(DEV), torch.from_numpy(G["Nw"]).to(DEV)
m.integrate_keyframe(Pw, Nw)
xyz = torch.from_numpy(G["q_world"][:4000]).to(DEV)
n = xyz.size(0)
for name, gs, gd in (("sdf only", 1.0, 0.0), ("std only", 0.0, 1.0), ("both", 1.0, 0.3)):
    res = {}
    for eng in (0, 1):
        lib.dfb_set_decoder_engine(eng)
        g = m._get_sdf_raw(xyz, torch.full((n,), gs, device=DEV), torch.full((n,), gd, device=DEV))
        s, sd, v = m._get_sdf_raw(xyz, None, None)
        torch.cuda.synchronize()
        res[eng] = (g.cpu().numpy(), s.cpu().numpy(), sd.cpu().numpy(), v.cpu().numpy())
    g0, g1 = res[0][0], res[1][0]
    err = np.abs(g1 - g0).max(1); mag = np.abs(g0).max(1) + 1e-9
    rel = err / np.abs(g0).max()
    bad = np.argsort(-err)[:5]
    print(name, "max|g0|", np.abs(g0).max(), "rel err: median %.2e p99 %.2e max %.2e" % (np.median(rel), np.quantile(rel, 0.99), rel.max()),
          "rows>1e-2:", int((rel > 1e-2).sum()))
    for b in bad:
        print("   row", b, "g0", g0[b], "g1", g1[b], "sdf", res[0][1][b], res[1][1][b], "std", res[0][2][b], res[1][2][b], "valid", res[0][3][b])
