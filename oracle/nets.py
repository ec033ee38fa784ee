"""Oracle restatement of the two networks.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/network/di_decoder.py:55-86 (decoder), network/di_encoder.py:26-30 +
utils/pt_util.py:76-127,193-206 (encoder), weights in the layout of ckpt/default/*.pth.tar
(exported verbatim to tests/golden/weights.npz by oracle/make_golden.py).
All math is torch-CPU fp32 so autograd gives the tracker's d(sdf/std)/dxyz (tracker.py:191-197).
"""
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_weights(path=None):
    """Returns {name: torch.float32 tensor} with the checkpoint's own key names
    (decoder keys prefixed 'dec.', encoder keys 'enc.')."""
    path = Path(path) if path is not None else GOLDEN / "weights.npz"
    z = np.load(path)
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def _wn(v, g):
    # nn.utils.weight_norm, dim=0: w = v * (g / ||v||_row)   (di_decoder.py:38-41)
    return v * (g / v.norm(dim=1, keepdim=True))


def decoder_effective_weights(W):
    """Weight-normed matrices W0..W4 (row = output unit) + biases, uncertainty head."""
    out = {}
    for l in range(5):
        out[f"W{l}"] = _wn(W[f"dec.lin{l}.weight_v"], W[f"dec.lin{l}.weight_g"])
        out[f"b{l}"] = W[f"dec.lin{l}.bias"]
    out["Wu"] = W["dec.uncertainty_layer.weight"]
    out["bu"] = W["dec.uncertainty_layer.bias"]
    return out


def decoder_forward(W, x):
    """di_decoder.py:55-86 with dims [32,128,128,128,128,1], latent_in=[3], weight_norm, eval mode.
    x: (N, 32) = [latent(29), xyz(3)].  Returns sdf (N,), std (N,)."""
    E = decoder_effective_weights(W)
    h = x
    std = None
    for layer in range(5):
        if layer == 3:
            h = torch.cat([h, x], 1)
        if layer == 4:
            std = 0.05 + 0.5 * F.softplus(F.linear(h, E["Wu"], E["bu"]))
        h = F.linear(h, E[f"W{layer}"], E[f"b{layer}"])
        if layer < 4:
            h = torch.relu(h)
    return torch.tanh(h).squeeze(-1), std.squeeze(-1)


def encoder_forward(W, x):
    """di_encoder.py:26-30 ('cnp' mode): SharedMLP 6->32->64->256->29, Conv1d(k=1, no bias)+BN(eval)+ReLU
    for the first three layers (pt_util.py:83: bias dropped when bn is set), plain Conv1d+bias last.
    x: (M, 6) = [rel xyz, normal].  Returns (M, 29)."""
    h = x
    for l in range(3):
        w = W[f"enc.mlp.layer{l}.conv.weight"].squeeze(-1)
        h = F.linear(h, w)
        h = F.batch_norm(h, W[f"enc.mlp.layer{l}.normlayer.bn.running_mean"],
                         W[f"enc.mlp.layer{l}.normlayer.bn.running_var"],
                         W[f"enc.mlp.layer{l}.normlayer.bn.weight"],
                         W[f"enc.mlp.layer{l}.normlayer.bn.bias"], False, 0.1, 1e-5)
        h = torch.relu(h)
    w = W["enc.mlp.layer3.conv.weight"].squeeze(-1)
    return F.linear(h, w, W["enc.mlp.layer3.conv.bias"])
