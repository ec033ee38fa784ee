"""Host-side weight folding and packing into the blobs the kernels read (layouts: csrc/decoder_simt.cuh DB_*,
csrc/integrate.cu EB_*).

Replaces network/utility.py:22-58 (load_model), the weight-norm recompute that runs on every decoder forward after
fix_weight_norm_pickle (network/utility.py:211-220; W = g * v / ||v||_row) and eval-mode BatchNorm in the encoder
(utils/pt_util.py:193-206), which is folded into the preceding 1x1 convolution.
Checkpoint tensors keep the reference's key names with a 'dec.' / 'enc.' prefix.
"""
from pathlib import Path

import numpy as np
import torch


def load_checkpoint(hyper_dir, epoch=300):
    """Read ckpt/default/{model,encoder}_<epoch>.pth.tar (reference format: {'epoch', 'model_state'})."""
    hyper_dir = Path(hyper_dir)
    out = {}
    sd = torch.load(hyper_dir / f"model_{epoch}.pth.tar", map_location="cpu", weights_only=False)["model_state"]
    for k, v in sd.items():
        out["dec." + k] = v.float()
    se = torch.load(hyper_dir / f"encoder_{epoch}.pth.tar", map_location="cpu", weights_only=False)["model_state"]
    for k, v in se.items():
        out["enc." + k] = v.float() if v.is_floating_point() else v
    return out


def load_npz(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def decoder_matrices(W):
    """Effective (weight-normed) decoder matrices, fp32, computed like torch._weight_norm: v * (g / ||v||)."""
    M = {}
    for l in range(5):
        v, g = W[f"dec.lin{l}.weight_v"].float(), W[f"dec.lin{l}.weight_g"].float()
        M[f"W{l}"] = (v * (g / v.norm(dim=1, keepdim=True))).numpy()
        M[f"b{l}"] = W[f"dec.lin{l}.bias"].float().numpy()
    M["Wu"] = W["dec.uncertainty_layer.weight"].float().numpy()
    M["bu"] = W["dec.uncertainty_layer.bias"].float().numpy()
    return M


def _pack_dense(mat):
    """mat (n_in, n_out) -> chunks of 64 output columns, each stored [n_in][cw] row-major."""
    n_in, n_out = mat.shape
    parts = []
    for c0 in range(0, n_out, 64):
        parts.append(np.ascontiguousarray(mat[:, c0:min(c0 + 64, n_out)]).reshape(-1))
    return np.concatenate(parts)


def pack_decoder(W):
    M = decoder_matrices(W)
    W0, W1, W2, W3, W4 = M["W0"], M["W1"], M["W2"], M["W3"], M["W4"]
    assert W0.shape == (128, 32) and W1.shape == (128, 128) and W2.shape == (96, 128) and W3.shape == (128, 128) and W4.shape == (1, 128)
    parts = [
        _pack_dense(W0.T), _pack_dense(W1.T), _pack_dense(W2.T), _pack_dense(W3.T),       # forward: [k][o]
        _pack_dense(W3[:, :96]), _pack_dense(W2), _pack_dense(W1),                          # reverse: [o][k]
    ]
    small = np.zeros(1512, dtype=np.float32)
    small[0:128] = M["b0"]; small[128:256] = M["b1"]; small[256:352] = M["b2"]; small[352:480] = M["b3"]
    small[480:608] = W4[0]; small[608:736] = M["Wu"][0]
    small[736:736 + 384] = W3[:, 125:128].reshape(-1)
    small[1120:1120 + 384] = W0[:, 29:32].reshape(-1)
    small[1504] = M["b4"][0]; small[1505] = M["bu"][0]
    blob = np.concatenate(parts + [small]).astype(np.float32)
    tc = pack_decoder_tc(M, small)
    return np.concatenate([blob, tc.view(np.float32)])


def _swizzled_image(mat):
    """mat (rows, K) float -> FP16 K-major SWIZZLE_128B image: blocks of 64 columns, each [rows][128 B]; the 16-byte
    chunk c of row r is stored at chunk position c ^ (r & 7) (csrc/decoder_tc.cu store_cols32 uses the same map)."""
    rows, K = mat.shape
    assert K % 64 == 0 and rows % 8 == 0
    h = mat.astype(np.float16)
    out = np.zeros((K // 64, rows, 8, 8), dtype=np.float16)
    r = np.arange(rows)
    for b in range(K // 64):
        blk = h[:, 64 * b:64 * b + 64].reshape(rows, 8, 8)
        for c in range(8):
            out[b, r, c ^ (r & 7)] = blk[:, c]
    return out.reshape(-1)


def pack_decoder_tc(M, small):
    """FP16 images for the tcgen05 engine + the FP32 small block, as bytes (uint8); layout csrc/decoder_tc.cu IMG_*.
    Every matrix is split hi = fp16(W), lo = fp16(W - hi) (the engine computes A_hi W_hi + A_lo W_hi + A_hi W_lo);
    one image per layer and half serves the forward pass (read K-major) and the reverse pass (read MN-major)."""
    W0, W1, W2, W3 = M["W0"], M["W1"], M["W2"], M["W3"]

    def hi(m):
        return m.astype(np.float16).astype(np.float32)
    w0 = np.concatenate([hi(W0), W0 - hi(W0)], axis=1)                      # (128, 64): hi | lo
    imgs = [_swizzled_image(w0), _swizzled_image(hi(W1)), _swizzled_image(W1 - hi(W1)), _swizzled_image(hi(W2)),
            _swizzled_image(W2 - hi(W2)), _swizzled_image(hi(W3)), _swizzled_image(W3 - hi(W3))]
    img = np.concatenate(imgs).view(np.uint8)
    assert img.size == 196608, img.size
    sm = np.zeros(6144 // 4, np.float32); sm[:small.size] = small
    return np.concatenate([img, sm.view(np.uint8)])


def encoder_matrices(W):
    """BN-folded encoder layers: list of (weight (out,in), bias (out,))."""
    layers = []
    for l in range(3):
        w = W[f"enc.mlp.layer{l}.conv.weight"].float().squeeze(-1).double()
        g = W[f"enc.mlp.layer{l}.normlayer.bn.weight"].double(); b = W[f"enc.mlp.layer{l}.normlayer.bn.bias"].double()
        mu = W[f"enc.mlp.layer{l}.normlayer.bn.running_mean"].double(); var = W[f"enc.mlp.layer{l}.normlayer.bn.running_var"].double()
        s = g / torch.sqrt(var + 1e-5)
        layers.append(((w * s[:, None]).float().numpy(), (b - s * mu).float().numpy()))
    layers.append((W["enc.mlp.layer3.conv.weight"].float().squeeze(-1).numpy(), W["enc.mlp.layer3.conv.bias"].float().numpy()))
    return layers


def pack_encoder(W):
    (w0, b0), (w1, b1), (w2, b2), (w3, b3) = encoder_matrices(W)
    assert w0.shape == (32, 6) and w1.shape == (64, 32) and w2.shape == (256, 64) and w3.shape == (29, 256)
    p0 = np.zeros((32, 8), np.float32); p0[:, :6] = w0
    p2 = np.ascontiguousarray(w2.reshape(32, 8, 64).transpose(0, 2, 1))            # [group][k][8]
    p3 = np.zeros((256, 32), np.float32); p3[:, :29] = w3.T
    pb3 = np.zeros(32, np.float32); pb3[:29] = b3
    blob = np.concatenate([p0.reshape(-1), b0, w1.reshape(-1), b1, p2.reshape(-1), b2, p3.reshape(-1), pb3]).astype(np.float32)
    return np.concatenate([blob, pack_encoder_tc(w0, b0, w1, b1, w2, b2, w3, b3).view(np.float32)])


def pack_encoder_tc(w0, b0, w1, b1, w2, b2, w3, b3):
    """FP16 SWIZZLE_128B images + FP32 biases for the tcgen05 encoder engine (csrc/encoder_tc.cu IMG_* / ES_*).
    Every matrix is split hi = fp16(W), lo = fp16(W - hi); the engine computes A_hi W_hi + A_lo W_hi + A_hi W_lo.
    Layers 0 and 1 keep hi and lo side by side in one 64-column block (lo at K step 1 / 2), layers 2 and 3 as two images."""
    def h(m):
        return m.astype(np.float16).astype(np.float32)
    w3p = np.zeros((32, 256), np.float32); w3p[:29] = w3
    i0 = np.zeros((32, 64), np.float32); i0[:, 0:6] = h(w0); i0[:, 16:22] = w0 - h(w0)
    i1 = np.concatenate([h(w1), w1 - h(w1)], axis=1)                                 # (64, 64)
    mats = [i0, i1, h(w2), w2 - h(w2), h(w3p), w3p - h(w3p)]
    img = np.concatenate([_swizzled_image(m) for m in mats]).view(np.uint8)
    assert img.size == 110592, img.size
    sm = np.zeros(2048 // 4, np.float32)
    sm[0:32] = b0; sm[32:96] = b1; sm[96:352] = b2; sm[352:381] = b3
    return np.concatenate([img, sm.view(np.uint8)])
