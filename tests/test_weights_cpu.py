"""Host logic (no GPU): the packed weight blobs, read back with the kernels' indexing, reproduce the oracle networks."""
import numpy as np
import torch

from oracle import nets


def _dense_from_blob(blob, off, n_in, n_out, x):
    """numpy twin of dense() in csrc/decoder_simt.cuh: chunks of 64 outputs, each [n_in][cw]."""
    out = np.zeros((x.shape[0], n_out), dtype=np.float64)
    for c0 in range(0, n_out, 64):
        cw = min(64, n_out - c0)
        m = blob[off + c0 * n_in: off + c0 * n_in + n_in * cw].reshape(n_in, cw).astype(np.float64)
        out[:, c0:c0 + cw] = x @ m
    return out


def test_decoder_blob_matches_oracle(dfb, weights):
    blob = dfb.weights.pack_decoder(weights)
    F0, F1, F2, F3 = 0, 4096, 4096 + 16384, 4096 + 16384 + 12288
    B3 = F3 + 16384; B2 = B3 + 12288; B1 = B2 + 12288; SM = B1 + 16384
    small = blob[SM:]
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.randn(64, 29) * 0.1, rng.rand(64, 3) - 0.5], 1).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    sdf, std = nets.decoder_forward(weights, xt)
    g = torch.autograd.grad((sdf / std.detach()).sum(), xt)[0][:, 29:].numpy()

    xd = x.astype(np.float64)
    a0 = _dense_from_blob(blob, F0, 32, 128, xd) + small[0:128]; h0 = np.maximum(a0, 0)
    a1 = _dense_from_blob(blob, F1, 128, 128, h0) + small[128:256]; h1 = np.maximum(a1, 0)
    a2 = _dense_from_blob(blob, F2, 128, 96, h1) + small[256:352]; h2 = np.maximum(a2, 0)
    a3 = _dense_from_blob(blob, F3, 128, 128, np.concatenate([h2, xd], 1)) + small[352:480]; h3 = np.maximum(a3, 0)
    z = h3 @ small[480:608] + small[1504]; u = h3 @ small[608:736] + small[1505]
    s = np.tanh(z); sd = 0.05 + 0.5 * np.log1p(np.exp(u))
    np.testing.assert_allclose(s, sdf.detach().numpy(), atol=2e-6)
    np.testing.assert_allclose(sd, std.detach().numpy(), atol=2e-6)
    seed = (1 - s * s) / sd
    d3 = seed[:, None] * small[480:608][None, :] * (a3 > 0)
    gx = d3 @ small[736:1120].reshape(128, 3)
    d2 = _dense_from_blob(blob, B3, 128, 96, d3) * (a2 > 0)
    d1 = _dense_from_blob(blob, B2, 96, 128, d2) * (a1 > 0)
    d0 = _dense_from_blob(blob, B1, 128, 128, d1) * (a0 > 0)
    gx = gx + d0 @ small[1120:1504].reshape(128, 3)
    np.testing.assert_allclose(gx, g, rtol=1e-4, atol=1e-5)


def test_encoder_blob_matches_oracle(dfb, weights):
    blob = dfb.weights.pack_encoder(weights)[:27264].astype(np.float64)
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.rand(50, 3) - 0.5, rng.randn(50, 3)], 1).astype(np.float32)
    ref = nets.encoder_forward(weights, torch.from_numpy(x)).numpy()
    o = 0
    w0 = blob[o:o + 256].reshape(32, 8)[:, :6]; o += 256
    b0 = blob[o:o + 32]; o += 32
    w1 = blob[o:o + 2048].reshape(64, 32); o += 2048
    b1 = blob[o:o + 64]; o += 64
    w2 = blob[o:o + 16384].reshape(32, 64, 8); o += 16384          # [group][k][8]
    b2 = blob[o:o + 256]; o += 256
    w3 = blob[o:o + 8192].reshape(256, 32); o += 8192
    b3 = blob[o:o + 32]
    h0 = np.maximum(x @ w0.T + b0, 0)
    h1 = np.maximum(h0 @ w1.T + b1, 0)
    h2 = np.maximum(np.einsum("nk,gkj->ngj", h1, w2).reshape(-1, 256) + b2, 0)
    out = (h2 @ w3 + b3)[:, :29]
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=2e-5)


def test_tc_images_follow_the_kernel_swizzle(dfb, weights):
    """FP16 SWIZZLE_128B images of the tcgen05 engine: element (row, k) of a layer lives at
    block (k // 64), row, 16-byte chunk ((k % 64) // 8) ^ (row & 7), lane k % 8 (csrc/decoder_tc.cu)."""
    blob = dfb.weights.pack_decoder(weights)
    assert blob.size == 91624 + 50688
    tc = blob[91624:].view(np.uint8)
    M = dfb.weights.decoder_matrices(weights)

    def read(img_off, rows, K):
        h = tc[img_off: img_off + rows * K * 2].view(np.float16).reshape(K // 64, rows, 8, 8)
        out = np.zeros((rows, K), np.float16)
        for r in range(rows):
            for c in range(K // 8):
                out[r, 8 * c: 8 * c + 8] = h[c // 8, r, (c % 8) ^ (r & 7)]
        return out
    def hi(m):
        return m.astype(np.float16)

    def lo(m):
        return (m - m.astype(np.float16).astype(np.float32)).astype(np.float16)
    w0 = read(0, 128, 64)                                    # hi | lo halves of W0 (csrc/decoder_tc.cu IMG_*)
    assert np.array_equal(w0[:, :32], hi(M["W0"])) and np.array_equal(w0[:, 32:], lo(M["W0"]))
    assert np.array_equal(read(16384, 128, 128), hi(M["W1"])) and np.array_equal(read(49152, 128, 128), lo(M["W1"]))
    assert np.array_equal(read(81920, 96, 128), hi(M["W2"])) and np.array_equal(read(106496, 96, 128), lo(M["W2"]))
    assert np.array_equal(read(131072, 128, 128), hi(M["W3"])) and np.array_equal(read(163840, 128, 128), lo(M["W3"]))
    # hi + lo carries ~22 significant bits of every weight
    for k, off, rows in (("W1", (16384, 49152), 128), ("W3", (131072, 163840), 128)):
        rec = read(off[0], rows, 128).astype(np.float32) + read(off[1], rows, 128).astype(np.float32)
        assert np.abs(rec - M[k]).max() <= 2.0 ** -21 * np.abs(M[k]).max()
    small = tc[196608:].view(np.float32)
    assert np.array_equal(small[:1512], blob[90112:90112 + 1512])
