"""Oracle for the convolutional autoencoder forward pass and reconstruction errors.

Serves improved_detection.py:125-131 with the architecture of
CAE_improved_modeltrain.py:188-216.  TensorFlow/Keras are not
installable here: the layer math is restated in fp32 with torch-CPU convolutions
-- PARITY UNPINNED against Keras.  Test infrastructure only.

Weights container (``CAEWeights`` below, the same plain-NumPy structure the product
loader produces): conv kernels HWIO float32 (3,3,Cin,Cout), biases, and for the six
BatchNormalization layers gamma/beta/moving_mean/moving_variance, epsilon 1e-3.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
# (Cin, Cout) of the seven Conv2D layers, train:191-216
CONV_SHAPES = [(1, 32), (32, 64), (64, 32), (32, 32), (32, 64), (64, 32), (32, 1)]


def bn_affine(gamma, beta, mean, var, eps=BN_EPS):
    """tf.nn.batch_normalization's inference form: y = x*inv + (beta - mean*inv)."""
    inv = (gamma.astype(np.float32) / np.sqrt(var.astype(np.float32) + np.float32(eps))).astype(np.float32)
    return inv, (beta.astype(np.float32) - mean.astype(np.float32) * inv).astype(np.float32)


def _conv(x, k, b, exact=True):
    """Conv2D 'same' cross-correlation + bias, float32.

    ``exact=True`` (parity tests): every output is the CORRECTLY ROUNDED float32 of the
    exact sum of products (float32 x float32 products are exact in float64; the float64
    accumulation error is ~1e-16), then the bias is added in float32 like TF's separate
    BiasAdd.  This is the order-independent centre of every float32 summation order a
    real TensorFlow/Eigen/oneDNN build may use (each deviates from it by its own
    ~sqrt(K)*2^-24 noise) -- which matters because the one-class SVM decision moves by
    ~7e-5 per 1e-6 relative feature noise (measured), i.e. the 1e-4 gate sits at the
    float32 noise floor of the reference itself.
    ``exact=False`` (CPU-baseline timing): plain torch/oneDNN float32 convolution."""
    w = torch.from_numpy(np.ascontiguousarray(k.transpose(3, 2, 0, 1)))  # HWIO -> OIHW
    if not exact:
        return F.conv2d(x, w, torch.from_numpy(b), padding=1)
    y = F.conv2d(x.double(), w.double(), None, padding=1).float()
    return y + torch.from_numpy(b)[None, :, None, None]


def _bn(x, bn):
    s, t = bn_affine(*bn)
    return x * torch.from_numpy(s)[None, :, None, None] + torch.from_numpy(t)[None, :, None, None]


@torch.no_grad()
def forward(X: np.ndarray, w, batch: int = 256, n_layers: int = 7, exact: bool = True):
    """X float32 [N,64,64,1] -> (recon float32 [N,64,64,1], encoded float32 [N,8,8,32]).

    ``w`` has attributes/keys ``kernels`` (7), ``biases`` (7), ``bns`` (6 x (gamma,
    beta, mean, var)).  ``n_layers=3`` runs the encoder only (encoder.keras, det:130).
    """
    recs, encs = [], []
    for s in range(0, len(X), batch):
        x = torch.from_numpy(np.ascontiguousarray(X[s:s + batch, :, :, 0]))[:, None]
        enc = None
        for i in range(n_layers):
            x = _conv(x, w["kernels"][i], w["biases"][i], exact)
            if i < 6:
                x = _bn(torch.relu(x), w["bns"][i])
                if i < 3:
                    x = F.max_pool2d(x, 2)
                    if i == 2:
                        enc = x
                else:
                    x = F.interpolate(x, scale_factor=2, mode="nearest")
            else:
                x = torch.sigmoid(x)
        encs.append(enc.permute(0, 2, 3, 1).contiguous().numpy())
        if n_layers == 7:
            recs.append(x.permute(0, 2, 3, 1).contiguous().numpy())
    rec = np.concatenate(recs) if recs else None
    return rec, np.concatenate(encs)


def recon_errors(X: np.ndarray, recon: np.ndarray):
    """det:126-127, float32 as NumPy computes them."""
    return (np.mean(np.square(X - recon), axis=(1, 2, 3)),
            np.mean(np.abs(X - recon), axis=(1, 2, 3)))
