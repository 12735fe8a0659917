"""Oracle (test infrastructure): a second, torch-free restatement of the policy/value net in float64 numpy,
written from the TensorFlow definitions the reference calls (genData/network.py:52-88,163-165), as a
cross-check of oracle/net.py -- the two share no convolution / dense / softmax code, so a layout or ordering
mistake would have to be made twice in different words to go unnoticed.

tf.layers.conv2d(data_format="channels_first", padding="SAME", strides 1) is a cross-correlation:
    out[b, o, i, j] = bias[o] + sum_{ky, kx, c} x[b, c, i + ky - kh//2, j + kx - kw//2] * K[ky, kx, c, o]
with zeros outside the board (odd kernels: SAME pads kh//2 on both sides); tf.layers.dense is x @ K + b with
K [in, out]; tf.reshape of an NCHW tensor flattens as c * S*S + i * S + j; ELU alpha = 1;
half_tanh(x) = tanh(x / 2) (network.py:163-165); prob = softmax over all S*S logits (network.py:88).

Parity: no TensorFlow output vectors exist (TF 1.x is absent); oracle/net.py, which this file cross-checks, is
pinned by the logged losses and by the reference's recorded game (see its header).
"""
from __future__ import annotations

import numpy as np


def _conv_same(x, K, b):
    """x f64[B, C, S, S], K f64[kh, kw, C, O] (TF HWIO), b f64[O] -> f64[B, O, S, S]."""
    B, C, S, _ = x.shape
    kh, kw, _, O = K.shape
    ph, pw = kh // 2, kw // 2
    xp = np.zeros((B, C, S + 2 * ph, S + 2 * pw))
    xp[:, :, ph:ph + S, pw:pw + S] = x
    out = np.zeros((B, O, S, S))
    for ky in range(kh):
        for kx in range(kw):
            out += np.einsum("bcij,co->boij", xp[:, :, ky:ky + S, kx:kx + S], K[ky, kx])
    return out + b[None, :, None, None]


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def forward(weights: dict, inputs) -> tuple[np.ndarray, np.ndarray]:
    """(prob f64[B, S*S], value f64[B]) for inputs [B, 3, S, S]."""
    w = {k: np.asarray(v, np.float64) for k, v in weights.items()}
    x = np.asarray(inputs, np.float64)

    def conv(t, name, act):
        y = _conv_same(t, w[name + "/kernel"], w[name + "/bias"])
        return _elu(y) if act else y

    def residual(t, name):                                   # network.py:52-56
        res = conv(t, name + "_res", False)
        f = conv(conv(t, name + "_conv1", True), name + "_conv2", False)
        return _elu(res + f)

    f = conv(x, "bone/conv1", True)                          # network.py:63
    f = residual(f, "bone/block1")
    f = residual(f, "bone/block2")
    v = residual(f, "value/block3")                          # network.py:68-76
    v = conv(v, "value/conv", True).reshape(x.shape[0], -1)
    v = _elu(v @ w["value/fc1/kernel"] + w["value/fc1/bias"])
    v = np.tanh((v @ w["value/fc2/kernel"] + w["value/fc2/bias"]) / 2.0)[:, 0]
    p = residual(f, "policy/block4")                         # network.py:79-85
    p = residual(p, "policy/block5")
    p = conv(p, "policy/conv", True).reshape(x.shape[0], -1)
    logits = p @ w["policy/fc/kernel"] + w["policy/fc/bias"]
    e = np.exp(logits - logits.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True), v
