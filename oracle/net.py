"""Oracle (test infrastructure): fp32 torch-CPU restatement of the policy/value net.

The reference builds the graph with TensorFlow 1.x ``tf.layers`` (genData/
network.py:52-97,163-165).  TensorFlow is not vendored, not pinned (no requirements
file; API use implies 1.x <= 1.15) and not installable here, so this restatement
is *the* CPU oracle for the network.  TF conventions reproduced:

* ``channels_first`` NCHW, cross-correlation, SAME padding (2 for 5x5, 1 for 3x3),
  kernels stored HWIO -> ``w.permute(3, 2, 0, 1)`` for torch's OIHW;
* ``dense``: ``x @ K + b`` with ``K`` stored ``[in, out]``; the NCHW flatten index
  is ``c * S*S + i * S + j`` (network.py:71-72, 83-84);
* ELU alpha = 1; value head ``tanh(x / 2)`` (network.py:163-165); softmax over all
  S*S cells with no legality mask (network.py:88);
* default initialisers: glorot-uniform kernels, zero biases.

Variable names are the TF scope names of the shipped checkpoint
(``bone/conv1/kernel`` ...).

Parity: no TF-produced output vectors exist.  Pinned (i) by the logged losses at
ckpt-6960 (tests/test_oracle_net.py: x-entropy 2.107 / value-MSE 0.324 / entropy
2.152 on the shipped replay sample vs 2.155 / 0.313 / 2.145 logged) and (ii) end to
end by the reference's own recorded game tmp/five_6960.gif (GUI.py:184-186): with
ckpt-6960 this net, driving the search restatement, plays the AI's 29 moves exactly
(tests/test_oracle_gui_game.py; one of them decided by 198 against 193 visits).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# (name, kh, cin, cout) for every conv in graph order; residual blocks are
# {name}_res (1x1), {name}_conv1 (3x3 + ELU), {name}_conv2 (3x3)   network.py:52-56
BLOCKS = [("bone/block1", 32, 64), ("bone/block2", 64, 128),
          ("value/block3", 128, 32), ("policy/block4", 128, 64), ("policy/block5", 64, 32)]


def layer_shapes(size: int) -> dict[str, tuple]:
    """All 42 variables (name -> shape) in TF layout for an ``size x size`` board."""
    C = size * size
    sh = {"bone/conv1/kernel": (5, 5, 3, 32), "bone/conv1/bias": (32,)}
    for name, cin, cout in BLOCKS:
        sh[f"{name}_res/kernel"] = (1, 1, cin, cout)
        sh[f"{name}_conv1/kernel"] = (3, 3, cin, cout)
        sh[f"{name}_conv2/kernel"] = (3, 3, cout, cout)
        for s in ("res", "conv1", "conv2"):
            sh[f"{name}_{s}/bias"] = (cout,)
    sh.update({"value/conv/kernel": (1, 1, 32, 4), "value/conv/bias": (4,),
               "value/fc1/kernel": (4 * C, 64), "value/fc1/bias": (64,),
               "value/fc2/kernel": (64, 1), "value/fc2/bias": (1,),
               "policy/conv/kernel": (1, 1, 32, 16), "policy/conv/bias": (16,),
               "policy/fc/kernel": (16 * C, C), "policy/fc/bias": (C,)})
    return sh


def glorot_weights(size: int, seed: int = 0) -> dict[str, np.ndarray]:
    """What an un-restored reference net holds: glorot-uniform kernels (limit
    sqrt(6 / (fan_in + fan_out)), fans include the receptive field) and zero biases
    -- the ``tf.layers`` defaults (network.py:53-55,63,70,73,76,82,85)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in layer_shapes(size).items():
        if name.endswith("bias"):
            out[name] = np.zeros(shape, np.float32)
            continue
        rf = int(np.prod(shape[:-2])) if len(shape) == 4 else 1
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        out[name] = ((torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * lim).numpy()
    return out


def param_count(weights) -> int:
    return int(sum(v.size for v in weights.values()))


class OracleNet:
    """``eval(inputs) -> (prob f32[B, S*S], value f32[B])`` like network.py:90-97."""

    def __init__(self, size: int, weights: dict[str, np.ndarray], dtype=torch.float32, threads=None):
        self.size = size
        self.dtype = dtype
        self.threads = threads
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}
        self.k = {k: v.permute(3, 2, 0, 1).contiguous() for k, v in self.w.items()
                  if k.endswith("kernel") and v.dim() == 4}

    def _conv(self, x, name, act):
        k = self.k[name + "/kernel"]
        y = F.conv2d(x, k, self.w[name + "/bias"], padding=k.shape[-1] // 2)
        return F.elu(y) if act else y

    def _residual(self, x, name):
        skip = self._conv(x, name + "_res", False)
        y = self._conv(self._conv(x, name + "_conv1", True), name + "_conv2", False)
        return F.elu(skip + y)

    def _dense(self, x, name):
        return x @ self.w[name + "/kernel"] + self.w[name + "/bias"]

    @torch.no_grad()
    def forward(self, inputs):
        if self.threads:
            torch.set_num_threads(self.threads)
        x = torch.as_tensor(np.asarray(inputs)).to(self.dtype)
        f = self._conv(x, "bone/conv1", True)
        f = self._residual(f, "bone/block1")
        f = self._residual(f, "bone/block2")
        v = self._residual(f, "value/block3")
        v = self._conv(v, "value/conv", True).flatten(1)
        v = F.elu(self._dense(v, "value/fc1"))
        v = torch.tanh(self._dense(v, "value/fc2") / 2).squeeze(1)
        p = self._residual(f, "policy/block4")
        p = self._residual(p, "policy/block5")
        p = self._conv(p, "policy/conv", True).flatten(1)
        logits = self._dense(p, "policy/fc")
        return logits, v

    def eval(self, inputs):
        logits, v = self.forward(inputs)
        prob = torch.softmax(logits, dim=1)
        return prob.to(torch.float32).numpy(), v.to(torch.float32).numpy()

    def losses(self, inputs, target_policy, target_value):
        """(x-entropy, value MSE, policy entropy) as network.py:40-46,86-87 define them."""
        logits, v = self.forward(inputs)
        logp = torch.log_softmax(logits, dim=1)
        tp = torch.as_tensor(target_policy).to(self.dtype)
        tv = torch.as_tensor(target_value).to(self.dtype)
        xent = -(tp * logp).sum(1).mean()
        mse = ((v - tv) ** 2).mean()
        ent = -(logp.exp() * logp).sum(1).mean()
        return float(xent), float(mse), float(ent)
