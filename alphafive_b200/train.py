"""The reference's training step on the device (network.py:40-50, main.py:36-39,57-68) in PyTorch.

SURVEY section 8(f) rank 3: the consumer of ``RandomStack.get_data`` that closes the
self-play -> replay -> train -> new weights loop without leaving the GPU.  The forward graph is
the one of network.py:58-88 (NCHW, SAME padding, ELU, TF variable names and layouts) run through
torch autograd (cuDNN, fp32 with TF32 off); the loss is

    total = -mean(w * sum(pi * log_softmax(logits))) + 2 * mean(w * (v - z)^2) + 4e-5 * sum_{non-bias} |theta|^2 / 2

(network.py:40-50; ``tf.nn.l2_loss`` halves the sum of squares) and the update is TensorFlow's
Adam (``tf.train.AdamOptimizer`` defaults beta1 0.9, beta2 0.999, eps 1e-8, with the
"epsilon-hat" form  theta -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)), the learning
rate following config.get_lr(step) (config.py:9,23-27).  Not reproduced: main.py:39 re-runs the
global initialiser after ``restore``, which discards the restored weights.

PyTorch is the engine here on purpose (the SURVEY names it): the hot path of this repository is
the search and the forward pass; this module only makes the loop complete.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

BLOCKS = [("bone/block1", 32, 64), ("bone/block2", 64, 128), ("value/block3", 128, 32),
          ("policy/block4", 128, 64), ("policy/block5", 64, 32)]


def forward_graph(w: dict, x: torch.Tensor):
    """logits [B, S*S], value [B] from TF-layout tensors ``w`` (network.py:58-88,163-165)."""
    def conv(t, name, act):
        k = w[name + "/kernel"].permute(3, 2, 0, 1)                       # HWIO -> OIHW
        y = F.conv2d(t, k, w[name + "/bias"], padding=k.shape[-1] // 2)
        return F.elu(y) if act else y

    def residual(t, name):
        return F.elu(conv(t, name + "_res", False) + conv(conv(t, name + "_conv1", True), name + "_conv2", False))

    f = conv(x, "bone/conv1", True)
    f = residual(residual(f, "bone/block1"), "bone/block2")
    v = conv(residual(f, "value/block3"), "value/conv", True).flatten(1)
    v = F.elu(v @ w["value/fc1/kernel"] + w["value/fc1/bias"])
    v = torch.tanh((v @ w["value/fc2/kernel"] + w["value/fc2/bias"]) / 2).squeeze(1)
    p = conv(residual(residual(f, "policy/block4"), "policy/block5"), "policy/conv", True).flatten(1)
    return p @ w["policy/fc/kernel"] + w["policy/fc/bias"], v


def losses(w: dict, boards, weights, values, policies, l2=4e-5):
    """(total, x-entropy, value MSE, policy entropy) of network.py:40-50,86-87."""
    logits, v = forward_graph(w, boards)
    logp = torch.log_softmax(logits, dim=1)
    xent = (policies * logp).sum(1)
    sq = (v - values) ** 2
    l2_loss = sum((t * t).sum() / 2 for k, t in w.items() if "bias" not in k)
    total = -(xent * weights).mean() + 2.0 * (sq * weights).mean() + l2 * l2_loss
    entropy = -(logp.exp() * logp).sum(1).mean()
    return total, -xent.mean(), sq.mean(), entropy


class Trainer:
    """``step(boards, weights, values, policies, lr)`` == one ``sess.run([..., opt])`` of main.py:65-68."""

    def __init__(self, board_size: int, weights: dict, device=None, dtype=torch.float32,
                 beta1=0.9, beta2=0.999, eps=1e-8, l2=4e-5):
        self.S = board_size
        self.device = torch.device(device if device is not None else "cuda")
        self.dtype = dtype
        self.w = {k: torch.tensor(np.asarray(v), dtype=dtype, device=self.device, requires_grad=True)
                  for k, v in weights.items()}
        self.m = {k: torch.zeros_like(t) for k, t in self.w.items()}
        self.v = {k: torch.zeros_like(t) for k, t in self.w.items()}
        self.t = 0
        self.beta1, self.beta2, self.eps, self.l2 = beta1, beta2, eps, l2

    def _t(self, a):
        return torch.as_tensor(a).to(self.device, self.dtype)

    def step(self, boards, weights, values, policies, lr: float):
        """One Adam step on a batch (numpy arrays or tensors, as ``get_data`` / ``get_data_device`` return
        them).  Returns (x-entropy, value MSE, entropy) -- what main.py:73 prints."""
        tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            total, xent, mse, ent = losses(self.w, self._t(boards), self._t(weights), self._t(values),
                                           self._t(policies).reshape(len(boards), -1), self.l2)
            grads = torch.autograd.grad(total, list(self.w.values()))
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        with torch.no_grad():
            for (k, p), g in zip(self.w.items(), grads):
                self.m[k].mul_(self.beta1).add_(g, alpha=1.0 - self.beta1)
                self.v[k].mul_(self.beta2).addcmul_(g, g, value=1.0 - self.beta2)
                p.addcdiv_(self.m[k], self.v[k].sqrt().add_(self.eps), value=-lr_t)
        return float(xent.detach()), float(mse.detach()), float(ent.detach())

    def weights(self) -> dict:
        """TF-layout fp32 numpy tensors (what ``DeviceNet.set_weights`` / ``ResNet.set_weights`` take)."""
        return {k: t.detach().to(torch.float32).cpu().numpy() for k, t in self.w.items()}

    def sync_to(self, net):
        """Load the current weights into a DeviceNet / ResNet facade (the search sees them from the next pass)."""
        net.set_weights(self.weights())
