"""Oracle (test infrastructure): float64 CPU restatement of the reference's training step --
loss of network.py:40-50 on the graph of ``oracle.net`` and TensorFlow's Adam
(``tf.train.AdamOptimizer(lr).minimize(total_loss)``, main.py:38-39) -- to check
``alphafive_b200.train.Trainer`` against.

Parity: the loss terms are pinned by the losses the reference logged at ckpt-6960
(tests/test_oracle_net.py); the optimiser has no first-party vectors (TensorFlow cannot run
here): its update rule is restated from TF 1.x's documented algorithm
(``m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2; theta <- theta - lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v)+eps)``)
and checked against a hand-computed scalar case -- parity of the optimiser is unpinned.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import net as onet


class OracleTrainer:
    def __init__(self, size, weights, beta1=0.9, beta2=0.999, eps=1e-8, l2=4e-5):
        self.net = onet.OracleNet(size, weights, dtype=torch.float64)
        for d in (self.net.w,):
            for k in d:
                d[k].requires_grad_(True)
        self.m = {k: torch.zeros_like(t) for k, t in self.net.w.items()}
        self.v = {k: torch.zeros_like(t) for k, t in self.net.w.items()}
        self.t, self.beta1, self.beta2, self.eps, self.l2 = 0, beta1, beta2, eps, l2

    def _refresh(self):
        self.net.k = {k: v.permute(3, 2, 0, 1) for k, v in self.net.w.items() if k.endswith("kernel") and v.dim() == 4}

    def loss_terms(self, boards, weights, values, policies):
        self._refresh()
        with torch.enable_grad():
            x = torch.as_tensor(np.asarray(boards)).to(torch.float64)
            f = self.net
            # OracleNet.forward runs under no_grad: call its undecorated body
            logits, v = onet.OracleNet.forward.__wrapped__(f, x)
            logp = torch.log_softmax(logits, dim=1)
            tp = torch.as_tensor(np.asarray(policies)).to(torch.float64).reshape(len(x), -1)
            tw = torch.as_tensor(np.asarray(weights)).to(torch.float64)
            tv = torch.as_tensor(np.asarray(values)).to(torch.float64)
            xent = (tp * logp).sum(1)                                           # network.py:41
            sq = (v - tv) ** 2                                                  # network.py:44
            l2_loss = sum((t * t).sum() / 2 for k, t in f.w.items() if "bias" not in k)   # network.py:47-48
            total = -(xent * tw).mean() + 2.0 * (sq * tw).mean() + self.l2 * l2_loss    # network.py:50
            ent = -(logp.exp() * logp).sum(1).mean()
        return total, -xent.mean(), sq.mean(), ent

    def step(self, boards, weights, values, policies, lr):
        total, xent, mse, ent = self.loss_terms(boards, weights, values, policies)
        names = list(self.net.w)
        grads = torch.autograd.grad(total, [self.net.w[k] for k in names])
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        with torch.no_grad():
            for k, g in zip(names, grads):
                self.m[k] = self.beta1 * self.m[k] + (1 - self.beta1) * g
                self.v[k] = self.beta2 * self.v[k] + (1 - self.beta2) * g * g
                self.net.w[k] -= lr_t * self.m[k] / (self.v[k].sqrt() + self.eps)
        return float(xent.detach()), float(mse.detach()), float(ent.detach())

    def weights(self):
        return {k: t.detach().numpy().copy() for k, t in self.net.w.items()}
