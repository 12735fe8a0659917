"""Oracle training step (oracle/train.py): loss terms against the values the reference logged at
ckpt-6960, TensorFlow-Adam's first step, and descent on a fixed batch."""
import numpy as np

from conftest import golden
from oracle import net as onet
from oracle.train import OracleTrainer


def _ckpt():
    z = golden("ckpt6960.npz")
    return {k.replace("__", "/"): z[k] for k in z.files}


def _batch(n=96):
    z = golden("replay_stack.npz")
    return z["batch_boards"][:n], z["batch_weights"][:n], z["batch_values"][:n], z["batch_policies"][:n]


def test_loss_terms_match_the_net_oracle_and_logged_range():
    b, w, v, p = _batch(256)
    tr = OracleTrainer(11, _ckpt())
    total, xent, mse, ent = (float(t.detach()) for t in tr.loss_terms(b, w, v, p))
    want = onet.OracleNet(11, _ckpt()).losses(b, p, v)
    assert abs(xent - want[0]) < 1e-4 and abs(mse - want[1]) < 1e-4 and abs(ent - want[2]) < 1e-4
    # main.py logged x-entropy 2.155 / mse 0.313 / entropy 2.145 around this checkpoint (whole-buffer averages)
    assert 1.7 < xent < 2.6 and 0.1 < mse < 0.5 and 1.7 < ent < 2.6
    # total = weighted terms + 4e-5 * sum |theta|^2 / 2 over the non-bias tensors (network.py:47-50)
    net = onet.OracleNet(11, _ckpt())
    logits, val = net.forward(b)
    logp = np.log(np.exp(logits.numpy() - logits.numpy().max(1, keepdims=True)) /
                  np.exp(logits.numpy() - logits.numpy().max(1, keepdims=True)).sum(1, keepdims=True))
    l2 = sum(float((np.asarray(t, np.float64) ** 2).sum()) / 2 for k, t in _ckpt().items() if "bias" not in k)
    want_total = -np.mean((p * logp).sum(1) * w) + 2.0 * np.mean((val.numpy() - v) ** 2 * w) + 4e-5 * l2
    assert abs(total - want_total) < 1e-4 * abs(want_total)


def test_first_adam_step_is_lr_times_sign_of_gradient():
    """TF Adam, t = 1: m = (1-b1) g, v = (1-b2) g^2, lr_t = lr sqrt(1-b2)/(1-b1)  =>  d theta = -lr g / (|g| + eps')."""
    b, w, v, p = _batch(48)
    tr = OracleTrainer(11, onet.glorot_weights(11, 0))
    before = tr.weights()
    import torch
    total = tr.loss_terms(b, w, v, p)[0]
    names = list(tr.net.w)
    grads = dict(zip(names, torch.autograd.grad(total, [tr.net.w[k] for k in names])))
    tr.step(b, w, v, p, lr=1e-3)
    after = tr.weights()
    for k in ("bone/conv1/kernel", "policy/fc/kernel", "value/fc2/bias"):
        g = grads[k].numpy()
        d = after[k] - before[k]
        big = np.abs(g) > 1e-3
        assert big.any()
        assert np.allclose(d[big], -1e-3 * np.sign(g[big]), rtol=1e-3, atol=0)
        assert np.allclose(d, -1e-3 * g / (np.abs(g) + 1e-8 / np.sqrt(1 - 0.999)), rtol=1e-7, atol=1e-15)
        assert (np.abs(d) <= 1e-3 * (1 + 1e-9)).all()


def test_steps_descend_on_a_fixed_batch():
    b, w, v, p = _batch(64)
    tr = OracleTrainer(11, onet.glorot_weights(11, 0))
    first = float(tr.loss_terms(b, w, v, p)[0].detach())
    for _ in range(6):
        tr.step(b, w, v, p, lr=1e-3)
    assert float(tr.loss_terms(b, w, v, p)[0].detach()) < first - 0.05
