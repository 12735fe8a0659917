"""oracle/net.py: the only first-party known answer for the network is the loss
curve logged by the reference at step ~6960."""
import numpy as np

from conftest import golden
from oracle import net, rules


def _ckpt():
    z = golden("ckpt6960.npz")
    return {k.replace("__", "/"): z[k] for k in z.files}


def test_param_count_matches_checkpoint():
    w = _ckpt()
    assert len(w) == 42 and net.param_count(w) == 754910
    sh = net.layer_shapes(11)
    assert {k: v.shape for k, v in w.items()} == sh
    assert net.param_count(net.glorot_weights(11)) == 754910
    assert net.param_count(net.glorot_weights(15)) == 1357382


def test_losses_reproduce_logged_values():
    g = golden("replay_sample.npz")
    x = np.stack([rules.input_planes(b, tuple(la) if la[0] >= 0 else None)
                  for b, la in zip(g["boards"], g["last_action"])])
    model = net.OracleNet(11, _ckpt())
    xent, mse, ent = model.losses(x, g["policy"].reshape(-1, 121), g["value"])
    logged = g["logged_losses"]
    assert abs(xent - logged[0]) < 0.12 and abs(mse - logged[1]) < 0.06 and abs(ent - logged[2]) < 0.12
    prob, value = model.eval(x[:8])
    assert prob.shape == (8, 121) and value.shape == (8,)
    np.testing.assert_allclose(prob.sum(1), 1.0, atol=1e-5)
    assert np.all(np.abs(value) < 1)


def test_fp64_agrees_with_fp32():
    import torch
    g = golden("replay_sample.npz")
    x = np.stack([rules.input_planes(b) for b in g["boards"][:64]])
    w = _ckpt()
    p32, v32 = net.OracleNet(11, w).eval(x)
    p64, v64 = net.OracleNet(11, w, dtype=torch.float64).eval(x)
    assert np.abs(p32 - p64).max() < 2e-5 and np.abs(v32 - v64).max() < 2e-5
