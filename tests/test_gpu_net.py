"""Device policy/value net vs the fp32 torch-CPU oracle: |dp|, |dv| <= 1e-4 (north star)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import net as onet, rules as orules

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _ckpt():
    z = golden("ckpt6960.npz")
    return {k.replace("__", "/"): z[k] for k in z.files}


def _planes(boards, last):
    return np.stack([orules.input_planes(b, tuple(la) if la[0] >= 0 else None) for b, la in zip(boards, last)])


def _modes():
    from alphafive_b200 import _lib
    return [_lib.NET_FP32, _lib.NET_TC]


@pytest.mark.parametrize("mode", [0, 1])
def test_trained_weights_parity(cuda_lib, mode):
    from alphafive_b200.net import DeviceNet
    g = golden("replay_sample.npz")
    x = _planes(g["boards"], g["last_action"])
    w = _ckpt()
    want_p, want_v = onet.OracleNet(11, w).eval(x)
    net = DeviceNet(11, 1024, w, mode=mode)
    got_p, got_v = net.eval(x)
    assert np.abs(got_p - want_p).max() <= TOL, np.abs(got_p - want_p).max()
    assert np.abs(got_v - want_v).max() <= TOL, np.abs(got_v - want_v).max()
    assert (got_p.argmax(1) == want_p.argmax(1)).mean() > 0.995


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("S", [11, 15])
def test_random_init_parity(cuda_lib, S, mode):
    from alphafive_b200.net import DeviceNet, glorot_init
    rng = np.random.default_rng(S)
    boards = np.stack([orules.random_board(rng, S) for _ in range(300)])
    last = rng.integers(-1, S, size=(300, 2))
    last[last[:, 0] < 0] = -1
    x = _planes(boards, last)
    w = glorot_init(S, seed=0)
    ow = onet.glorot_weights(S, seed=0)
    assert all(np.array_equal(w[k], ow[k]) for k in w)         # same "random-init" definition
    want_p, want_v = onet.OracleNet(S, w).eval(x)
    net = DeviceNet(S, 128, w, mode=mode)                      # 300 boards -> 3 chunks, one ragged
    got_p, got_v = net.eval(x)
    assert np.abs(got_p - want_p).max() <= TOL, np.abs(got_p - want_p).max()
    assert np.abs(got_v - want_v).max() <= TOL, np.abs(got_v - want_v).max()
    np.testing.assert_allclose(got_p.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("S", [11, 15])
def test_small_batch_latency_path_parity(cuda_lib, S):
    """A5_NET_SMALL (one persistent fp32 kernel, the leaf evaluator of a single Player search): exact fp32, so it
    sits within fp32 summation-order noise of the oracle -- well inside the 1e-4 bar -- for every batch 1..8,
    on the trained checkpoint (11x11) and random-init weights, and repeats bit for bit."""
    from alphafive_b200 import _lib
    from alphafive_b200.net import DeviceNet, glorot_init
    rng = np.random.default_rng(100 + S)
    if S == 11:
        g = golden("replay_sample.npz")
        x = _planes(g["boards"][:40], g["last_action"][:40])
        w = _ckpt()
    else:
        boards = np.stack([orules.random_board(rng, S) for _ in range(40)])
        x = _planes(boards, np.full((40, 2), -1))
        w = glorot_init(S, seed=0)
    want_p, want_v = onet.OracleNet(S, w).eval(x)
    net = DeviceNet(S, 16, w)
    xs = torch.from_numpy(np.ascontiguousarray(x > 0.5)).to(torch.int8).cuda()
    i = 0
    for n in (1, 2, 3, 8, 1, 5, 8, 4, 8):
        p, v = net.forward(xs[i:i + n], mode=_lib.NET_SMALL)
        p2, v2 = net.forward(xs[i:i + n], mode=_lib.NET_SMALL)
        assert torch.equal(p, p2) and torch.equal(v, v2)
        p, v = p.cpu().numpy(), v.cpu().numpy()
        assert np.abs(p - want_p[i:i + n]).max() <= 2e-5, (n, np.abs(p - want_p[i:i + n]).max())
        assert np.abs(v - want_v[i:i + n]).max() <= 2e-5, (n, np.abs(v - want_v[i:i + n]).max())
        np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-5)
        i += n
    from alphafive_b200._lib import A5Error
    with pytest.raises(A5Error):
        net.forward(xs[:9], mode=_lib.NET_SMALL)


def test_eval_is_the_reference_pv_fn_seam(cuda_lib):
    """ResNet.eval contract (network.py:90-97): f32 [B,3,S,S] -> (f32 [B,S*S], f32 [B])."""
    from alphafive_b200.net import DeviceNet
    net = DeviceNet(11, 8)
    x = np.zeros((1, 3, 11, 11), np.float32)
    p, v = net.eval(x)
    assert p.shape == (1, 121) and v.shape == (1,) and p.dtype == np.float32 and v.dtype == np.float32


@pytest.mark.parametrize("S,n", [(11, 1), (11, 300), (11, 1500), (15, 77), (9, 40)])
def test_chunk_major_megakernel_is_bitwise_identical(cuda_lib, S, n, monkeypatch):
    """A5_TC_MEGA=1 runs the eight block-conv launches as one depth-first persistent kernel (every CTA walks its own
    boards chunk by chunk through all layers).  Same MMAs in the same order per position, so the outputs
    are bit for bit those of the layer-per-launch path, ragged batches and all board sizes included."""
    from alphafive_b200.net import DeviceNet, glorot_init
    rng = np.random.default_rng(n)
    planes = torch.from_numpy((rng.random((n, 3, S, S)) < 0.25).astype(np.int8)).cuda()
    w = glorot_init(S, seed=2)
    monkeypatch.setenv("A5_TC_MEGA", "0")
    a = DeviceNet(S, n, w)
    monkeypatch.setenv("A5_TC_MEGA", "1")
    b = DeviceNet(S, n, w)
    pa, va = a.forward(planes)
    pb, vb = b.forward(planes)
    pb2, vb2 = b.forward(planes)
    assert torch.equal(pa, pb) and torch.equal(va, vb) and torch.equal(pb, pb2) and torch.equal(vb, vb2)
    a.close(); b.close()
