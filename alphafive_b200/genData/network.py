"""``ResNet`` with the inference surface of the reference's ``genData.network.ResNet``
(network.py:15-160), backed by the device network (``a5_net_*``).

What drivers touch: ``ResNet(board_size, graph=None)``, ``.eval`` (the ``pv_fn``
seam, network.py:90-97), ``.get_prob`` / ``.get_value`` (network.py:99-111), ``.restore``
(network.py:113-122: a directory resolves to its latest checkpoint, anything else is an exact
prefix, failure raises ``FileNotFoundError``), ``.get_pipes`` (network.py:124-134),
``.load_pretrained`` (network.py:136-155) and ``.close``.  The trainer-only TensorFlow objects
(``sess``, ``saver``, the loss tensors) are out of scope.

Weights are fp32 PyTorch tensors on the device; a fresh model is glorot-uniform with zero
biases, which is what ``tf.layers`` gives when no initializer is passed (network.py:53-55).
"""
from __future__ import annotations

import contextlib

import numpy as np

from .. import _lib, ckpt
from ..net import DeviceNet, glorot_init, tensor_shapes
from .networkAPI import NetworkAPI


class _Graph:
    """Stand-in for ``tf.Graph``: NetworkAPI does ``with agent_model.graph.as_default():``
    (networkAPI.py:67)."""

    def as_default(self):
        return contextlib.nullcontext(self)


class ResNet(object):
    def __init__(self, board_size, graph=None, max_batch=4096, mode=None, seed=0):
        self.board_size = board_size
        self.graph = graph if graph is not None else _Graph()
        if mode is None:
            mode = _lib.NET_TC
        self.device_net = DeviceNet(board_size, max_batch, glorot_init(board_size, seed), mode=mode)
        self.api = None

    # -- inference (network.py:90-111) -----------------------------------------------------
    def eval(self, inputs):
        return self.device_net.eval(inputs)

    def get_prob(self, inputs):
        return self.device_net.eval(inputs)[0]

    def get_value(self, inputs):
        return self.device_net.eval(inputs)[1]

    # -- weights (network.py:113-122, 136-155) ---------------------------------------------
    def restore(self, ckpt_path):
        try:
            prefix = ckpt.resolve_prefix(ckpt_path)
            weights = ckpt.read_bundle(prefix)
        except (OSError, KeyError, ValueError) as e:
            raise FileNotFoundError("Could not find old network weights") from e
        self.set_weights(weights)
        print("Successfully loaded:", prefix)

    def load_pretrained(self, data_path):
        """Assign every checkpoint variable whose name and shape match; report the rest."""
        found = ckpt.read_bundle(ckpt.resolve_prefix(data_path))
        shapes = tensor_shapes(self.board_size)
        cur = {k: v.cpu().numpy() for k, v in self.device_net.params.items()}
        ok, missed = [], []
        for name, value in found.items():
            if name in shapes and tuple(value.shape) == shapes[name]:
                cur[name] = value
                ok.append(name)
            else:
                missed.append(name)
        self.set_weights(cur)
        print("loaded successed: ")
        for v in ok:
            print(v)
        print("=" * 80)
        print("missed:")
        for v in missed:
            print(v)

    def save(self, save_path, global_step=None):
        """``self.saver.save(self.sess, save_path=..., global_step=...)`` (main.py:74-77): a TF checkpoint-v2
        bundle ``<save_path>-<step>.{index,data-00000-of-00001}`` + the ``checkpoint`` marker, loadable by
        the reference's ``restore`` (no ``.meta``: the reference rebuilds the graph in code)."""
        prefix = save_path if global_step is None else f"{save_path}-{global_step}"
        ckpt.write_bundle(prefix, {k: v.cpu().numpy() for k, v in self.device_net.params.items()})
        return prefix

    def set_weights(self, weights: dict):
        self.device_net.set_weights({k: np.asarray(v, np.float32) for k, v in weights.items()})

    # -- pipe server (network.py:124-134) --------------------------------------------------
    def get_pipes(self, config, reload=True):
        if self.api is None:
            self.api = NetworkAPI(config, self)
            self.api.start(reload)
        return self.api.get_pipe(reload)

    def close(self):
        if self.api is not None:
            self.api.close()
            self.api = None
        self.device_net.close()
