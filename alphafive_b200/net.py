"""DeviceNet: the policy/value network on the GPU (a5_net_*), weights owned by PyTorch."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import NET_FP32, NET_SMALL, NET_SMALL_MAX, NET_TC, check, ptr, stream_ptr  # noqa: F401


def tensor_names():
    lib = _lib.load()
    return [lib.a5_net_tensor_name(i).decode() for i in range(_lib.NUM_TENSORS)]


def tensor_shapes(S: int) -> dict[str, tuple]:
    """TF-layout shapes of the 42 variables (network.py:58-88)."""
    Cc = S * S
    sh = {"bone/conv1/kernel": (5, 5, 3, 32), "bone/conv1/bias": (32,)}
    for name, cin, cout in (("bone/block1", 32, 64), ("bone/block2", 64, 128), ("value/block3", 128, 32),
                            ("policy/block4", 128, 64), ("policy/block5", 64, 32)):
        sh[f"{name}_res/kernel"] = (1, 1, cin, cout)
        sh[f"{name}_conv1/kernel"] = (3, 3, cin, cout)
        sh[f"{name}_conv2/kernel"] = (3, 3, cout, cout)
        for s in ("res", "conv1", "conv2"):
            sh[f"{name}_{s}/bias"] = (cout,)
    sh.update({"value/conv/kernel": (1, 1, 32, 4), "value/conv/bias": (4,),
               "value/fc1/kernel": (4 * Cc, 64), "value/fc1/bias": (64,),
               "value/fc2/kernel": (64, 1), "value/fc2/bias": (1,),
               "policy/conv/kernel": (1, 1, 32, 16), "policy/conv/bias": (16,),
               "policy/fc/kernel": (16 * Cc, Cc), "policy/fc/bias": (Cc,)})
    return sh


def glorot_init(S: int, seed: int = 0) -> dict[str, np.ndarray]:
    """Random-init weights as tf.layers leaves them: glorot-uniform kernels, zero biases
    (network.py:53-55; no initializer is passed anywhere)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in tensor_shapes(S).items():
        if name.endswith("bias"):
            out[name] = np.zeros(shape, np.float32)
            continue
        rf = int(np.prod(shape[:-2])) if len(shape) == 4 else 1
        lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
        out[name] = ((torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * lim).numpy()
    return out


class DeviceNet:
    """eval(planes) -> (prob, value) on the device.  ``mode``: ``NET_TC`` (default) is the tcgen05 path;
    ``NET_FP32`` is the exact-fp32 CUDA-core path kept as the on-device reference for debugging; ``NET_SMALL``
    (per call, at most ``NET_SMALL_MAX`` boards) is the one-kernel fp32 latency path a single ``Player`` search
    evaluates its leaves with."""

    def __init__(self, S: int, max_batch: int, weights: dict | None = None, mode: int = NET_TC, device=None):
        self.lib = _lib.load()
        self.S, self.C, self.max_batch, self.mode = S, S * S, max_batch, mode
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.a5_net_create(S, max_batch, C.byref(h)))
        self.handle = h
        self.names = tensor_names()
        self.version = 0                        # bumped by set_weights (evaluation caches key on it)
        self.params: dict[str, torch.Tensor] = {}
        self.set_weights(weights if weights is not None else glorot_init(S))

    def set_weights(self, weights: dict):
        shapes = tensor_shapes(self.S)
        params = {}
        for i, name in enumerate(self.names):
            w = torch.as_tensor(np.ascontiguousarray(weights[name]), dtype=torch.float32)
            assert tuple(w.shape) == shapes[name], (name, tuple(w.shape), shapes[name])
            assert w.numel() == self.lib.a5_net_tensor_size(i, self.S)
            params[name] = w.to(self.device).contiguous()
        arr = (C.c_void_p * _lib.NUM_TENSORS)(*[params[n].data_ptr() for n in self.names])
        with torch.cuda.device(self.device):
            check(self.lib.a5_net_set_weights(self.handle, arr, stream_ptr()))
        self.params = params          # PyTorch keeps ownership of the fp32 masters
        self.version += 1

    def forward(self, planes: torch.Tensor, prob: torch.Tensor | None = None, value: torch.Tensor | None = None,
                mode: int | None = None):
        """planes int8 [n, 3, S, S] (CUDA) -> prob f32 [n, S*S], value f32 [n]."""
        assert planes.is_cuda and planes.dtype == torch.int8
        n = planes.shape[0]
        planes = planes.contiguous()
        if prob is None:
            prob = torch.empty((n, self.C), dtype=torch.float32, device=planes.device)
        if value is None:
            value = torch.empty((n,), dtype=torch.float32, device=planes.device)
        self.forward_raw(planes.data_ptr(), n, prob, value, mode)
        return prob, value

    def forward_raw(self, planes_ptr: int, n: int, prob, value, mode=None):
        check(self.lib.a5_net_forward(self.handle, C.c_void_p(planes_ptr), n, ptr(prob), ptr(value),
                                      self.mode if mode is None else mode, stream_ptr()))

    def eval(self, inputs):
        """The reference's pv_fn seam (network.py:90-97): np.float32 [B, 3, S, S] host
        array in, (np.float32 [B, S*S], np.float32 [B]) out."""
        x = np.asarray(inputs)
        out_p, out_v = [], []
        for i in range(0, x.shape[0], self.max_batch):
            chunk = torch.from_numpy(np.ascontiguousarray(x[i:i + self.max_batch] > 0.5)).to(torch.int8)
            p, v = self.forward(chunk.to(self.device, non_blocking=True))
            out_p.append(p.cpu().numpy())
            out_v.append(v.cpu().numpy())
        return np.concatenate(out_p), np.concatenate(out_v)

    def close(self):
        if self.handle:
            self.lib.a5_net_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
