"""``NetworkAPI`` with the surface of the reference's ``genData.networkAPI.NetworkAPI``
(networkAPI.py:10-83): one daemon thread serves N ``multiprocessing`` Pipes, gathers
every pending request into one batch, evaluates it once on the device and scatters the
results back pipe by pipe.

Wire protocol (player.py:194-197 <-> networkAPI.py:49-78), unchanged:
  worker  -> server : a list of ``f32[3, S, S]`` input planes (the reference always sends one)
  server  -> worker : a list of ``(policy f32[S*S], float(value))`` of the same length
A pipe whose peer has gone away raises ``EOFError`` on ``recv``; it is logged, closed and --
unlike the reference, which keeps polling the dead handle -- dropped
from the wait set.
"""
from __future__ import annotations

from logging import getLogger
from multiprocessing import Pipe, connection
from threading import Thread

import numpy as np

logger = getLogger(__name__)


class NetworkAPI(object):
    def __init__(self, cfg=None, agent_model=None):
        self.agent_model = agent_model
        self.config = cfg
        self.pipes = []
        self.reload = True
        self.prediction_worker = None
        self.done = False

    def start(self, reload):
        self.reload = reload
        self.prediction_worker = Thread(target=self.predict_batch_worker, name="prediction_worker", daemon=True)
        self.prediction_worker.start()

    def get_pipe(self, reload=True):
        mine, yours = Pipe()
        self.pipes.append(mine)
        self.reload = reload
        return yours

    def _drain(self, pipe, requests):
        """Append (pipe, planes list) for every message waiting on ``pipe``."""
        try:
            while pipe.poll():
                requests.append((pipe, pipe.recv()))
        except (EOFError, OSError) as e:
            logger.error(f"EOF error: {e}")
            pipe.close()
            if pipe in self.pipes:
                self.pipes.remove(pipe)

    def predict_batch_worker(self):
        device_net = getattr(self.agent_model, "device_net", None)
        if device_net is not None:
            import torch
            torch.cuda.set_device(device_net.device)          # the thread inherits no CUDA context
        graph = getattr(self.agent_model, "graph", None)
        while not self.done:
            live = [p for p in self.pipes if not p.closed]
            if not live:
                self._sleep()
                continue
            try:
                ready = connection.wait(live, timeout=0.001)
            except OSError:
                continue
            if not ready:
                continue
            requests = []
            for pipe in ready:
                self._drain(pipe, requests)
            if not requests:
                continue
            batch = np.asarray([x for _, planes in requests for x in planes], dtype=np.float32)
            if graph is not None and hasattr(graph, "as_default"):
                with graph.as_default():
                    policy, value = self.agent_model.eval(batch)
            else:
                policy, value = self.agent_model.eval(batch)
            k = 0
            for pipe, planes in requests:
                reply = [(policy[k + j], float(value[k + j])) for j in range(len(planes))]
                k += len(planes)
                try:
                    pipe.send(reply)
                except (BrokenPipeError, OSError) as e:
                    logger.error(f"EOF error: {e}")

    @staticmethod
    def _sleep():
        import time
        time.sleep(0.001)

    def close(self):
        self.done = True
        for pipe in self.pipes:
            pipe.close()
