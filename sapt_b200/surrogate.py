"""Surrogate training off the hot path: PyTorch restatement of
`surrogate.train` + `ParallelTempering.surrogate_trainer/normalize_likelihood`
(_revamp.py:217-295, 976-1026). Keras is a third-party, un-pinned dependency of the reference
and is not installed here, so this part is "parity unpinned" against Keras itself; it follows
the call sites: Dense(h1, relu) -> Dense(h2, relu) -> Dense(1, sigmoid), kernel_initializer
'uniform' (U(-0.05, 0.05)), zero biases, adam + mse, batch 50, 20 epochs, validation_split 0.1
(the LAST 10 % of the rows, before shuffling), EarlyStopping(val_loss, patience 5), warm start
from the previous signature's weights (_revamp.py:247-266)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import SURROGATE_HIDDEN


class SurrogateTrainer:
    """`graph=True` (default on CUDA devices) replays ONE captured CUDA graph per full mini-batch — forward, backward
    and the Adam update — instead of ~40 eager launches; the arithmetic is the same."""

    def __init__(self, P: int, topology_id: int, device, seed: int = 0, epochs: int = 20, batch_size: int = 50, graph: Optional[bool] = None):
        self.P = P
        self.h1, self.h2 = SURROGATE_HIDDEN[int(topology_id)]
        self.device = torch.device(device)
        self.epochs, self.batch_size = epochs, batch_size
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.params: Optional[Dict[str, torch.Tensor]] = None
        self.train_mse = []
        self.use_graph = (self.device.type == "cuda") if graph is None else bool(graph)
        self._opt = None
        self._graph = None

    def _init_values(self):
        uni = lambda *shape: ((torch.rand(*shape, generator=self.gen) - 0.5) * 0.1).to(self.device)
        z = lambda n: torch.zeros(n, device=self.device)
        return {"W1": uni(self.P, self.h1), "b1": z(self.h1), "W2": uni(self.h1, self.h2), "b2": z(self.h2),
                "W3": uni(self.h2, 1), "b3": z(1)}

    @staticmethod
    def forward(p, x):
        a = torch.relu(x @ p["W1"] + p["b1"])
        a = torch.relu(a @ p["W2"] + p["b2"])
        return torch.sigmoid(a @ p["W3"] + p["b3"]).squeeze(-1)

    def _reset(self, fresh: bool):
        """Parameter tensors and the optimizer are created once (a captured graph holds their addresses); a fresh
        model re-draws the values in place, and every training starts from zeroed Adam moments like a newly
        compiled Keras model (_revamp.py:265)."""
        if self.params is None:
            self.params = {k: v.requires_grad_(True) for k, v in self._init_values().items()}
            self._opt = torch.optim.Adam(list(self.params.values()), lr=1e-3, eps=1e-7, capturable=self.use_graph)
        elif fresh:
            with torch.no_grad():
                for k, v in self._init_values().items():
                    self.params[k].copy_(v)
        for st in self._opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    def _step_eager(self, xb, yb):
        self._opt.zero_grad(set_to_none=True)
        loss = torch.mean((self.forward(self.params, xb) - yb) ** 2)
        loss.backward()
        self._opt.step()

    def _capture(self):
        B = self.batch_size
        self._xb = torch.zeros((B, self.P), dtype=torch.float32, device=self.device)
        self._yb = torch.zeros(B, dtype=torch.float32, device=self.device)
        saved = {k: v.detach().clone() for k, v in self.params.items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up on a side stream (allocator, Adam state, autograd buffers)
                self._step_eager(self._xb, self._yb)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        self._opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph):
            loss = torch.mean((self.forward(self.params, self._xb) - self._yb) ** 2)
            loss.backward()
            self._opt.step()
        with torch.no_grad():                        # undo the warm-up steps
            for k, v in saved.items():
                self.params[k].copy_(v)
        for st in self._opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    def __call__(self, batch: np.ndarray, signature: int):
        """batch = [n, P+1] rows (theta', L_untempered) of one interval (device tensor or numpy). Returns numpy weights."""
        batch = torch.as_tensor(batch, device=self.device)
        X = batch[:, : self.P].to(torch.float32)
        Y = batch[:, self.P].to(torch.float64)
        min_y, max_y = 2.0 * float(Y.min().item()), -1.0            # _revamp.py:989-993 (every call)
        Yn = ((Y - min_y) / (max_y - min_y)).to(torch.float32)       # _revamp.py:1024
        self._reset(fresh=(signature == 1 or self.params is None))  # fresh model, else warm start (_revamp.py:247-255)
        params = self.params
        if self.use_graph and self._graph is None:
            self._capture()
        n = X.shape[0]
        n_val = int(n * 0.1)
        n_tr = n - n_val
        Xt, Yt, Xv, Yv = X[:n_tr], Yn[:n_tr], X[n_tr:], Yn[n_tr:]
        best, wait = float("inf"), 0
        for _ in range(self.epochs):
            perm = torch.randperm(n_tr, generator=self.gen).to(self.device)
            for s in range(0, n_tr, self.batch_size):
                idx = perm[s:s + self.batch_size]
                if self._graph is not None and idx.numel() == self.batch_size:
                    self._xb.copy_(Xt[idx]); self._yb.copy_(Yt[idx])
                    self._graph.replay()
                else:
                    self._step_eager(Xt[idx], Yt[idx])
            if n_val > 0:
                with torch.no_grad():
                    vl = float(torch.mean((self.forward(params, Xv) - Yv) ** 2))
                if vl < best:
                    best, wait = vl, 0
                else:
                    wait += 1
                    if wait >= 5:
                        break
        with torch.no_grad():
            self.train_mse.append(float(torch.mean((self.forward(params, X) - Yn) ** 2)))  # _revamp.py:268-285
        return {k: v.detach().cpu().numpy().copy() for k, v in params.items()}   # copy: on a CPU device .numpy() aliases the parameter
