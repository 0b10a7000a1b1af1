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
    def __init__(self, P: int, topology_id: int, device, seed: int = 0, epochs: int = 20, batch_size: int = 50):
        self.P = P
        self.h1, self.h2 = SURROGATE_HIDDEN[int(topology_id)]
        self.device = device
        self.epochs, self.batch_size = epochs, batch_size
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.params: Optional[Dict[str, torch.Tensor]] = None
        self.train_mse = []

    def _init_params(self):
        def uni(*shape):
            return ((torch.rand(*shape, generator=self.gen) - 0.5) * 0.1).to(self.device).requires_grad_(True)
        z = lambda n: torch.zeros(n, device=self.device, requires_grad=True)
        return {"W1": uni(self.P, self.h1), "b1": z(self.h1), "W2": uni(self.h1, self.h2), "b2": z(self.h2),
                "W3": uni(self.h2, 1), "b3": z(1)}

    @staticmethod
    def forward(p, x):
        a = torch.relu(x @ p["W1"] + p["b1"])
        a = torch.relu(a @ p["W2"] + p["b2"])
        return torch.sigmoid(a @ p["W3"] + p["b3"]).squeeze(-1)

    def __call__(self, batch: np.ndarray, signature: int):
        """batch = [n, P+1] rows (theta', L_untempered) of one interval (device tensor or numpy). Returns numpy weights."""
        batch = torch.as_tensor(batch, device=self.device)
        X = batch[:, : self.P].to(torch.float32)
        Y = batch[:, self.P].to(torch.float64)
        min_y, max_y = 2.0 * float(Y.min().item()), -1.0            # _revamp.py:989-993 (every call)
        Yn = ((Y - min_y) / (max_y - min_y)).to(torch.float32)       # _revamp.py:1024
        if signature == 1 or self.params is None:                   # fresh model, else warm start (_revamp.py:247-255)
            self.params = self._init_params()
        params = self.params
        opt = torch.optim.Adam(list(params.values()), lr=1e-3, eps=1e-7)
        n = X.shape[0]
        n_val = int(n * 0.1)
        n_tr = n - n_val
        Xt, Yt, Xv, Yv = X[:n_tr], Yn[:n_tr], X[n_tr:], Yn[n_tr:]
        best, wait = float("inf"), 0
        for _ in range(self.epochs):
            perm = torch.randperm(n_tr, generator=self.gen).to(self.device)
            for s in range(0, n_tr, self.batch_size):
                idx = perm[s:s + self.batch_size]
                opt.zero_grad(set_to_none=True)
                loss = torch.mean((self.forward(params, Xt[idx]) - Yt[idx]) ** 2)
                loss.backward()
                opt.step()
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
        return {k: v.detach().cpu().numpy() for k, v in params.items()}
