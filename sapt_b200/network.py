"""`Network` and `ptReplica` — the reference's model/likelihood surface on the GPU.

Same names, arguments and return conventions as _revamp.py:47-174 (`Network`) and
_revamp.py:380-410 (`ptReplica.rmse/accuracy/likelihood_func/prior_likelihood`); the arithmetic
runs in the CUDA kernels behind include/sapt.h. `data` arrays keep the reference layout
(float [n, D+1], label in column D). Data sets are uploaded once and cached by CONTENT (shape + a
digest of the bytes), so a temporary slice, a recycled `id()` or an in-place edit can never be
evaluated against a stale device copy; the cache is a small LRU whose evicted entries free their HBM.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .problem import Problem, prior as _prior


class Network:
    def __init__(self, Topo, Train, Test, learn_rate, device=None, path=_lib.PATH_AUTO):
        self.Top = [int(v) for v in Topo]
        self.TrainData = Train
        self.TestData = Test
        self.lrate = learn_rate
        self._path = path
        self._device = device
        self._problems = OrderedDict()  # content key -> (Problem, which); LRU, the constructor's two sets are pinned
        self._main = Problem(self.Top, Train, Test, device=device, path=path)
        self._pinned = {self._key(Train)}
        self._problems[self._key(Train)] = (self._main, 0)
        if Test is not None and len(Test):
            self._pinned.add(self._key(Test))
            self._problems.setdefault(self._key(Test), (self._main, 1))
        d, h, o = self.Top
        # same draw sequence as the reference constructor (_revamp.py:54-57)
        self.W1 = np.random.randn(d, h) / np.sqrt(d)
        self.B1 = np.random.randn(1, h) / np.sqrt(h)
        self.W2 = np.random.randn(h, o) / np.sqrt(h)
        self.B2 = np.random.randn(1, o) / np.sqrt(h)

    # -- a1 -------------------------------------------------------------------------------
    def decode(self, w):
        d, h, o = self.Top
        w = np.asarray(w)
        a, b = d * h, d * h + h * o
        self.W1 = np.reshape(w[0:a], (d, h))
        self.W2 = np.reshape(w[a:b], (h, o))
        self.B1 = w[b:b + h].reshape(1, h)
        self.B2 = w[b + h:b + h + o].reshape(1, o)

    def encode(self):
        return np.concatenate([self.W1.ravel(), self.W2.ravel(), self.B1.ravel(), self.B2.ravel()])

    # -- data-set cache -------------------------------------------------------------------
    MAX_CACHED = 8          # extra data sets kept on the device besides TrainData / TestData
    _FULL_DIGEST = 64 << 20  # arrays up to this size are digested in full, larger ones by a strided row sample + both ends

    @classmethod
    def _key(cls, data):
        a = np.asarray(data)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        h = hashlib.blake2b(digest_size=16)
        h.update(repr((a.shape, a.dtype.str)).encode())
        if a.nbytes <= cls._FULL_DIGEST or a.ndim < 1 or a.shape[0] < 4096:
            h.update(memoryview(a).cast("B"))
        else:
            step = max(1, a.shape[0] // 4096)
            h.update(np.ascontiguousarray(a[::step]).tobytes())
            h.update(np.ascontiguousarray(a[:64]).tobytes()); h.update(np.ascontiguousarray(a[-64:]).tobytes())
            h.update(np.asarray(a.sum(axis=0, dtype=np.float64)).tobytes())
        return h.digest()

    def _resolve(self, data):
        key = self._key(data)
        hit = self._problems.get(key)
        if hit is not None:
            self._problems.move_to_end(key)
            return hit
        entry = (Problem(self.Top, data, None, device=self._device, path=self._path), 0)
        self._problems[key] = entry
        extra = [k for k in self._problems if k not in self._pinned]
        while len(extra) > self.MAX_CACHED:
            old = extra.pop(0)
            pr, _ = self._problems.pop(old)
            if pr is not self._main:
                pr.close()
        return entry

    def _theta(self, w, pr):
        w = np.asarray(w, dtype=np.float32)
        single = w.ndim == 1
        t = torch.as_tensor(np.atleast_2d(w)).to(pr.device).contiguous()
        return t, single

    # -- a4 -------------------------------------------------------------------------------
    def evaluate_proposal(self, data, w):
        """-> (fx[n] float class indices, prob[n, O]) like _revamp.py:155-174."""
        pr, which = self._resolve(data)
        self.decode(w)
        t, _ = self._theta(w, pr)
        fx, prob = pr.evaluate_proposal(t[0], which)
        return fx.cpu().numpy().astype(np.float64), prob.cpu().numpy().astype(np.float64)

    def loglik_batch(self, data, W):
        """Batched extension: W[R, P] -> (L[R], ncorrect[R], sqerr[R]) numpy."""
        pr, which = self._resolve(data)
        t, _ = self._theta(W, pr)
        L, nc, sq = pr.loglik(t, which)
        return L.cpu().numpy(), nc.cpu().numpy(), sq.cpu().numpy()

    # -- a9 -------------------------------------------------------------------------------
    def langevin_gradient(self, data, w, depth, mode="sgd_exact", batch_scale=1.0):
        """One epoch (per unit of depth) of per-row online SGD, LG:138-156. w may be [P] or [R, P]."""
        pr, which = self._resolve(data)
        if which != 0:   # the Langevin kernels walk the TRAIN slot of a problem: give this array its own
            key = b"train:" + self._key(data)
            if key not in self._problems:
                self._problems[key] = (Problem(self.Top, data, None, device=self._device, path=self._path), 0)
                self._pinned.add(key)
            pr = self._problems[key][0]
        t, single = self._theta(w, pr)
        for _ in range(int(depth)):
            t = pr.langevin(t, self.lrate, 0 if mode == "sgd_exact" else 1, batch_scale)
        out = t.cpu().numpy().astype(np.float64)
        if single:
            self.decode(out[0])
            return out[0]
        return out


class ptReplica:
    """The numeric helpers of the reference's worker class, without the process: the chain loop
    itself is `PTEngine` (all chains at once)."""

    def __init__(self, topology, temperature=1.0):
        self.topology = [int(v) for v in topology]
        self.temperature = float(temperature)
        self.adapttemp = float(temperature)

    def rmse(self, pred, actual):
        return np.sqrt(((np.asarray(pred) - np.asarray(actual)) ** 2).mean())

    def accuracy(self, pred, actual):
        pred = np.asarray(pred); actual = np.asarray(actual)
        return 100 * (int(np.sum(pred == actual)) / pred.shape[0])

    def likelihood_func(self, fnn: Network, data, w):
        """-> [L/T, fx, rmse, L] (_revamp.py:391-402)."""
        y = np.asarray(data)[:, self.topology[0]]
        fx, _ = fnn.evaluate_proposal(data, w)
        L, nc, sq = fnn.loglik_batch(data, np.asarray(w)[None, :])
        lhood = float(L[0])
        return [lhood / self.adapttemp, fx, self.rmse(fx, y), lhood]

    def prior_likelihood(self, sigma_squared, nu_1, nu_2, w):
        """_revamp.py:404-410 (nu_1, nu_2 unused there too)."""
        P = len(w)
        t = torch.as_tensor(np.asarray(w, dtype=np.float32)[None, :]).cuda()
        return float(_prior(t, P, sigma_squared)[0].item())
