"""Device-resident data set + topology (wraps sapt_problem_* of include/sapt.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def current_stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Problem:
    """Uploads traindata/testdata once (reference arrays: float64 [n, D+1], label in column D,
    _revamp.py:1425-1519) and offers the batched device ops on them."""

    def __init__(self, topology, traindata, testdata=None, device=None, path=_lib.PATH_AUTO):
        if not torch.cuda.is_available():
            raise RuntimeError("sapt_b200 needs a CUDA device; there is no CPU fallback")
        self.topology = [int(v) for v in topology]
        D, H, O = self.topology
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.path = path
        L = _lib.lib()
        h = C.c_void_p()
        tr = np.ascontiguousarray(traindata, dtype=np.float64)
        if tr.ndim != 2 or tr.shape[1] != D + 1:
            raise ValueError(f"traindata must be [n, {D + 1}] (features + label column), got {tr.shape}")
        if testdata is not None and len(testdata):
            te = np.ascontiguousarray(testdata, dtype=np.float64)
            if te.ndim != 2 or te.shape[1] != D + 1:
                raise ValueError(f"testdata must be [n, {D + 1}], got {te.shape}")
        else:
            te = None
        _lib.check(L.sapt_problem_create(C.byref(h), D, H, O, tr.ctypes.data, tr.shape[0],
                                         te.ctypes.data if te is not None else None,
                                         te.shape[0] if te is not None else 0, self.device.index), "sapt_problem_create")
        self._h = h
        self.P = D * H + H * O + H + O
        self.n = (tr.shape[0], te.shape[0] if te is not None else 0)

    @classmethod
    def from_f32(cls, topology, xtrain, ytrain, xtest=None, ytest=None, device=None, path=_lib.PATH_AUTO):
        self = cls.__new__(cls)
        if not torch.cuda.is_available():
            raise RuntimeError("sapt_b200 needs a CUDA device; there is no CPU fallback")
        self.topology = [int(v) for v in topology]
        D, H, O = self.topology
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.path = path
        L = _lib.lib()
        h = C.c_void_p()
        xt = np.ascontiguousarray(xtrain, dtype=np.float32); yt = np.ascontiguousarray(ytrain, dtype=np.int32)
        nte = 0 if xtest is None else len(xtest)
        xs = np.ascontiguousarray(xtest, dtype=np.float32) if nte else None
        ys = np.ascontiguousarray(ytest, dtype=np.int32) if nte else None
        _lib.check(L.sapt_problem_create_f32(C.byref(h), D, H, O, xt.ctypes.data, yt.ctypes.data, xt.shape[0],
                                             xs.ctypes.data if nte else None, ys.ctypes.data if nte else None, nte,
                                             self.device.index), "sapt_problem_create_f32")
        self._h = h
        self.P = D * H + H * O + H + O
        self.n = (xt.shape[0], nte)
        return self

    @property
    def handle(self):
        return self._h

    @property
    def tc_supported(self) -> bool:
        return bool(_lib.lib().sapt_problem_tc_supported(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().sapt_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batched ops (device tensors) --------------------------------------------------
    def loglik(self, theta: torch.Tensor, which: int = 0, rows=None, n_rows=None, out=None, path=None):
        """theta fp32 [R, ld>=P] on this device -> (L f64 [R], ncorrect i32 [R], sqerr f64 [R])."""
        assert theta.dtype == torch.float32 and theta.is_cuda and theta.stride(1) == 1
        R = theta.shape[0]
        if out is None:
            out = (torch.zeros(R, dtype=torch.float64, device=theta.device),
                   torch.zeros(R, dtype=torch.int32, device=theta.device),
                   torch.zeros(R, dtype=torch.float64, device=theta.device))
        _lib.check(_lib.lib().sapt_loglik(self._h, which, _ptr(theta), _ld(theta), R, _ptr(rows), _ptr(n_rows),
                                          _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                          self.path if path is None else path, current_stream_ptr()), "sapt_loglik")
        return out

    def loglik_host(self, theta_host: np.ndarray, which: int = 0, path=None):
        """Host-buffer entry point (e2e): numpy fp32 [R, P] -> numpy (L, ncorrect, sqerr)."""
        th = np.ascontiguousarray(theta_host, dtype=np.float32)
        R = th.shape[0]
        L = np.empty(R, np.float64); nc = np.empty(R, np.int32); sq = np.empty(R, np.float64)
        _lib.check(_lib.lib().sapt_loglik_host(self._h, which, th.ctypes.data, th.shape[1], R, L.ctypes.data, nc.ctypes.data,
                                               sq.ctypes.data, self.path if path is None else path), "sapt_loglik_host")
        return L, nc, sq

    def evaluate_proposal(self, theta_one: torch.Tensor, which: int = 0):
        n = self.n[which]
        fx = torch.empty(n, dtype=torch.int32, device=theta_one.device)
        prob = torch.empty((n, self.topology[2]), dtype=torch.float32, device=theta_one.device)
        _lib.check(_lib.lib().sapt_evaluate_proposal(self._h, which, _ptr(theta_one), _ptr(fx), _ptr(prob), current_stream_ptr()),
                   "sapt_evaluate_proposal")
        return fx, prob

    def langevin(self, theta: torch.Tensor, lrate: float, mode: int = 0, batch_scale: float = 1.0, rows=None, n_rows=None, out=None):
        if out is None:
            out = theta.clone()
        _lib.check(_lib.lib().sapt_langevin(self._h, _ptr(theta), _ptr(out), _ld(theta), theta.shape[0], _ptr(rows), _ptr(n_rows),
                                            float(lrate), int(mode), float(batch_scale), current_stream_ptr()), "sapt_langevin")
        return out


def _ld(t: torch.Tensor) -> int:
    """Row stride in elements; a single-row tensor may report any stride for dim 0."""
    return max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] == 1 else int(t.stride(0))


def prior(theta: torch.Tensor, P: int, sigma_squared: float = 25.0):
    theta = theta.contiguous()
    out = torch.empty(theta.shape[0], dtype=torch.float64, device=theta.device)
    _lib.check(_lib.lib().sapt_prior(_ptr(theta), _ld(theta), theta.shape[0], P, float(sigma_squared), _ptr(out),
                                     current_stream_ptr()), "sapt_prior")
    return out
