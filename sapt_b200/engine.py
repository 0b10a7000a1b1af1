"""Device-resident parallel-tempering engine.

Replaces R x ``ptReplica.run`` (one OS process per chain, _revamp.py:412-780 /
surrogate_rw_common_interval.py:413-704 / surrogate_pt_classifier_langevingrad.py:428-908) plus
the coordinator loop of ``ParallelTempering.run_chains`` (RW:1068-1129) by one host thread per
GPU issuing a handful of kernels per MCMC step through the C-ABI (include/sapt.h). All chain
state (theta, temperatures, likelihoods, priors, windows, counters, histories, surrogate
training sets) lives in HBM; the host never reads it inside the loop.

Multi-GPU: chains (theta rows) are sharded contiguously over ranks; an exchange moves only the
slot binding, after a masked all-reduce of the per-chain scalars (torch.distributed, NCCL on
GPUs / gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, parallel
from .problem import Problem, _ptr, current_stream_ptr

SURROGATE_HIDDEN = {1: (64, 16), 2: (120, 40), 3: (200, 50)}  # _revamp.py:221-229


@dataclass
class VariantSpec:
    """Behavioural switches that differ between the reference's forks."""
    name: str
    swap: bool
    swap_at_zero: bool          # LG:804 swaps at i % swap_interval == 0 including i == 0
    post_tempered: bool         # LG:805 vs RW:623
    take_post_lik: bool         # LG:813-815
    gate_on_swap_interval: bool  # RW:507
    common_interval: bool       # RW: swap + surrogate rendezvous share i % swap_interval == 0, i != 0
    denorm_half_min: bool       # LG:325 vs RV:320
    compare_surrogate: bool     # RW:371 / RV:370 True, LG:381 False
    lg_flow: bool               # LG: redraw on surrogate, Langevin re-proposal, first-eval bookkeeping
    ship_zero_row: bool         # RW:496,626 / LG:543,822 ship surr_train_set with its all-zero first row
    local_surrogate: bool = False  # LS fork: one 4-10-1 network per chain over its own likelihood history


VARIANTS = {
    "rw": VariantSpec("rw", True, False, False, False, True, True, False, True, False, True),
    "lg": VariantSpec("lg", True, True, True, True, False, False, True, False, True, True),
    "rv": VariantSpec("rv", False, False, False, False, False, False, False, True, False, False),
    # localsurrogates/surrogate_pt_classifier_rw_localsurrogate.py: exchanges like LG (tempered L posted at
    # i % swap_interval == 0 including 0, LS:756-757) but only w is taken back (LS:765); the surrogate is gated on
    # surrogate_interval (LS:588) and always compared with the true train likelihood (LS:607-608)
    "ls": VariantSpec("ls", True, True, True, False, False, False, False, True, False, False, True),
}

LS_LRATE, LS_EPOCHS, LS_BLOCK = 0.1, 200, 5   # LS:262, LS:253, LS:548


@dataclass
class EngineConfig:
    variant: str = "rw"
    samples: int = 100                 # per chain (= int(NumSample / num_chains), _revamp.py:800)
    swap_interval: int = 50
    surrogate_interval: int = 50
    surrogate_prob: float = 0.5
    use_surrogate: bool = True
    use_langevin: bool = False
    learn_rate: float = 0.5
    langevin_mode: str = "sgd_exact"   # or "batch" 
    langevin_batch_scale: float = 1.0  # "sum": 1.0, "mean": 1/N
    surrogate_topology: int = 1
    seed: int = 1
    step_w: float = 0.025              # _revamp.py:449
    sigma_squared: float = 25.0        # _revamp.py:456
    l_prob: float = 0.5                # LG:385
    record_history: bool = True
    record_pos: bool = True
    train_cap: Optional[int] = None    # rows kept per chain per interval (default: the interval length)
    path: int = _lib.PATH_AUTO
    rank: int = 0
    world: int = 1


class Surrogate:
    """Device copy of the global surrogate MLP + the de-normalisation constants the workers hold."""

    def __init__(self, P: int, topology_id: int, device):
        self.h1, self.h2 = SURROGATE_HIDDEN[int(topology_id)]
        self.P = P
        self.device = device
        self.weights: Optional[Dict[str, torch.Tensor]] = None
        self.flat: Optional[torch.Tensor] = None
        self.min_y = 0.0
        self.max_y = 0.0
        self.signature = 0

    @property
    def shapes(self):
        return ((self.P, self.h1), (self.h1,), (self.h1, self.h2), (self.h2,), (self.h2, 1), (1,))

    def set_weights(self, w):
        """w: dict W1,b1,W2,b2,W3,b3 (numpy or tensors, Keras Dense layout). Stored as views into ONE flat fp32 device
        buffer, so that a later set of weights can be taken on the device without the host looking at it."""
        sizes = [int(np.prod(sh)) for sh in self.shapes]
        if self.flat is None:
            self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=self.device)
            views, o = {}, 0
            for k, n in zip(parallel.SURROGATE_WEIGHT_KEYS, sizes):
                views[k] = self.flat[o:o + n]
                o += n
            self._views = views
        for k, n in zip(parallel.SURROGATE_WEIGHT_KEYS, sizes):
            v = w[k]
            v = v.to(torch.float32) if torch.is_tensor(v) else torch.as_tensor(np.ascontiguousarray(v, dtype=np.float32))
            if v.numel() != n:
                raise ValueError(f"surrogate weight {k}: {v.numel()} values, expected {n} for P={self.P}, hidden {self.h1}/{self.h2}")
            self._views[k].copy_(v.reshape(-1).to(self.device))
        self.weights = self._views

    @property
    def ready(self) -> bool:
        return self.weights is not None


class TrainingBatch:
    """The rendezvous batch on the trainer rank: a preallocated fp64 [n_max, P+1] buffer and the number of valid rows as
    a DEVICE scalar. `tensor()` is the exact [n, P+1] view (one host read of n); a trainer that sets
    `accepts_lazy_batch = True` receives this object instead and decides itself whether to look."""

    def __init__(self, buf: torch.Tensor, n_dev: torch.Tensor):
        self.buf, self.n_dev, self._n = buf, n_dev, None

    @property
    def n(self) -> int:
        if self._n is None:
            self._n = int(self.n_dev.item())
        return self._n

    def tensor(self) -> torch.Tensor:
        return self.buf[: self.n]


class PTEngine:
    def __init__(self, problem: Problem, temperatures, cfg: EngineConfig, inject: Optional[dict] = None,
                 surrogate_trainer=None, process_group=None):
        self.p = problem
        self.cfg = cfg
        self.spec = VARIANTS[cfg.variant]
        self.dev = problem.device
        self.L = _lib.lib()
        self.Rg = int(len(temperatures))
        self.row0, self.R = parallel.shard_bounds(self.Rg, cfg.rank, cfg.world)
        self.P = problem.P
        self.ld = (self.P + 3) & ~3
        self.pg = process_group
        self.trainer = surrogate_trainer
        S = cfg.samples
        dev = self.dev
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.theta = torch.zeros((self.R, self.ld), dtype=torch.float32, device=dev)
        self.prop = torch.zeros_like(self.theta)
        self.theta_gd = torch.zeros_like(self.theta) if cfg.use_langevin else None
        self.T = torch.as_tensor(np.asarray(temperatures, dtype=np.float64)).to(dev)
        self.lik_cur = torch.zeros(self.Rg, **f64)
        self.lik_post = torch.zeros(self.Rg, **f64)
        self.prior_cur = torch.zeros(self.Rg, **f64)
        self.mov = torch.zeros((self.Rg, 3), **f64)
        self.likl_wo_temp = torch.zeros(self.Rg, **f64)
        self.last = torch.zeros((4, self.Rg), **f64)
        self.naccept = torch.zeros(self.Rg, **i32)
        self.slot_of_row = torch.arange(self.row0, self.row0 + self.R, **i32)
        self.row_of_slot = torch.arange(self.Rg, **i32)
        # per-row scratch
        self.L_train = torch.zeros(self.R, **f64); self.nc_train = torch.zeros(self.R, **i32); self.sq_train = torch.zeros(self.R, **f64)
        self.L_test = torch.zeros(self.R, **f64); self.nc_test = torch.zeros(self.R, **i32); self.sq_test = torch.zeros(self.R, **f64)
        self.prior_prop = torch.zeros(self.R, **f64)
        self.sur_s = torch.zeros(self.R, dtype=torch.float32, device=dev)
        self.is_sur = torch.zeros(self.R, **i32)
        self.surg_first = torch.zeros(self.R, **f64)
        self.true_rows = torch.zeros(self.R, **i32); self.sur_rows = torch.zeros(self.R, **i32); self.lang_rows = torch.zeros(self.R, **i32)
        self.counters = torch.zeros(4, **i32)  # n_true, n_sur, n_lang, num_swap
        # history [S][Rg]
        self.hist = None
        if cfg.record_history:
            self.hist = {k: torch.zeros((S, self.Rg), **f64) for k in
                         ("lhood_list", "rmse_train", "rmse_test", "acc_train", "acc_test", "likeh", "accept_list")}
            self.hist["surg"] = torch.zeros((S, self.Rg, 3), **f64)
            self.hist["accepted"] = torch.zeros((S, self.Rg), dtype=torch.uint8, device=dev)
            self.hist["is_sur"] = torch.zeros((S, self.Rg), dtype=torch.uint8, device=dev)
            if cfg.record_pos:
                self.hist["pos_w"] = torch.zeros((S, self.Rg, self.P), dtype=torch.float32, device=dev)
        # surrogate
        self.surrogate = Surrogate(self.P, cfg.surrogate_topology, dev)
        self.sweep_count = 0
        self.swap_log: List[torch.Tensor] = []
        self.total_swap_proposals = 0
        self.train_cap = 0
        self.train_theta = self.train_L = self.train_count = None
        self.ls = None
        if self.spec.local_surrogate and cfg.use_surrogate:
            if self.hist is None:
                raise ValueError("the local-surrogate fork reads its inputs from the surg_likeh_list history: record_history is required")
            if cfg.surrogate_interval < LS_BLOCK:
                raise ValueError("local surrogate: surrogate_interval must be >= 5 (LS:781 builds windows of five)")
            self.ls_weights = torch.zeros((self.Rg, _lib.LS_WEIGHTS), **f64)
            self.ls_limit = torch.zeros(self.Rg, **f64)
            self.ls_pred = torch.zeros(self.R, **f64); self.ls_avg = torch.zeros(self.R, **f64)
            self.ls_rmse = torch.zeros(self.Rg, **f64)
            self.ls_ready = False
            self.ls_synced_row = 0          # surg history rows < this are complete on every rank
            self.ls_train_rmse: List[torch.Tensor] = []
            self.ls = _lib.LocalSurrogate()
            self.ls.weights = self.ls_weights.data_ptr(); self.ls.limit = self.ls_limit.data_ptr()
            self.ls.surg = self.hist["surg"].data_ptr(); self.ls.n_slots = self.Rg
            self.ls.slot_of_row = self.slot_of_row.data_ptr(); self.ls.R = self.R
        if cfg.use_surrogate and not self.spec.local_surrogate:
            interval = cfg.swap_interval if self.spec.common_interval else cfg.surrogate_interval
            self.train_cap = int(cfg.train_cap if cfg.train_cap is not None else min(interval + 1, S))
            self.train_theta = torch.zeros((self.Rg, self.train_cap, self.P), dtype=torch.float32, device=dev)
            self.train_L = torch.zeros((self.Rg, self.train_cap), **f64)
            self.train_count = torch.zeros(self.Rg, **i32)
            # rows stay on the rank that wrote them (parallel.py "surrogate rendezvous"): which entries this rank owns, the
            # counts at the last synchronisation, and the fixed-size payload it ships to the trainer rank
            self.train_seen = torch.zeros(self.Rg, **i32)
            self.train_mine = torch.zeros((self.Rg, self.train_cap), dtype=torch.bool, device=dev)
            self.train_npad = self.R * max(self.train_cap, interval + 1)   # a rank appends <= R rows per step of the interval
            self.train_payload = torch.zeros((self.train_npad, self.P + 4), dtype=torch.float32, device=dev)
            self._batch_buf = None
        self.trainer_rank = 0
        self._train_synced = False
        self.surr_batches: List[torch.Tensor] = []   # device tensors of every rendezvous (what the parity tests compare)
        self.keep_batches = True                     # False: a batch is dropped once the trainer has seen it (long / huge runs)
        # injected streams (parity tests): numpy arrays indexed [step, chain(, P)]
        self.inj = None
        if inject is not None:
            self.inj = {}
            for k in ("z1", "z2"):
                if inject.get(k) is not None:
                    self.inj[k] = torch.as_tensor(np.ascontiguousarray(inject[k], dtype=np.float32)).to(dev)
            for k in ("ku", "u", "lx2", "swap_u"):
                if inject.get(k) is not None:
                    self.inj[k] = torch.as_tensor(np.ascontiguousarray(inject[k], dtype=np.float64)).to(dev)
            self.inj_weights = inject.get("surrogate_weights_per_signature")
            if inject.get("net_init") is not None:   # LS: the 61 create_network normals of every chain
                self.inj["net_init"] = torch.as_tensor(np.ascontiguousarray(inject["net_init"], dtype=np.float64)).to(dev)
        self.flags = 0
        sp = self.spec
        if sp.post_tempered: self.flags |= _lib.F_POST_TEMPERED
        if sp.take_post_lik: self.flags |= _lib.F_TAKE_POST_LIK
        if sp.denorm_half_min: self.flags |= _lib.F_DENORM_HALF_MIN
        if sp.compare_surrogate: self.flags |= _lib.F_COMPARE_SURROGATE
        if sp.lg_flow: self.flags |= _lib.F_REDRAW_ON_SURROGATE
        if sp.local_surrogate: self.flags |= _lib.F_LOCAL_SURROGATE
        if cfg.record_history and cfg.record_pos: self.flags |= _lib.F_RECORD_POS
        self.evals = 0  # replica x row likelihood evaluations issued (host-side count of the worst case)
        # device-side step / sweep counters: kernels read step = base + *counter, so a captured CUDA graph of
        # one exchange interval can be replayed for every later interval
        self.step_dev = torch.zeros(2, dtype=torch.int32, device=dev)
        self._graph_mode = False
        self._graphs = {}
        self._segments = [(0, self._hosted_mask())]  # (first history row, hosted mask) per binding epoch

    # ------------------------------------------------------------------------------------
    def _chain_state(self) -> _lib.ChainState:
        st = _lib.ChainState()
        st.slot_of_row = self.slot_of_row.data_ptr(); st.temperature = self.T.data_ptr()
        st.lik_cur = self.lik_cur.data_ptr(); st.lik_post = self.lik_post.data_ptr(); st.prior_cur = self.prior_cur.data_ptr()
        st.mov = self.mov.data_ptr(); st.likl_wo_temp = self.likl_wo_temp.data_ptr()
        st.last_rmse_train = self.last[0].data_ptr(); st.last_rmse_test = self.last[1].data_ptr()
        st.last_acc_train = self.last[2].data_ptr(); st.last_acc_test = self.last[3].data_ptr()
        st.naccept = self.naccept.data_ptr()
        return st

    def _history(self) -> _lib.History:
        h = _lib.History()
        h.n_slots = self.Rg
        if self.hist is not None:
            for k in ("lhood_list", "rmse_train", "rmse_test", "acc_train", "acc_test", "surg", "likeh", "accept_list", "accepted", "is_sur"):
                setattr(h, k, self.hist[k].data_ptr())
            if "pos_w" in self.hist:
                h.pos_w = self.hist["pos_w"].data_ptr()
        return h

    def _inj(self, key, step):
        if self.inj is None or key not in self.inj:
            return None
        return self.inj[key][step]

    # ------------------------------------------------------------------------------------
    def initialize(self, w0):
        """All chains start from the same w (_revamp.py:940-943); initial likelihood/prior as in
        ptReplica.run before the loop (_revamp.py:452-468)."""
        w0 = torch.as_tensor(np.asarray(w0, dtype=np.float32)).to(self.dev)
        self.theta.zero_()
        self.theta[:, : self.P] = w0[None, :]
        self.prop.copy_(self.theta)
        Ltr, _, _ = self.p.loglik(self.theta, 0, path=self.cfg.path)
        Lte = None
        if self.p.n[1] > 0:
            Lte, _, _ = self.p.loglik(self.theta, 1, path=self.cfg.path)
        pr = torch.empty(self.R, dtype=torch.float64, device=self.dev)
        _lib.check(self.L.sapt_prior(_ptr(self.theta), self.ld, self.R, self.P, self.cfg.sigma_squared, _ptr(pr), current_stream_ptr()))
        slots = self.slot_of_row.long()
        lik = Ltr / self.T[slots]
        self.lik_cur[slots] = lik
        self.lik_post[slots] = lik
        self.prior_cur[slots] = pr
        # `likl_without_temp` is overwritten by the TEST-set call during initialisation (RW:466-467)
        self.likl_wo_temp[slots] = Lte if Lte is not None else Ltr
        if self.hist is not None:
            self.hist["likeh"][0].fill_(-100.0)  # RW:473
            if "pos_w" in self.hist:
                self.hist["pos_w"][0].fill_(1.0)  # pos_w = np.ones (RW:431)

    # ------------------------------------------------------------------------------------
    def _propose_args(self, i, sdev, gate, use_sur, sur_ready) -> _lib.ProposeArgs:
        cfg, sp = self.cfg, self.spec
        a = _lib.ProposeArgs()
        a.base = self.theta.data_ptr(); a.proposal = self.prop.data_ptr(); a.ld = self.ld; a.R = self.R; a.P = self.P
        a.slot_of_row = self.slot_of_row.data_ptr(); a.slot_offset = 0
        a.step_w = cfg.step_w; a.sigma_squared = cfg.sigma_squared; a.prior_prop = self.prior_prop.data_ptr()
        a.seed = cfg.seed & 0xFFFFFFFF; a.kind = _lib.STREAM_Z1; a.step = i; a.step_dev = sdev
        z1 = self._inj("z1", i); a.z_inject = z1.data_ptr() if z1 is not None else None
        a.decide = 1 if use_sur else 0
        a.redraw = 1 if sp.lg_flow else 0
        a.surrogate_ready = 1 if sur_ready else 0
        a.surrogate_prob = cfg.surrogate_prob; a.gate = gate
        z2 = self._inj("z2", i); a.z2_inject = z2.data_ptr() if z2 is not None else None
        ku = self._inj("ku", i); a.ku_inject = ku.data_ptr() if ku is not None else None
        a.is_sur = self.is_sur.data_ptr(); a.true_rows = self.true_rows.data_ptr(); a.n_true = self.counters[0:1].data_ptr()
        a.sur_rows = self.sur_rows.data_ptr(); a.n_sur = self.counters[1:2].data_ptr()
        return a

    def _accept_args(self, i, sdev, use_sur) -> _lib.AcceptArgs:
        cfg, sp = self.cfg, self.spec
        ntr, nte = self.p.n
        m = _lib.AcceptArgs()
        m.theta = self.theta.data_ptr(); m.proposal = self.prop.data_ptr(); m.ld = self.ld; m.R = self.R; m.P = self.P
        m.st = self._chain_state(); m.hist = self._history(); m.step = i; m.step_dev = sdev; m.flags = self.flags
        m.is_sur = self.is_sur.data_ptr() if use_sur else None
        m.L_train = self.L_train.data_ptr(); m.nc_train = self.nc_train.data_ptr(); m.sq_train = self.sq_train.data_ptr()
        m.L_test = self.L_test.data_ptr(); m.nc_test = self.nc_test.data_ptr(); m.sq_test = self.sq_test.data_ptr()
        m.n_train = ntr; m.n_test = nte
        m.prior_prop = self.prior_prop.data_ptr(); m.sur_s = self.sur_s.data_ptr()
        m.min_y = self.surrogate.min_y; m.max_y = self.surrogate.max_y
        m.seed = cfg.seed & 0xFFFFFFFF
        u = self._inj("u", i); m.u_inject = u.data_ptr() if u is not None else None
        if self.train_theta is not None:
            m.train_theta = self.train_theta.data_ptr(); m.train_L = self.train_L.data_ptr()
            m.train_count = self.train_count.data_ptr(); m.train_cap = self.train_cap
        m.capture = 0 if sp.lg_flow else 1
        m.surg_first = self.surg_first.data_ptr() if sp.lg_flow else None
        if self.ls is not None:
            m.ls_pred = self.ls_pred.data_ptr(); m.ls_avg = self.ls_avg.data_ptr()
        return m

    def step(self, i: int):
        """Enqueue one MCMC step of every local chain (no host synchronisation). In graph mode `i` is the
        offset inside the captured segment and the absolute step lives in self.step_dev[0]."""
        cfg, sp, L = self.cfg, self.spec, self.L
        st = current_stream_ptr()
        sdev = self.step_dev[0:1].data_ptr() if self._graph_mode else None
        if self._graph_mode:
            i = 0  # the device counter is advanced after every captured step: the absolute step is *step_dev
        use_sur = bool(cfg.use_surrogate)
        self._train_synced = False     # this step may append training rows on the hosting rank
        self.counters[:3].zero_()
        gate = cfg.swap_interval if sp.gate_on_swap_interval else cfg.surrogate_interval
        sur_ready = self.ls_ready if self.ls is not None else self.surrogate.ready
        a = self._propose_args(i, sdev, gate, use_sur, sur_ready)
        z2 = self._inj("z2", i)
        _lib.check(L.sapt_propose(C.byref(a), st), "sapt_propose")

        sur_possible = use_sur and sur_ready and cfg.surrogate_prob > 0 and (self._graph_mode or i >= gate + 1)
        true_rows = self.true_rows if use_sur else None
        n_true = self.counters[0:1] if use_sur else None
        if sur_possible and self.ls is not None:
            _lib.check(L.sapt_local_surrogate_predict(C.byref(self.ls), i, sdev, _ptr(self.sur_rows), _ptr(self.counters[1:2]),
                                                      _ptr(self.ls_pred), _ptr(self.ls_avg), st), "sapt_local_surrogate_predict")
        elif sur_possible:
            w = self.surrogate.weights
            _lib.check(L.sapt_surrogate_predict(_ptr(self.prop), self.ld, self.R, self.P, _ptr(self.sur_rows), _ptr(self.counters[1:2]),
                                                _ptr(w["W1"]), _ptr(w["b1"]), self.surrogate.h1, _ptr(w["W2"]), _ptr(w["b2"]),
                                                self.surrogate.h2, _ptr(w["W3"]), _ptr(w["b3"]), _ptr(self.sur_s), st), "sapt_surrogate_predict")
        out_tr = (self.L_train, self.nc_train, self.sq_train)
        out_te = (self.L_test, self.nc_test, self.sq_test)
        ntr, nte = self.p.n
        if sp.lg_flow:
            self.p.loglik(self.prop, 0, rows=true_rows, n_rows=n_true, out=out_tr, path=cfg.path)
            f = _lib.LgFirstArgs()
            f.proposal = self.prop.data_ptr(); f.ld = self.ld; f.R = self.R; f.P = self.P
            f.st = self._chain_state(); f.step = i; f.step_dev = sdev; f.seed = cfg.seed & 0xFFFFFFFF
            f.true_rows = _ptr(true_rows); f.n_true = _ptr(n_true)
            f.L_train = self.L_train.data_ptr(); f.surg_first = self.surg_first.data_ptr()
            f.use_langevin = 1 if cfg.use_langevin else 0; f.l_prob = cfg.l_prob
            lx2 = self._inj("lx2", i); f.lx2_inject = lx2.data_ptr() if lx2 is not None else None
            f.lang_rows = self.lang_rows.data_ptr(); f.n_lang = self.counters[2:3].data_ptr()
            if self.train_theta is not None:
                f.train_theta = self.train_theta.data_ptr(); f.train_L = self.train_L.data_ptr()
                f.train_count = self.train_count.data_ptr(); f.train_cap = self.train_cap
            _lib.check(L.sapt_lg_first_eval(C.byref(f), st), "sapt_lg_first_eval")
            self.evals += self.R * ntr
            if cfg.use_langevin:
                mode = 0 if cfg.langevin_mode == "sgd_exact" else 1
                # Eq. 8: w_gd = langevin_gradient(w) (LG:650); the second call on w_proposal (LG:652) only
                # feeds diff_prop, which never reaches the MH ratio (LG:665-668 vs 687-691) and is skipped.
                self.p.langevin(self.theta, cfg.learn_rate, mode, cfg.langevin_batch_scale, rows=self.lang_rows,
                                n_rows=self.counters[2:3], out=self.theta_gd)
                b = _lib.ProposeArgs()
                b.base = self.theta_gd.data_ptr(); b.proposal = self.prop.data_ptr(); b.ld = self.ld; b.R = self.R; b.P = self.P
                b.slot_of_row = self.slot_of_row.data_ptr(); b.rows = self.lang_rows.data_ptr(); b.n_rows = self.counters[2:3].data_ptr()
                b.step_w = cfg.step_w; b.sigma_squared = cfg.sigma_squared; b.prior_prop = self.prior_prop.data_ptr()
                b.seed = cfg.seed & 0xFFFFFFFF; b.kind = _lib.STREAM_Z2; b.step = i; b.step_dev = sdev
                b.z_inject = z2.data_ptr() if z2 is not None else None
                b.decide = 0
                _lib.check(L.sapt_propose(C.byref(b), st), "sapt_propose(langevin)")
                # LG:676 re-evaluates every true row; only the Langevin rows have a new proposal, the
                # others would reproduce the first evaluation bit for bit.
                self.p.loglik(self.prop, 0, rows=self.lang_rows, n_rows=self.counters[2:3], out=out_tr, path=cfg.path)
                self.evals += self.R * ntr // 2
            if nte > 0:
                self.p.loglik(self.prop, 1, rows=true_rows, n_rows=n_true, out=out_te, path=cfg.path)
                self.evals += self.R * nte
        else:
            if sp.compare_surrogate or not use_sur:
                self.p.loglik(self.prop, 0, out=out_tr, path=cfg.path)   # RV:543-544: train set on every step
            else:
                self.p.loglik(self.prop, 0, rows=true_rows, n_rows=n_true, out=out_tr, path=cfg.path)
            self.evals += self.R * ntr
            if nte > 0:
                self.p.loglik(self.prop, 1, rows=true_rows, n_rows=n_true, out=out_te, path=cfg.path)
                self.evals += self.R * nte

        m = self._accept_args(i, sdev, use_sur)
        _lib.check(L.sapt_mh_accept(C.byref(m), st), "sapt_mh_accept")

    # ------------------------------------------------------------------------------------
    def _hosted_mask(self):
        return parallel.hosted_mask(self.slot_of_row, self.Rg)

    def sync_slots(self):
        """Masked all-reduce of the per-chain scalars: every rank contributes the slots its rows host."""
        if self.cfg.world == 1:
            return
        args = (self.lik_cur, self.lik_post, self.prior_cur, self.likl_wo_temp, self.last, self.mov, self.naccept)
        pack = parallel.sync_hosted(parallel.pack_state(*args), self._hosted_mask(), self.pg)
        parallel.unpack_state(pack, *args)
        if self.ls is not None:   # the local network and its limit belong to the chain (slot), not to the theta row
            net = parallel.sync_hosted(torch.cat([self.ls_weights.t(), self.ls_limit[None]], dim=0), self._hosted_mask(), self.pg)
            self.ls_weights.copy_(net[:-1].t()); self.ls_limit.copy_(net[-1])

    def swap(self, i: int):
        """Replica exchange at a rendezvous (RW:1095-1099)."""
        if self.Rg < 2:
            return
        self.sync_slots()
        self._sync_training_counts()
        if self.cfg.world > 1 and self.hist is not None and "pos_w" in self.hist:
            # a rejected step copies pos_w history row i (RV:608): the row must follow the slot across ranks
            import torch.distributed as dist
            row = self.hist["pos_w"][i + 1]
            mask = self._hosted_mask().to(row.dtype)
            row.mul_(mask[:, None])
            dist.all_reduce(row, group=self.pg)
        u = None
        if self.inj is not None and "swap_u" in self.inj:
            u = self.inj["swap_u"][self.sweep_count]
        swapped = torch.zeros(self.Rg - 1, dtype=torch.uint8, device=self.dev)
        _lib.check(self.L.sapt_swap_sweep(self.Rg, _ptr(self.lik_post), _ptr(self.T), _ptr(self.row_of_slot), self.flags,
                                          self.cfg.seed & 0xFFFFFFFF, self.sweep_count, _ptr(u), _ptr(swapped),
                                          _ptr(self.counters[3:4]), None, self.row0, self.R, _ptr(self.slot_of_row),
                                          current_stream_ptr()), "sapt_swap_sweep")
        if self.ls is not None and self.cfg.world > 1:
            self._ls_sync_history(i + 2)
        self.swap_log.append(swapped)
        self._segments.append((i + 2, self._hosted_mask()))  # history rows >= i+2 are written under the new binding
        self.sweep_count += 1
        self.total_swap_proposals += self.Rg - 1

    def _sync_training_counts(self):
        """A chain's training rows are appended on whichever rank hosts the chain, and they STAY there. What follows the
        chain is its row count: before an exchange re-binds the chains (and before a rendezvous) every rank marks the
        entries it has written since the last call as its own and the counts are taken from the current hosts (one
        [R_global] int32 all-reduce), so the next host continues at the right index."""
        if self.train_theta is None or self._train_synced:
            return
        mask = self._hosted_mask()
        parallel.mark_owned(self.train_mine, self.train_seen, self.train_count, mask)
        if self.cfg.world > 1:
            parallel.sync_train_counts(self.train_count, self.train_seen, mask, self.pg)
        else:
            self.train_seen.copy_(self.train_count)
        self._train_synced = True

    def gather_training_rows(self):
        """(theta', L_untempered) rows of the interval's true steps, chain by chain (RW:1105-1116), as ONE device tensor
        [n, P+1] (fp64) on the trainer rank: the training set never visits the host (the reference pickles it through
        queues). Every rank packs the rows it owns with their final batch position (sapt_train_rows_pack), the fixed-size
        payloads are gathered on the trainer rank and scattered into the batch (sapt_train_rows_scatter). One GPU: all shapes
        are static, no host synchronisation; several ranks: one host read per rendezvous (how many payload rows carry data), so
        that the padding does not travel. Returns a TrainingBatch on the trainer rank, None elsewhere; `self.last_n_out`
        is the device scalar with the number of rows (known on every rank)."""
        self._sync_training_counts()
        src, dest, zero_rows, n_out, n_owned = parallel.training_row_plan(self.train_count, self.train_mine, self.train_npad,
                                                                          skip_first=(self.spec.name == "rv"),      # RV ships surr_train_set[1:count_real] (RV:683)
                                                                          zero_row=self.spec.ship_zero_row,         # RW:496,626 / LG:543,822
                                                                          return_owned=True)
        self.last_n_out = n_out
        st = current_stream_ptr()
        _lib.check(self.L.sapt_train_rows_pack(_ptr(self.train_theta), _ptr(self.train_L), _ptr(src), _ptr(dest), self.train_npad, self.P,
                                               _ptr(self.train_payload), st), "sapt_train_rows_pack")
        allp = parallel.gather_payloads(self.train_payload, self.trainer_rank, self.cfg.world, self.pg, n_owned=n_owned)
        batch = None
        if self.cfg.rank == self.trainer_rank:
            if self._batch_buf is None:
                zero = 1 if self.spec.ship_zero_row else 0
                self._batch_buf = torch.zeros((self.Rg * (self.train_cap + zero), self.P + 1), dtype=torch.float64, device=self.dev)
            _lib.check(self.L.sapt_train_rows_scatter(_ptr(allp), allp.shape[0], self.P, _ptr(zero_rows) if zero_rows.numel() else None,
                                                      int(zero_rows.numel()), _ptr(self._batch_buf), st), "sapt_train_rows_scatter")
            batch = TrainingBatch(self._batch_buf, n_out)
        self.train_count.zero_(); self.train_seen.zero_(); self.train_mine.zero_()
        self._train_synced = False
        return batch

    def _ls_sync_history(self, upto: int):
        """Multi-rank LS: a chain's surg_likeh_list rows live on whichever rank hosted the chain when they were
        written. Between two exchanges the binding is fixed, so in the rows written since the last sync every
        (row, chain) entry is non-zero on exactly one rank: a plain sum makes them complete everywhere (NaN
        markers survive the sum), and the new host of a chain can feed its network."""
        lo = self.ls_synced_row
        if upto > lo:
            parallel.sync_history_rows(self.hist["surg"][lo:upto], self.pg)
            self.ls_synced_row = upto

    def local_surrogate_train(self, i: int):
        """LS:771-832 for every local chain: first call = series surg[0:i, 2], limit, fresh network; later calls
        = the last surrogate_interval entries, same limit, training continues."""
        first = not self.ls_ready
        lo = 0 if first else i - self.cfg.surrogate_interval
        init = self.inj.get("net_init") if (self.inj is not None and first) else None
        _lib.check(self.L.sapt_local_surrogate_train(C.byref(self.ls), lo, i, 1 if first else 0, _ptr(init), self.cfg.seed & 0xFFFFFFFF,
                                                     LS_LRATE, LS_EPOCHS, _ptr(self.ls_rmse), current_stream_ptr()),
                   "sapt_local_surrogate_train")
        self.ls_train_rmse.append(self.ls_rmse.clone())
        self.ls_ready = True

    def surrogate_rendezvous(self, i: int):
        """RW:1105-1116 / RV:976-1026: collect the interval's rows, retrain, hand the model back to the chains. With
        several ranks the rows go to ONE trainer rank and only the weights (and, once, minY / maxY) come back."""
        if self.ls is not None:
            return self.local_surrogate_train(i)
        batch = self.gather_training_rows()
        if not self.spec.ship_zero_row and int(self.last_n_out.item()) == 0:
            return            # RV ships no zero rows: an interval of surrogate steps only leaves nothing to train on
        sur = self.surrogate
        sur.signature += 1
        first = sur.signature == 1
        multi = self.cfg.world > 1
        w, min_y, max_y = None, sur.min_y, sur.max_y
        if self.cfg.rank == self.trainer_rank:
            if self.keep_batches:
                self.surr_batches.append(batch.tensor().clone())
            if first:  # minmax.txt is written once and read once (RV:996-997, RV:707-710)
                min_y, max_y = float(torch.min(batch.tensor()[:, self.P]).item() * 2), -1.0
            if self.inj is not None and self.inj_weights is not None:
                w = self.inj_weights[sur.signature - 1]
            elif self.trainer is not None:
                w = self.trainer(batch if getattr(self.trainer, "accepts_lazy_batch", False) else batch.tensor(), sur.signature)
        if not multi:
            sur.min_y, sur.max_y = min_y, max_y
            if w is not None:
                sur.set_weights(w)
            return
        head, flat = parallel.broadcast_surrogate_device(w, sur.shapes, min_y, max_y, self.trainer_rank, self.dev, self.pg)
        if first or not sur.ready:
            h = head.cpu()     # the only host read: until a model exists every rank must learn whether one arrived (and minY / maxY once)
            if first:
                sur.min_y, sur.max_y = float(h[1]), float(h[2])
            if h[0] > 0:
                sur.set_weights(parallel.split_flat(flat, sur.shapes))
        else:                  # later models are taken on the device: new weights if the trainer produced any, else the old ones
            sur.flat.copy_(torch.where(head[0] > 0, flat, sur.flat))

    def rendezvous(self, i: int):
        cfg, sp = self.cfg, self.spec
        do_swap = sp.swap and (i % cfg.swap_interval == 0) and (sp.swap_at_zero or i != 0)
        if sp.common_interval:
            do_train = do_swap and cfg.use_surrogate
        else:
            do_train = cfg.use_surrogate and (i % cfg.surrogate_interval == 0) and (i != 0)
        if do_swap:
            self.swap(i)
        if do_train:
            self.surrogate_rendezvous(i)

    # ------------------------------------------------------------------------------------
    def fused_supported(self) -> bool:
        """Small problems (the reference's own runnable cases) run n steps per launch in the fused kernel."""
        return (not self.spec.lg_flow and self.ls is None and self.inj is None and (self.p.n[0] + self.p.n[1]) <= 8192
                and self.p.topology[0] <= 64 and self.p.topology[2] <= 20 and self.p.topology[1] <= 128)

    def fused_steps(self, i0: int, n: int):
        """Steps i0 .. i0+n-1 of every local chain in ONE launch (sapt_fused_steps)."""
        cfg, sp = self.cfg, self.spec
        self._train_synced = False
        f = _lib.FusedArgs()
        f.theta = self.theta.data_ptr(); f.ld = self.ld; f.R = self.R; f.P = self.P
        f.step_w = cfg.step_w; f.sigma_squared = cfg.sigma_squared; f.seed = cfg.seed & 0xFFFFFFFF
        f.use_surrogate = 1 if cfg.use_surrogate else 0
        f.surrogate_ready = 1 if self.surrogate.ready else 0
        f.surrogate_prob = cfg.surrogate_prob
        f.gate = cfg.swap_interval if sp.gate_on_swap_interval else cfg.surrogate_interval
        f.h1 = self.surrogate.h1; f.h2 = self.surrogate.h2
        if self.surrogate.ready:
            w = self.surrogate.weights
            f.W1 = w["W1"].data_ptr(); f.b1 = w["b1"].data_ptr(); f.W2 = w["W2"].data_ptr(); f.b2 = w["b2"].data_ptr()
            f.W3 = w["W3"].data_ptr(); f.b3 = w["b3"].data_ptr()
        m = f.acc
        m.theta = self.theta.data_ptr(); m.proposal = self.prop.data_ptr(); m.ld = self.ld; m.R = self.R; m.P = self.P
        m.st = self._chain_state(); m.hist = self._history(); m.step = i0; m.flags = self.flags
        m.min_y = self.surrogate.min_y; m.max_y = self.surrogate.max_y; m.seed = cfg.seed & 0xFFFFFFFF
        if self.train_theta is not None:
            m.train_theta = self.train_theta.data_ptr(); m.train_L = self.train_L.data_ptr()
            m.train_count = self.train_count.data_ptr(); m.train_cap = self.train_cap
        f.n_steps = n
        _lib.check(self.L.sapt_fused_steps(self.p.handle, C.byref(f), current_stream_ptr()), "sapt_fused_steps")
        ntr, nte = self.p.n
        self.evals += n * self.R * (ntr + nte)

    def native_supported(self) -> bool:
        """sapt_step drives whole steps of the random-walk forks inside the library (no Python between the kernels)."""
        return not self.spec.lg_flow and self.ls is None and self.inj is None

    def native_steps(self, i0: int, n: int):
        """Steps i0 .. i0+n-1 of every local chain through ONE C call (sapt_step): same kernels, same order as step()."""
        cfg, sp = self.cfg, self.spec
        use_sur = bool(cfg.use_surrogate)
        self._train_synced = False
        gate = cfg.swap_interval if sp.gate_on_swap_interval else cfg.surrogate_interval
        sa = _lib.StepArgs()
        sa.prop = self._propose_args(i0, None, gate, use_sur, self.surrogate.ready)
        sa.acc = self._accept_args(i0, None, use_sur)
        sa.use_surrogate = 1 if use_sur else 0
        sa.compare_surrogate = 1 if sp.compare_surrogate else 0
        sa.h1 = self.surrogate.h1; sa.h2 = self.surrogate.h2
        if self.surrogate.ready:
            w = self.surrogate.weights
            sa.W1 = w["W1"].data_ptr(); sa.b1 = w["b1"].data_ptr(); sa.W2 = w["W2"].data_ptr(); sa.b2 = w["b2"].data_ptr()
            sa.W3 = w["W3"].data_ptr(); sa.b3 = w["b3"].data_ptr()
        sa.sur_s = self.sur_s.data_ptr()
        sa.L_train = self.L_train.data_ptr(); sa.nc_train = self.nc_train.data_ptr(); sa.sq_train = self.sq_train.data_ptr()
        sa.L_test = self.L_test.data_ptr(); sa.nc_test = self.nc_test.data_ptr(); sa.sq_test = self.sq_test.data_ptr()
        sa.counters = self.counters.data_ptr(); sa.path = cfg.path; sa.n_steps = n
        _lib.check(self.L.sapt_step(self.p.handle, C.byref(sa), current_stream_ptr()), "sapt_step")
        ntr, nte = self.p.n
        self.evals += n * self.R * (ntr + nte)

    def _run_native(self, n):
        pts = self._rendezvous_steps(n)
        i = 0
        while i < n:
            j = i
            while j not in pts and j < n - 1:
                j += 1
            self.native_steps(i, j - i + 1)
            self.rendezvous(j)
            i = j + 1

    def _run_fused(self, n):
        pts = self._rendezvous_steps(n)
        i = 0
        while i < n:
            j = i
            while j not in pts and j < n - 1:
                j += 1
            self.fused_steps(i, j - i + 1)
            self.rendezvous(j)
            i = j + 1

    def run(self, steps: Optional[int] = None, graph: Optional[bool] = None, fused: Optional[bool] = None, native: Optional[bool] = None):
        """Run the chain loop. `graph=True` captures the launches of one interval between rendezvous points in a
        CUDA graph and replays it (the step index is read from device memory), removing the per-launch host cost.
        Measured at C1 (8 chains x 699 rows) the eager loop is already GPU-latency bound (~67 us per step for ~10
        dependent small kernels) and capture + instantiation only pays off for runs of >~10^4 steps, so the
        default is off; the mode is kept for host-bound deployments (slow CPUs, many short kernels).
        `native` (default: on where it applies) runs every stretch between two rendezvous through ONE sapt_step call —
        the C driver of the random-walk forks' loop body — instead of ~6 ctypes calls per step."""
        n = self.cfg.samples - 1 if steps is None else int(steps)
        if self.hist is not None and n > self.cfg.samples - 1:
            # the kernels write history row step+1 with no bound check: the buffers hold cfg.samples rows
            raise ValueError(f"run(steps={n}) with record_history needs samples >= steps + 1 (samples = {self.cfg.samples})")
        if fused is None:
            fused = self.fused_supported()
        if fused:
            if not self.fused_supported():
                raise NotImplementedError("fused step kernel: random-walk forks, Philox streams, <= 8192 rows only")
            return self._run_fused(n)
        if graph is None:
            graph = False
        if not graph or self.inj is not None or self.cfg.world != 1:
            if native is None:
                native = self.native_supported()
            if native:
                if not self.native_supported():
                    raise NotImplementedError("sapt_step drives the random-walk forks with Philox streams only")
                return self._run_native(n)
            for i in range(n):
                self.step(i)
                self.rendezvous(i)
            return
        self._run_graphed(n)

    def _rendezvous_steps(self, n):
        cfg, sp = self.cfg, self.spec
        pts = set()
        for i in range(n):
            do_swap = sp.swap and (i % cfg.swap_interval == 0) and (sp.swap_at_zero or i != 0)
            do_train = (do_swap and cfg.use_surrogate) if sp.common_interval else (
                cfg.use_surrogate and (i % cfg.surrogate_interval == 0) and (i != 0))
            if do_swap or do_train:
                pts.add(i)
        return pts

    def _run_graphed(self, n):
        pts = self._rendezvous_steps(n)
        i = 0
        self.step(0)                       # eager warm-up step: grows workspaces outside any capture
        self.rendezvous(0)
        i = 1
        while i < n:
            j = i                          # segment [i, j]: ends at the next rendezvous step or at n-1
            while j not in pts and j < n - 1:
                j += 1
            length = j - i + 1
            # everything a captured step bakes in: the segment length, the global surrogate's weights (signature) and
            # whether surrogate steps exist at all — for the local-surrogate fork that is ls_ready, which the
            # rendezvous flips without touching the signature
            ready = self.ls_ready if self.ls is not None else self.surrogate.ready
            key = (length, self.surrogate.signature, bool(ready))
            g = self._graphs.get(key)
            self.step_dev[0] = i
            if g is None and length >= 4:
                if ready and self.ls is None:   # library workspaces grow on first use: that must not happen under capture
                    w = self.surrogate.weights
                    _lib.check(self.L.sapt_surrogate_predict(_ptr(self.prop), self.ld, self.R, self.P, None, None, _ptr(w["W1"]), _ptr(w["b1"]),
                                                             self.surrogate.h1, _ptr(w["W2"]), _ptr(w["b2"]), self.surrogate.h2, _ptr(w["W3"]),
                                                             _ptr(w["b3"]), _ptr(self.sur_s), current_stream_ptr()), "sapt_surrogate_predict")
                self._graph_mode = True
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for k in range(length):
                            self.step(k)
                            _lib.check(self.L.sapt_advance(self.step_dev[0:1].data_ptr(), 1, current_stream_ptr()), "sapt_advance")
                finally:
                    self._graph_mode = False
                self._graphs = {key: g}    # one live graph: surrogate state changes invalidate older ones
                self.step_dev[0] = i       # capture does not execute
            if g is not None:
                g.replay()
            else:
                for k in range(i, j + 1):
                    self.step(k)
            self.rendezvous(j)
            i = j + 1

    # ------------------------------------------------------------------------------------
    def results(self) -> dict:
        """Histories as numpy arrays shaped like the reference's per-chain arrays ([chain, sample...])."""
        self.sync_slots()
        out = {}
        if self.hist is not None:
            hist = self.hist
            if self.cfg.world > 1:
                hist = {k: v for k, v in self.hist.items()}
                if "pos_w" in hist:
                    # rows synchronised at exchanges are present on every rank: keep only hosted-at-write entries
                    hist["pos_w"] = self._pos_w_owned()
                if self.ls is not None:
                    self._ls_sync_history(self.cfg.samples)   # surg is then complete on every rank: not merged again
                hist = {k: (v if (k == "surg" and self.ls is not None) else parallel.merge_history(v, self.pg)) for k, v in hist.items()}
                hist["likeh"][0].fill_(-100.0)
                if "pos_w" in hist:
                    hist["pos_w"][0].fill_(1.0)
            for k in ("lhood_list", "rmse_train", "rmse_test", "acc_train", "acc_test", "accept_list"):
                out[k] = hist[k].t().cpu().numpy()
            out["surg_likeh_list"] = hist["surg"].permute(1, 0, 2).cpu().numpy()
            lk = np.zeros((self.Rg, self.cfg.samples, 2)); lk[:, 0, :] = -100
            lk[:, :, 0] = hist["likeh"].t().cpu().numpy()
            out["likeh_list"] = lk
            out["accepted"] = hist["accepted"].t().cpu().numpy().astype(bool)
            out["is_surrogate_step"] = hist["is_sur"].t().cpu().numpy().astype(bool)
            if "pos_w" in hist:
                out["pos_w"] = hist["pos_w"].permute(1, 0, 2).cpu().numpy().astype(np.float64)
        out["naccept"] = self.naccept.cpu().numpy()
        out["swap_decisions"] = [s.cpu().numpy().astype(bool) for s in self.swap_log]
        out["num_swap"] = int(self.counters[3].item())
        out["total_swap_proposals"] = self.total_swap_proposals
        out["surr_batches"] = [b.cpu().numpy() for b in self.surr_batches]
        out["min_y"] = self.surrogate.min_y; out["max_y"] = self.surrogate.max_y; out["signature"] = self.surrogate.signature
        out["row_of_slot"] = self.row_of_slot.cpu().numpy()
        out["likelihood"] = self.lik_post.cpu().numpy(); out["likelihood_copy"] = self.lik_cur.cpu().numpy()
        out["prior_current"] = self.prior_cur.cpu().numpy()
        if self.ls is not None:
            out["ls_weights"] = self.ls_weights.cpu().numpy(); out["ls_limit"] = self.ls_limit.cpu().numpy()
            out["ls_train_rmse"] = [t.cpu().numpy() for t in self.ls_train_rmse]
        return out

    def _pos_w_owned(self):
        """pos_w history with every entry that this rank did not write itself zeroed (rows that were
        all-reduced at an exchange exist on all ranks and must be counted once in the merge)."""
        owned = self.hist["pos_w"].clone()
        S = owned.shape[0]
        owner = torch.zeros((S, self.Rg), dtype=owned.dtype, device=owned.device)
        segs = self._segments + [(S, None)]
        for (r0, mask), (r1, _) in zip(segs[:-1], segs[1:]):
            owner[r0:r1] = mask[None, :].to(owned.dtype)
        owned.mul_(owner[:, :, None])
        return owned

    def export_streams(self, steps: Optional[int] = None) -> dict:
        """The engine's in-kernel Philox streams as arrays indexed [step, chain(, P)] (and swap_u [sweep, R-1]): exactly
        the numbers the kernels draw for this seed (sapt_philox_fill runs the same device functions). Feeding them to
        another implementation of the chain loop (the oracle's run_pt, or `inject=` of a second engine) gives a
        same-seed run: the posterior-parity tests are built on this."""
        n = self.cfg.samples - 1 if steps is None else int(steps)
        R, P, seed = self.Rg, self.P, self.cfg.seed & 0xFFFFFFFF
        z1 = torch.zeros((n, R, P), dtype=torch.float32, device=self.dev); z2 = torch.zeros_like(z1)
        sc = {k: torch.zeros((n, R), dtype=torch.float32, device=self.dev) for k in ("ku", "lx2", "u")}
        n_sweeps = n // max(1, self.cfg.swap_interval) + 2
        swap_u = torch.zeros((n_sweeps, max(1, R - 1)), dtype=torch.float32, device=self.dev)
        fill, st = self.L.sapt_philox_fill, current_stream_ptr()
        for i in range(n):
            for k in range(R):
                _lib.check(fill(seed, _lib.STREAM_Z1, i, k, P, 1, z1[i, k].data_ptr(), st))
                _lib.check(fill(seed, _lib.STREAM_Z2, i, k, P, 1, z2[i, k].data_ptr(), st))
                for name, kind in (("ku", _lib.STREAM_KU), ("lx2", _lib.STREAM_LX2), ("u", _lib.STREAM_U)):
                    _lib.check(fill(seed, kind, i, k, 1, 0, sc[name][i, k:k + 1].data_ptr(), st))
        if R > 1:
            for s in range(n_sweeps):
                _lib.check(fill(seed, _lib.STREAM_SWAP, s, 0, R - 1, 0, swap_u[s].data_ptr(), st))
        out = dict(z1=z1.cpu().numpy().astype(np.float64), z2=z2.cpu().numpy().astype(np.float64),
                   swap_u=swap_u.cpu().numpy().astype(np.float64))
        out.update({k: v.cpu().numpy().astype(np.float64) for k, v in sc.items()})
        return out

    def current_w(self) -> np.ndarray:
        """[chain, P] parameter vector each chain currently holds (single rank)."""
        th = self.theta[:, : self.P].cpu().numpy().astype(np.float64)
        ros = self.row_of_slot.cpu().numpy()
        return th[ros - self.row0] if self.cfg.world == 1 else th
