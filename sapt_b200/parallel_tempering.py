"""`ParallelTempering` — the reference's coordinator surface (_revamp.py:782-1398) over `PTEngine`.

Constructor, `initialize_chains(burn_in)` and `run_chains()` keep the reference's names,
positional order and return tuple. What changes underneath: no processes, queues or events; the
R chains are rows of one device-resident state and every step is a few kernel launches.
Extra keyword arguments (all optional) select the reference fork to mimic and the seed. Calling it with the
local-surrogate fork's 15 positional arguments (no surrogate_topology, LS:894) selects variant "ls"
(that fork comments out its pos_w files, LS:846; the engine keeps pos_w unless record_pos=False)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import EngineConfig, PTEngine
from .ladder import assign_temperatures, default_beta_ladder
from .problem import Problem
from .results import aggregate
from .surrogate import SurrogateTrainer


class ParallelTempering:
    def __init__(self, use_surrogate, use_langevin_gradients, learn_rate, save_surrogate_data, traindata, testdata,
                 topology, num_chains, maxtemp, NumSample, swap_interval, surrogate_interval, surrogate_prob, path,
                 path_db, surrogate_topology=None, *, variant=None, seed=1, langevin_mode="sgd_exact", langevin_batch_scale=1.0,
                 kernel_path=_lib.PATH_AUTO, record_pos=True, process_group=None):
        self.traindata = traindata
        self.testdata = testdata
        self.topology = [int(v) for v in topology]
        d, h, o = self.topology
        self.num_param = d * h + h * o + h + o
        self.swap_interval = int(swap_interval)
        self.path = path
        self.path_db = path_db
        self.maxtemp = maxtemp
        self.num_swap = 0
        self.total_swap_proposals = 0
        self.num_chains = int(num_chains)
        self.chains = []
        self.temperatures = []
        self.NumSamples = int(NumSample / self.num_chains)          # _revamp.py:800
        self.sub_sample_size = max(1, int(0.05 * self.NumSamples))
        self.surrogate_interval = int(surrogate_interval)
        self.surrogate_prob = surrogate_prob
        self.geometric = True
        self.minlim_param = 0.0
        self.maxlim_param = 0.0
        self.minY = np.zeros((1, 1))
        self.maxY = np.ones((1, 1))
        self.model_signature = 0.0
        self.use_surrogate = bool(use_surrogate)
        self.surrogate_topology = surrogate_topology
        self.save_surrogate_data = save_surrogate_data
        self.use_langevin_gradients = bool(use_langevin_gradients)
        self.learn_rate = learn_rate
        # the 15-argument constructor (no surrogate_topology) is the local-surrogate fork's (LS:894): per-chain networks
        if surrogate_topology is None and variant is None:
            variant = "ls"
        if surrogate_topology is None:
            self.surrogate_topology = surrogate_topology = 1
        self.variant = variant or ("lg" if self.use_langevin_gradients else "rw")
        self.seed = int(seed)
        self._langevin_mode = langevin_mode
        self._langevin_batch_scale = langevin_batch_scale
        self._kernel_path = kernel_path
        self._record_pos = record_pos
        self._pg = process_group
        self.engine = None

    def default_beta_ladder(self, ndim, ntemps, Tmax):
        return default_beta_ladder(ndim, ntemps, Tmax)

    def assign_temperatures(self):
        self.temperatures = list(assign_temperatures(self.num_chains, self.maxtemp, self.geometric))

    def initialize_chains(self, burn_in, w=None):
        """_revamp.py:933-943: temperatures, prior limits, ONE random w shared by all chains."""
        self.burn_in = burn_in
        self.assign_temperatures()
        self.minlim_param = np.repeat([-100], self.num_param)
        self.maxlim_param = np.repeat([100], self.num_param)
        if w is None:
            w = np.random.RandomState(self.seed).randn(self.num_param)
        self.w_init = np.asarray(w, dtype=np.float64)
        rank, world = 0, 1
        if self._pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            import torch.distributed as dist
            rank, world = dist.get_rank(self._pg), dist.get_world_size(self._pg)
        self.problem = Problem(self.topology, self.traindata, self.testdata, path=self._kernel_path)
        cfg = EngineConfig(variant=self.variant, samples=self.NumSamples, swap_interval=self.swap_interval,
                           surrogate_interval=self.surrogate_interval, surrogate_prob=float(self.surrogate_prob),
                           use_surrogate=self.use_surrogate, use_langevin=self.use_langevin_gradients,
                           learn_rate=float(self.learn_rate), langevin_mode=self._langevin_mode,
                           langevin_batch_scale=self._langevin_batch_scale, surrogate_topology=int(self.surrogate_topology),
                           seed=self.seed, record_history=True, record_pos=self._record_pos, path=self._kernel_path,
                           rank=rank, world=world)
        trainer = None
        if self.use_surrogate and self.variant != "ls":   # 'ls' trains its per-chain networks on the device
            trainer = SurrogateTrainer(self.num_param, int(self.surrogate_topology), self.problem.device, seed=self.seed)
        self._trainer = trainer
        self.engine = PTEngine(self.problem, self.temperatures, cfg, surrogate_trainer=trainer, process_group=self._pg)
        self.engine.initialize(self.w_init)

    def run_chains(self):
        """-> (pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, accept_list, swap_perc,
        likelihood_vec, rmse_surr, surr_list, accept) with the shapes of _revamp.py:1227,1299-1313."""
        eng = self.engine
        eng.run()
        torch.cuda.synchronize()
        res = eng.results()
        self.results = res
        ntr, nte = self.traindata.shape[0], (self.testdata.shape[0] if self.testdata is not None else 0)
        (posterior, fx_train_all, fx_test_all, rmse_train, rmse_test, acc_train, acc_test, likelihood_vec_t, accept_list, rmse_surr,
         surr_list, accept) = aggregate(res, self.NumSamples, self.num_chains, self.num_param, self.burn_in, self.use_surrogate, ntr, nte)
        self.num_swap = res["num_swap"]
        self.total_swap_proposals = res["total_swap_proposals"]
        swap_perc = self.num_swap * 100 / self.total_swap_proposals if self.total_swap_proposals else 0
        return (posterior, fx_train_all, fx_test_all, rmse_train, rmse_test, acc_train, acc_test, accept_list, swap_perc,
                likelihood_vec_t, rmse_surr, surr_list, accept)
