"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden

Two kinds of fixtures:

1. ``unit_<problem>.npz`` — outputs of the reference's own ``Network.evaluate_proposal``,
   ``ptReplica.likelihood_func``, ``ptReplica.prior_likelihood``, ``Network.langevin_gradient``,
   ``ParallelTempering.swap_procedure`` and ``default_beta_ladder`` on seeded inputs.

2. ``run_<variant>_<tag>.npz`` — full traces of R chains of the reference's ``ptReplica.run``
   driven by the reference's coordinator, with
     * ``multiprocessing`` replaced by thread-based equivalents (same API) so the chains run
       in this process,
     * the module-level ``np.random`` / ``random`` replaced by per-chain recorded streams,
     * Keras replaced by a fake ``load_model`` returning a fixed MLP (the reference's own
       ``surrogate.predict`` de-normalisation code still runs), ``surrogate.train`` a no-op.
   The recorded streams and traces pin ``oracle.sapt_oracle.run_pt`` (and through it the
   device engine) step by step.
"""
from __future__ import annotations

import contextlib
import io
import os
import queue
import random as _pyrandom
import sys
import tempfile
import threading
import types

import numpy as np

from . import ref_loader
from . import sapt_oracle as so

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DATA = os.path.join(ref_loader.REFERENCE_ROOT, "DATA")


# ----------------------------------------------------------------------------------------
# data sets (the reference's own preprocessing, RV:1425-1535)
# ----------------------------------------------------------------------------------------

def load_problem(name: str):
    if name == "cancer":  # RV:1455-1465 (hidden 12 is the commented alternative used by BASELINE C1)
        tr = np.genfromtxt(os.path.join(DATA, "Cancer/ftrain.txt"), delimiter=' ')[:, :-1]
        te = np.genfromtxt(os.path.join(DATA, "Cancer/ftest.txt"), delimiter=' ')[:, :-1]
        return tr, te, [9, 12, 2]
    if name == "ions":  # RV:1467-1474
        tr = np.genfromtxt(os.path.join(DATA, "Ions/Ions/ftrain.csv"), delimiter=',')[:, :-1]
        te = np.genfromtxt(os.path.join(DATA, "Ions/Ions/ftest.csv"), delimiter=',')[:, :-1]
        return tr, te, [34, 50, 2]
    if name == "pendigit":  # RV:1490-1503: z-score features per column
        tr = np.genfromtxt(os.path.join(DATA, "PenDigit/train.csv"), delimiter=',')
        te = np.genfromtxt(os.path.join(DATA, "PenDigit/test.csv"), delimiter=',')
        for k in range(16):
            mt, st = np.mean(tr[:, k]), np.std(tr[:, k])
            tr[:, k] = (tr[:, k] - mt) / st
            mt, st = np.mean(te[:, k]), np.std(te[:, k])
            te[:, k] = (te[:, k] - mt) / st
        return tr, te, [16, 30, 10]
    raise KeyError(name)


# ----------------------------------------------------------------------------------------
# unit fixtures
# ----------------------------------------------------------------------------------------

def make_unit(problem: str, n_w: int, n_rows_langevin: int, seed: int, max_rows=None):
    ref = ref_loader.load_reference("LG")
    tr, te, topo = load_problem(problem)
    if max_rows:
        tr, te = tr[:max_rows], te[:max_rows // 2]
    P = so.num_param(topo)
    rs = np.random.RandomState(seed)
    W = rs.randn(n_w, P)
    W[n_w // 2:] *= 0.3  # a few "trained-looking" smaller vectors too
    temps = np.linspace(1.0, 4.0, n_w)
    net = ref.Network(topo, tr, te, 0.5)
    out = dict(train=tr, test=te, topology=np.array(topo), W=W, temps=temps)
    fx_tr, prob_tr, L_tr, Lt_tr, rm_tr, fx_te, L_te, rm_te, prior, acc_tr = ([] for _ in range(10))
    for k in range(n_w):
        rep = ref_loader.DuckReplica(ref, topo, temps[k])
        fx, prob = net.evaluate_proposal(tr, W[k])
        lt, fx2, rm, l = rep.likelihood_func(net, tr, W[k])
        assert np.array_equal(fx, fx2)
        fx_tr.append(fx); prob_tr.append(prob); L_tr.append(l); Lt_tr.append(lt); rm_tr.append(rm)
        acc_tr.append(rep.accuracy(fx, tr[:, topo[0]]))
        lt, fx3, rm, l = rep.likelihood_func(net, te, W[k])
        fx_te.append(fx3); L_te.append(l); rm_te.append(rm)
        prior.append(rep.prior_likelihood(25, 0, 0, W[k]))
    out.update(fx_train=np.array(fx_tr), prob_train=np.array(prob_tr), L_train=np.array(L_tr),
               Ltemp_train=np.array(Lt_tr), rmse_train=np.array(rm_tr), acc_train=np.array(acc_tr),
               fx_test=np.array(fx_te), L_test=np.array(L_te), rmse_test=np.array(rm_te), prior=np.array(prior))
    # Langevin: sequential SGD epoch on the first n_rows_langevin rows, lr 0.5 (LG:1725) and 0.01
    sub = tr[:n_rows_langevin]
    lg = []
    for lr in (0.5, 0.01):
        net_l = ref.Network(topo, sub, te, lr)
        lg.append(np.array([net_l.langevin_gradient(sub, W[k].copy(), 1) for k in range(n_w)]))
    out.update(langevin_rows=np.array(n_rows_langevin), langevin_lr=np.array([0.5, 0.01]), langevin_out=np.array(lg))
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"unit_{problem}.npz"), **out)
    print(f"unit_{problem}: N={tr.shape[0]}/{te.shape[0]} P={P} L_train[0]={L_tr[0]:.6f}")


def make_swap_and_ladder(seed: int):
    ref = ref_loader.load_reference("RW")
    out = {}
    # temperature ladders (RV:835-930)
    for R, mt in ((8, 2.0), (10, 4.0), (4, 4.0), (16, 6.0)):
        duck = types.SimpleNamespace(num_chains=R, maxtemp=mt, geometric=True, temperatures=[])
        duck.default_beta_ladder = lambda ndim, ntemps, Tmax: ref.ParallelTempering.default_beta_ladder(duck, ndim, ntemps, Tmax)
        ref.ParallelTempering.assign_temperatures(duck)
        out[f"ladder_R{R}_T{mt}"] = np.array(duck.temperatures, dtype=np.float64)
    # the ptemcee step table (RV:854-881): Tmax derived from ntemps for every tabulated dimension and beyond
    duck = types.SimpleNamespace()
    out["ladder_ndim"] = np.arange(1, 111)
    out["ladder_betas_ntemps5"] = np.array([ref.ParallelTempering.default_beta_ladder(duck, int(nd), 5, None) for nd in range(1, 111)])
    out["ladder_ntemps_Tmax50"] = np.array([len(ref.ParallelTempering.default_beta_ladder(duck, int(nd), None, 50.0)) for nd in range(1, 111)])
    # swap sweeps: drive the reference's swap_procedure through queues exactly like RW:1095-1099
    rs = np.random.RandomState(seed)
    P = 5
    cases = []
    for R, spread in ((8, 1.0), (8, 50.0), (16, 0.2), (5, 2000.0)):
        lik = rs.randn(R) * spread - 300
        us = rs.uniform(size=R - 1)
        qs = [queue.Queue() for _ in range(R)]
        for k in range(R):
            qs[k].put(np.concatenate([np.full(P, float(k)), [0.0], [lik[k]], [1.0 + k], [7.0]]))
        duck = types.SimpleNamespace(num_param=P, num_swap=0, total_swap_proposals=0)
        it = iter(us)
        real_uniform = ref.np.random.uniform
        swapped = []
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                ref.np.random.uniform = lambda *a, **k: next(it)
                for k in range(R - 1):
                    p1, p2, s = ref.ParallelTempering.swap_procedure(duck, qs[k], qs[k + 1])
                    qs[k].put(p1); qs[k + 1].put(p2)
                    swapped.append(s)
            finally:
                ref.np.random.uniform = real_uniform
        src = np.array([int(qs[k].get()[0]) for k in range(R)])
        cases.append((lik, us, src, np.array(swapped), duck.num_swap))
    for n, (lik, us, src, sw, ns) in enumerate(cases):
        out[f"swap{n}_lik"] = lik; out[f"swap{n}_u"] = us; out[f"swap{n}_src"] = src
        out[f"swap{n}_swapped"] = sw; out[f"swap{n}_num_swap"] = np.array(ns)
    out["n_swap_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "swap_ladder.npz"), **out)
    print("swap_ladder: cases", len(cases))


# ----------------------------------------------------------------------------------------
# full-run fixtures: threads + recorded streams
# ----------------------------------------------------------------------------------------

class _ThreadProcess(threading.Thread):
    """multiprocessing.Process look-alike (start/join/is_alive/run)."""

    def __init__(self, *a, **k):
        threading.Thread.__init__(self, daemon=True)


class _Q(queue.Queue):
    def close(self):
        pass

    def join_thread(self):
        pass

    def empty(self):  # same meaning
        return queue.Queue.empty(self)


class _JQ(_Q):
    pass


class _AutoResetEvent(threading.Event):
    """Event whose wait() consumes the set(). LG chains never clear the events they wait on
    (LG:808,827), so a faithful barrier needs one wake-up per set()."""

    def wait(self, timeout=None):
        r = threading.Event.wait(self, timeout)
        self.clear()
        return r


def _fake_multiprocessing(auto_reset: bool):
    m = types.ModuleType("multiprocessing")
    m.Process = _ThreadProcess
    m.Queue = _Q
    m.JoinableQueue = _JQ
    m.Event = _AutoResetEvent if auto_reset else threading.Event
    return m


def _load_threaded(variant: str):
    """Load a private copy of the reference script whose `multiprocessing` is thread based."""
    ref_loader._install_stubs()
    import importlib.util
    real_mp = sys.modules.get("multiprocessing")
    sys.modules["multiprocessing"] = _fake_multiprocessing(auto_reset=(variant in ("LG", "LS")))
    try:
        path = os.path.join(ref_loader.REFERENCE_ROOT, ref_loader.VARIANTS[variant])
        spec = importlib.util.spec_from_file_location(f"_sapt_reference_threaded_{variant}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.modules["multiprocessing"] = real_mp
    return mod


_KIND_ID = {"normal": 1, "randn": 2, "uniform": 3, "py": 4}


class _Recorder:
    """Per-thread random streams. Every draw is produced by a seeded RandomState private to
    (thread, kind) and appended to a log so the fixture can replay it."""

    def __init__(self, seed: int):
        self.seed = seed
        self.lock = threading.Lock()
        self.states = {}
        self.log = {}
        self.index_of = {}  # thread -> chain index (main thread = -1)

    def _key(self):
        t = threading.current_thread()
        return self.index_of.get(t, -1)

    def _rs(self, kind):
        k = (self._key(), kind)
        with self.lock:
            if k not in self.states:
                self.states[k] = np.random.RandomState((self.seed * 1000003 + (k[0] + 7) * 101 + _KIND_ID[kind]) % (2 ** 31))
                self.log[k] = []
            return self.states[k], self.log[k]

    # numpy.random API used by the reference -------------------------------------------
    def normal(self, loc=0.0, scale=1.0, size=None):
        rs, log = self._rs("normal")
        z = rs.standard_normal(size)
        log.append(("normal", np.array(z, dtype=np.float64).reshape(-1)))
        return loc + scale * z

    def randn(self, *shape):
        rs, log = self._rs("randn")
        z = rs.standard_normal(shape if shape else None)
        log.append(("randn", np.array(z, dtype=np.float64).reshape(-1)))
        return z

    def uniform(self, low=0.0, high=1.0, size=None):
        rs, log = self._rs("uniform")
        u = rs.uniform(size=size)
        log.append(("uniform_arr" if size is not None else "uniform", np.array(u, dtype=np.float64).reshape(-1)))
        return low + (high - low) * u

    def seed(self, *a):
        pass

    # python random API -------------------------------------------------------------------
    def py_uniform(self, a, b):
        rs, log = self._rs("py")
        u = float(rs.uniform())
        log.append(("py", np.array([u])))
        return a + (b - a) * u


class _NpProxy:
    def __init__(self, rec: _Recorder):
        self.random = types.SimpleNamespace(normal=rec.normal, randn=rec.randn, uniform=rec.uniform, seed=rec.seed)

    def __getattr__(self, name):
        return getattr(np, name)


class _FakeKerasModel:
    def __init__(self, weights):
        self.w = weights

    def compile(self, *a, **k):
        pass

    def predict(self, X, *a, **k):
        return so.surrogate_forward(X, self.w).reshape(-1, 1)


def make_run(variant: str, tag: str, problem: str, R: int, samples_per_chain: int, swap_interval: int,
             surrogate_interval: int, surrogate_prob: float, maxtemp: float, use_langevin: bool,
             learn_rate: float, surrogate_topology: int, seed: int, n_train: int, n_test: int):
    mod = _load_threaded({"rw": "RW", "lg": "LG", "rv": "RV"}[variant])
    tr, te, topo = load_problem(problem)
    tr, te = tr[:n_train].copy(), te[:n_test].copy()
    P = so.num_param(topo)
    rec = _Recorder(seed)
    mod.np = _NpProxy(rec)
    mod.random = types.SimpleNamespace(uniform=rec.py_uniform)
    n_sig = samples_per_chain // max(1, min(swap_interval, surrogate_interval)) + 2
    sur_w = [so.make_surrogate_weights(P, surrogate_topology, seed * 31 + s) for s in range(n_sig)]

    def fake_load_model(path):
        sig = int(float(os.path.basename(path).split("_")[2]))
        return _FakeKerasModel(sur_w[sig - 1])

    mod.load_model = fake_load_model
    mod.surrogate.train = lambda self, model_signature: None

    tmp = tempfile.mkdtemp(prefix="sapt_golden_")
    for sub in ("posterior/pos_w", "posterior/pos_likelihood", "posterior/surg_likelihood", "posterior/accept_list",
                "predictions", "surrogate", "surrogate/learnsurrogate_data"):
        os.makedirs(os.path.join(tmp, sub), exist_ok=True)

    NumSample = samples_per_chain * R
    with contextlib.redirect_stdout(io.StringIO()):
        pt = mod.ParallelTempering(True, use_langevin, learn_rate, False, tr, te, topo, R, maxtemp, NumSample,
                                   swap_interval, surrogate_interval, surrogate_prob, tmp, tmp, surrogate_topology)
        pt.initialize_chains(0.25)
    for k, ch in enumerate(pt.chains):
        rec.index_of[ch] = k
    w0 = rec.log[(-1, "randn")][-1][1].copy()  # initialize_chains: w = np.random.randn(num_param)
    temps = np.array(pt.temperatures, dtype=np.float64)

    # capture per-chain history arrays by intercepting np.savetxt inside the module
    saved = {}

    def savetxt(fname, arr, *a, **k):
        saved[os.path.relpath(str(fname), tmp) if not hasattr(fname, "write") else "<file>"] = np.array(arr, dtype=np.float64)
        np.savetxt(fname, arr, *a, **k)

    mod.np.savetxt = savetxt
    # record every swap decision by wrapping the reference's swap_procedure
    swap_log = []
    orig_swap = mod.ParallelTempering.swap_procedure

    def logged_swap(self, q1, q2):
        res = orig_swap(self, q1, q2)
        swap_log.append(res[2] if (res is not None and len(res) == 3) else None)
        return res

    mod.ParallelTempering.swap_procedure = logged_swap
    trainer_batches = []
    orig_trainer = mod.ParallelTempering.surrogate_trainer

    def logged_trainer(self, params):
        trainer_batches.append(np.array(params, dtype=np.float64).copy())
        return orig_trainer(self, params)

    mod.ParallelTempering.surrogate_trainer = logged_trainer
    ref_show_results = mod.ParallelTempering.show_results    # run below on the files the chains leave behind
    pt.show_results = lambda: tuple([None] * 12)

    with contextlib.redirect_stdout(io.StringIO()):
        if variant == "lg":
            _run_lg_barrier(mod, pt)
        else:
            pt.run_chains()

    # the reference's own aggregation (RV:1231-1392 and the forks' copies): burn-in slicing, reshape order, rmse_surr —
    # fed by the per-chain files the chains have just written (with the reference's own text formats)
    show = {}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            shown = ref_show_results(pt)
    except ValueError:
        shown = None   # a run without a single surrogate step: the reference's plot code dies on np.amax([]) (RW:1280)
    if shown is not None:
        show_names = ("posterior", "fx_train", "fx_test", "rmse_train", "rmse_test", "acc_train", "acc_test", "likelihood_vec",
                      "accept_list", "rmse_surr", "surr_list", "accept")
        show = {f"show_{n}": np.array(v, dtype=np.float64) for n, v in zip(show_names, shown) if not n.startswith("fx_")}
        show["show_fx_shapes"] = np.array([np.shape(shown[1]), np.shape(shown[2])])
        show["show_burn_in"] = np.array(pt.burn_in)
        for fn in ("pos_param.txt", "likelihood.txt", "acc_train.txt", "acc_test.txt"):
            show["showfile_" + fn.replace(".txt", "")] = np.loadtxt(os.path.join(tmp, fn))

    # ---- assemble streams [step, chain, ...] from the per-thread logs ----
    S = samples_per_chain
    z1 = np.zeros((S - 1, R, P)); z2 = np.zeros((S - 1, R, P)); has_z2 = np.zeros((S - 1, R), dtype=bool)
    lx1 = np.zeros((S - 1, R)); lx2 = np.zeros((S - 1, R)); has_lx2 = np.zeros((S - 1, R), dtype=bool)
    ku = np.zeros((S - 1, R)); u = np.zeros((S - 1, R))
    for k in range(R):
        normals = [v for (t, v) in rec.log[(k, "normal")]]
        unis = [v for (t, v) in rec.log[(k, "uniform")]] if (k, "uniform") in rec.log else []
        pys = [v for (t, v) in rec.log[(k, "py")]]
        assert len(pys) == 2 * (S - 1), (len(pys), S)
        # walk the step structure: each step starts with uniform_arr (lx); normals: 1 or 2 per step;
        # second uniform_arr only on LG true steps. Use the reference's own traces to decide.
        ku[:, k] = np.array(pys[0::2]).reshape(-1)
        u[:, k] = np.array(pys[1::2]).reshape(-1)
        rec.log[(k, "_normals")] = normals
        rec.log[(k, "_unis")] = unis
    meta = dict(z1=z1, z2=z2, has_z2=has_z2, lx1=lx1, lx2=lx2, has_lx2=has_lx2, ku=ku, u=u)
    _assemble_np_streams(rec, meta, variant, R, S, P, surrogate_prob, swap_interval, surrogate_interval, use_langevin)

    main_u = [v for (t, v) in rec.log.get((-1, "uniform"), [])]
    swap_u = np.array(main_u, dtype=np.float64).reshape(-1, R - 1) if main_u else np.zeros((0, R - 1))

    mm = os.path.join(tmp, "surrogate", "minmax.txt")  # what the workers de-normalise with (RV:707-710)
    minmax = np.loadtxt(mm) if os.path.exists(mm) else np.array([0.0, 0.0])
    traces = {}
    for k in range(R):
        tstr = str(pt.temperatures[k])
        traces.setdefault("pos_w", []).append(saved[f"posterior/pos_w/chain_{tstr}.txt"])
        traces.setdefault("rmse_test", []).append(saved[f"predictions/rmse_test_chain_{tstr}.txt"])
        traces.setdefault("rmse_train", []).append(saved[f"predictions/rmse_train_chain_{tstr}.txt"])
        traces.setdefault("acc_test", []).append(saved[f"predictions/acc_test_chain_{tstr}.txt"])
        traces.setdefault("acc_train", []).append(saved[f"predictions/acc_train_chain_{tstr}.txt"])
        traces.setdefault("surg_likeh_list", []).append(saved[f"posterior/surg_likelihood/chain_{tstr}.txt"])
        traces.setdefault("likeh_list", []).append(saved[f"posterior/pos_likelihood/chain_{tstr}.txt"])
        traces.setdefault("accept_list", []).append(saved[f"posterior/accept_list/chain_{tstr}.txt"])
        traces.setdefault("accept_ratio", []).append(saved[f"posterior/accept_list/chain_{tstr}_accept.txt"])
    out = {f"trace_{k}": np.array(v) for k, v in traces.items()}
    out.update(train=tr, test=te, topology=np.array(topo), temperatures=temps, w0=w0, samples=np.array(S),
               swap_interval=np.array(swap_interval), surrogate_interval=np.array(surrogate_interval),
               surrogate_prob=np.array(surrogate_prob), use_langevin=np.array(use_langevin),
               learn_rate=np.array(learn_rate), surrogate_topology=np.array(surrogate_topology),
               swap_u=swap_u, swap_log=np.array([(-1 if s is None else int(s)) for s in swap_log]),
               num_swap=np.array(pt.num_swap), total_swap_proposals=np.array(pt.total_swap_proposals),
               n_trainer_calls=np.array(len(trainer_batches)), min_y=np.array(minmax[0]),
               max_y=np.array(minmax[1]), n_signatures=np.array(min(n_sig, len(trainer_batches))), **meta, **show)
    for s, b in enumerate(trainer_batches):
        out[f"trainer_batch_{s}"] = b
    for s in range(min(n_sig, len(trainer_batches))):  # only the signatures the run reached
        for name, arr in sur_w[s].items():
            out[f"sur{s}_{name}"] = arr
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"run_{variant}_{tag}.npz"), **out)
    acc = [float(a[0]) if np.ndim(a) else float(a) for a in traces["accept_ratio"]]
    print(f"run_{variant}_{tag}: R={R} S={S} swaps={pt.num_swap}/{pt.total_swap_proposals} "
          f"trainer_calls={len(trainer_batches)} accept%={np.round(acc, 1)}")


def _assemble_np_streams(rec, meta, variant, R, S, P, surrogate_prob, swap_interval, surrogate_interval, use_langevin):
    """Map the flat per-chain numpy draw logs onto [step, chain]. The per-step draw pattern is
    a deterministic function of (ku, lx2) and the variant (see LG:530-651, RW:492-495):
      every step : uniform_arr (lx, unused), normal(P) = z1
      LG surrogate step : + normal(P) = z2 (redraw)
      LG true step : + uniform_arr = lx2 ; + normal(P) = z2 iff use_langevin and lx2 < 0.5
    """
    gate = swap_interval if variant == "rw" else surrogate_interval
    for k in range(R):
        log = rec.log[(k, "normal")]
        ulog = [v for (t, v) in rec.log.get((k, "uniform"), []) if t == "uniform_arr"]
        ni = 0  # index into normal log (skip nothing: the pre-loop draw is randn, logged separately)
        ui = 0
        for i in range(S - 1):
            meta["lx1"][i, k] = ulog[ui][0]; ui += 1
            meta["z1"][i, k] = log[ni][1]; ni += 1
            is_sur = (meta["ku"][i, k] < surrogate_prob) and (i >= gate + 1)
            if variant == "lg":
                if is_sur:
                    meta["z2"][i, k] = log[ni][1]; ni += 1; meta["has_z2"][i, k] = True
                else:
                    meta["lx2"][i, k] = ulog[ui][0]; ui += 1; meta["has_lx2"][i, k] = True
                    if use_langevin and meta["lx2"][i, k] < 0.5:
                        meta["z2"][i, k] = log[ni][1]; ni += 1; meta["has_z2"][i, k] = True
        assert ni == len(log), (variant, k, ni, len(log))
        assert ui == len(ulog), (variant, k, ui, len(ulog))


def _run_lg_barrier(mod, pt):
    """Deterministic coordinator for LG chains. LG's own run_chains is racy (it waits on a stale
    index and swap_procedure silently skips pairs whose queue is momentarily empty,
    LG:1083,1108-1110,1239-1240), so — as SURVEY.md §3.5 prescribes — the chains are driven with
    RW's barrier semantics: wait for ALL chains to post, sweep k = 0..R-2 with the reference's
    own LG swap_procedure, release all; same for the surrogate rendezvous."""
    R = pt.num_chains
    for ch in pt.chains:
        ch.start()
    S = pt.NumSamples
    swap_steps = [i for i in range(S - 1) if i % pt.swap_interval == 0]
    sur_steps = [i for i in range(S - 1) if (i % pt.surrogate_interval == 0 and i != 0)]
    for i in range(S - 1):
        if i in swap_steps:
            for k in range(R):
                pt.wait_chain[k].wait()
            for k in range(R - 1):
                res = pt.swap_procedure(pt.parameter_queue[k], pt.parameter_queue[k + 1])
                assert res is not None
                p1, p2 = res
                pt.parameter_queue[k].put(p1); pt.parameter_queue[k + 1].put(p2)
            for k in range(R):
                pt.event[k].set()
        if i in sur_steps:
            for k in range(R):
                pt.surrogate_start_events[k].wait()
            params = [pt.surrogate_parameterqueues[k].get() for k in range(R)]
            pt.surrogate_trainer(np.concatenate(params, axis=0))
            for k in range(R):
                pt.surrogate_resume_events[k].set()
    for ch in pt.chains:
        ch.join()


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if not ref_loader.reference_available():
        raise SystemExit("reference not available: fixtures can only be generated in the build container")
    make_unit("cancer", n_w=6, n_rows_langevin=120, seed=11)
    make_unit("ions", n_w=4, n_rows_langevin=60, seed=12)
    make_unit("pendigit", n_w=4, n_rows_langevin=40, seed=13, max_rows=600)
    make_swap_and_ladder(seed=21)
    make_run("rw", "cancer", "cancer", R=4, samples_per_chain=48, swap_interval=8, surrogate_interval=8,
             surrogate_prob=0.5, maxtemp=4.0, use_langevin=False, learn_rate=0.5, surrogate_topology=1,
             seed=31, n_train=96, n_test=40)
    make_run("rw", "nosurr", "cancer", R=5, samples_per_chain=40, swap_interval=5, surrogate_interval=5,
             surrogate_prob=0.0, maxtemp=2.0, use_langevin=False, learn_rate=0.5, surrogate_topology=1,
             seed=32, n_train=80, n_test=30)
    make_run("lg", "cancer", "cancer", R=4, samples_per_chain=40, swap_interval=6, surrogate_interval=8,
             surrogate_prob=0.4, maxtemp=4.0, use_langevin=True, learn_rate=0.5, surrogate_topology=2,
             seed=33, n_train=64, n_test=32)
    make_run("rv", "cancer", "cancer", R=4, samples_per_chain=44, swap_interval=5, surrogate_interval=7,
             surrogate_prob=0.5, maxtemp=2.0, use_langevin=False, learn_rate=0.5, surrogate_topology=1,
             seed=35, n_train=90, n_test=36)
    make_run("lg", "ions", "ions", R=3, samples_per_chain=24, swap_interval=100000, surrogate_interval=6,
             surrogate_prob=0.5, maxtemp=2.0, use_langevin=True, learn_rate=0.01, surrogate_topology=3,
             seed=34, n_train=48, n_test=24)
    make_run_local("cancer", "cancer", R=4, samples_per_chain=60, swap_interval=7, surrogate_interval=12,
                   surrogate_prob=0.4, maxtemp=3.0, seed=36, n_train=80, n_test=30)



# ----------------------------------------------------------------------------------------
# N4: local-surrogate fork (LS). Same thread/stream harness; the fork's own coordinator is racy in the same
# way as LG's (LS:1108 waits on a stale index), so the chains are driven by a barrier that sweeps the fork's
# own swap_procedure, as for LG. The per-chain 4-10-1 networks train inside the reference code (numpy only).
# ----------------------------------------------------------------------------------------
def _run_ls_barrier(pt):
    R = pt.num_chains
    for ch in pt.chains:
        ch.start()
    S = pt.NumSamples
    for i in range(S - 1):
        if i % pt.swap_interval == 0:
            for k in range(R):
                pt.wait_chain[k].wait()
            for k in range(R - 1):
                res = pt.swap_procedure(pt.parameter_queue[k], pt.parameter_queue[k + 1])
                assert res is not None
                p1, p2 = res
                pt.parameter_queue[k].put(p1); pt.parameter_queue[k + 1].put(p2)
            for k in range(R):
                pt.event[k].set()
    for ch in pt.chains:
        ch.join()


def make_run_local(tag: str, problem: str, R: int, samples_per_chain: int, swap_interval: int, surrogate_interval: int,
                   surrogate_prob: float, maxtemp: float, seed: int, n_train: int, n_test: int):
    mod = _load_threaded("LS")
    tr, te, topo = load_problem(problem)
    tr, te = tr[:n_train].copy(), te[:n_test].copy()
    P = so.num_param(topo)
    rec = _Recorder(seed)
    mod.np = _NpProxy(rec)
    mod.random = types.SimpleNamespace(uniform=rec.py_uniform)
    tmp = tempfile.mkdtemp(prefix="sapt_golden_")
    for sub in ("posterior/pos_w", "posterior/pos_likelihood", "posterior/surg_likelihood", "posterior/accept_list",
                "predictions", "surrogate", "surrogate/learnsurrogate_data"):
        os.makedirs(os.path.join(tmp, sub), exist_ok=True)
    train_log = []   # (chain, init flag, windows/limit, rmse, weights after)
    orig_train = mod.surrogate.train

    def logged_train(self, model_signature, temperature, X, Y):
        rmse = orig_train(self, model_signature, temperature, X, Y)
        k = rec.index_of.get(threading.current_thread(), -1)
        f = self.fnn
        train_log.append((k, bool(model_signature), np.column_stack((X, Y)).copy(), float(rmse),
                          np.concatenate([f.W1.ravel(), f.W2.ravel(), f.B1.ravel(), f.B2.ravel()]).copy()))
        return rmse

    mod.surrogate.train = logged_train
    NumSample = samples_per_chain * R
    with contextlib.redirect_stdout(io.StringIO()):
        pt = mod.ParallelTempering(True, False, 0.5, False, tr, te, topo, R, maxtemp, NumSample, swap_interval,
                                   surrogate_interval, surrogate_prob, tmp, tmp)
        pt.initialize_chains(0.25)
    for k, ch in enumerate(pt.chains):
        rec.index_of[ch] = k
    w0 = rec.log[(-1, "randn")][-1][1].copy()
    temps = np.array(pt.temperatures, dtype=np.float64)
    saved = {}

    def savetxt(fname, arr, *a, **k):
        if not hasattr(fname, "write"):
            saved[os.path.relpath(str(fname), tmp)] = np.array(arr, dtype=np.float64)
        np.savetxt(fname, arr, *a, **k)

    mod.np.savetxt = savetxt
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _run_ls_barrier(pt)
    S = samples_per_chain
    z1 = np.zeros((S - 1, R, P)); ku = np.zeros((S - 1, R)); u = np.zeros((S - 1, R))
    net_init = np.zeros((R, 61))
    for k in range(R):
        normals = [v for (t, v) in rec.log[(k, "normal")]]
        pys = [v for (t, v) in rec.log[(k, "py")]]
        assert len(normals) == S - 1 and len(pys) == 2 * (S - 1), (len(normals), len(pys))
        z1[:, k] = np.array(normals)
        ku[:, k] = np.array(pys[0::2]).reshape(-1); u[:, k] = np.array(pys[1::2]).reshape(-1)
        # randn log of a chain: w_proposal (LS:473), the classifier's create_network (4 draws, overwritten by
        # decode), then the local network's create_network at the first training: W1, B1, W2, B2
        rl = [v for (t, v) in rec.log[(k, "randn")]]
        assert [len(v) for v in rl[-4:]] == [40, 10, 10, 1], [len(v) for v in rl]
        net_init[k] = np.concatenate(rl[-4:])
    main_u = [v for (t, v) in rec.log.get((-1, "uniform"), [])]
    swap_u = np.array(main_u, dtype=np.float64).reshape(-1, R - 1)
    traces = {}
    for k in range(R):
        tstr = str(pt.temperatures[k])
        traces.setdefault("rmse_test", []).append(saved[f"predictions/rmse_test_chain_{tstr}.txt"])
        traces.setdefault("rmse_train", []).append(saved[f"predictions/rmse_train_chain_{tstr}.txt"])
        traces.setdefault("acc_test", []).append(saved[f"predictions/acc_test_chain_{tstr}.txt"])
        traces.setdefault("acc_train", []).append(saved[f"predictions/acc_train_chain_{tstr}.txt"])
        traces.setdefault("surg_likeh_list", []).append(saved[f"posterior/surg_likelihood/chain_{tstr}.txt"])
        traces.setdefault("likeh_list", []).append(saved[f"posterior/pos_likelihood/chain_{tstr}.txt"])
        traces.setdefault("accept_list", []).append(saved[f"posterior/accept_list/chain_{tstr}.txt"])
        traces.setdefault("accept_ratio", []).append(saved[f"posterior/accept_list/chain_{tstr}_accept.txt"])
    out = {f"trace_{k}": np.array(v) for k, v in traces.items()}
    # trainer log, ordered (call number, chain)
    n_calls = len(train_log) // R
    tl = sorted(range(len(train_log)), key=lambda j: (sum(1 for q in train_log[:j] if q[0] == train_log[j][0]), train_log[j][0]))
    for pos, j in enumerate(tl):
        k, init, data, rmse, wts = train_log[j]
        c = pos // R
        assert k == pos % R
        out[f"ls_train{c}_data_{k}"] = data
        out[f"ls_train{c}_weights_{k}"] = wts
        out[f"ls_train{c}_rmse_{k}"] = np.array(rmse)
        out[f"ls_train{c}_init_{k}"] = np.array(init)
    out.update(train=tr, test=te, topology=np.array(topo), temperatures=temps, w0=w0, samples=np.array(S),
               swap_interval=np.array(swap_interval), surrogate_interval=np.array(surrogate_interval),
               surrogate_prob=np.array(surrogate_prob), swap_u=swap_u, num_swap=np.array(pt.num_swap),
               total_swap_proposals=np.array(pt.total_swap_proposals), n_train_calls=np.array(n_calls),
               z1=z1, ku=ku, u=u, net_init=net_init)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"run_ls_{tag}.npz"), **out)
    acc = [float(a[0]) if np.ndim(a) else float(a) for a in traces["accept_ratio"]]
    sur_steps = int(np.sum(~np.isnan(np.array(traces["surg_likeh_list"])[:, 1:, 1])))
    print(f"run_ls_{tag}: R={R} S={S} swaps={pt.num_swap}/{pt.total_swap_proposals} train_calls={n_calls} "
          f"surrogate_steps={sur_steps} accept%={np.round(acc, 1)}")


if __name__ == "__main__":
    main()
