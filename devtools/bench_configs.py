"""Secondary measurements for the other BASELINE.json configs (C1, C2, C4, C5) — parity is covered by tests/,
this only reports throughput. Writes gpurun_out/configs.json (copied to profiles/ by hand)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
from oracle import sapt_oracle as so   # only for the fixed random surrogate weights

def synth(D, O, n, seed):
    rs = np.random.RandomState(seed)
    return rs.standard_normal((n, D)).astype(np.float32), rs.randint(0, O, n).astype(np.int32)

def timed_steps(eng, warm, steps):
    for i in range(warm):
        eng.step(i); eng.rendezvous(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0 = eng.evals
    e0.record()
    for i in range(warm, warm + steps):
        eng.step(i); eng.rendezvous(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return ms / steps, (eng.evals - ev0) / (ms * 1e-3)

out = {}
which = sys.argv[1:] or ["c4", "c5"]
if "c4" in which:
    D, H, O, N, R = 54, 64, 7, 581_012, 256
    x, y = synth(D, O, N, 1); xt, yt = synth(D, O, 65_536, 2)
    prob = Problem.from_f32([D, H, O], x, y, xt, yt)
    cfg = EngineConfig(variant="lg", samples=64, swap_interval=100000, surrogate_interval=100000, surrogate_prob=0.0, use_surrogate=False,
                       use_langevin=True, learn_rate=0.5, langevin_mode="batch", langevin_batch_scale=1.0 / N, seed=4, record_pos=False)
    eng = PTEngine(prob, assign_temperatures(R, 4.0), cfg)
    eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
    ms, eps = timed_steps(eng, 2, 4)
    def time_op(fn, n=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    th = eng.theta if hasattr(eng, "theta") else torch.randn(R, prob.P, device="cuda") * 0.3
    ms_ll = time_op(lambda: prob.loglik(th, 0))
    ms_lg = time_op(lambda: prob.langevin(th, 0.5, mode=1, batch_scale=1.0 / N))
    out["C4 covtype-shape 54-64-7, 581012 rows, 256 chains, Langevin(batch, mean) p=0.5"] = dict(
        ms_per_step=ms, loglik_evals_per_s=eps, samples_per_s=R / (ms * 1e-3), path="auto (tcgen05)",
        loglik_train_256_chains_ms=ms_ll, langevin_batch_256_chains_ms=ms_lg)
    print(json.dumps(out), flush=True)
    del eng, prob
    torch.cuda.empty_cache()
if "c5" in which:
    D, H, O, N = 6, 25, 18, 16_833
    x, y = synth(D, O, N, 5); xt, yt = synth(D, O, 11_223, 6)
    for R in (64, 4096, 65536):
        prob = Problem.from_f32([D, H, O], x, y, xt, yt)
        w = so.make_surrogate_weights(prob.P, 3, 7)
        cfg = EngineConfig(variant="rw", samples=40, swap_interval=4, surrogate_interval=4, surrogate_prob=0.5, use_surrogate=True,
                           surrogate_topology=3, seed=4, record_pos=False, train_cap=4)
        eng = PTEngine(prob, assign_temperatures(R, 4.0), cfg, surrogate_trainer=lambda batch, sig: (w if sig == 1 else None)); eng.surr_batches = type("Drop", (list,), {"append": lambda self, x: None})()
        eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
        ms, eps = timed_steps(eng, 10, 8)
        out[f"C5 chess-shape 6-25-18, 16833+11223 rows, {R} chains, surrogate_prob 0.5 (fixed MLP P-200-50-1)"] = dict(
            ms_per_step=ms, loglik_evals_per_s=eps, samples_per_s=R / (ms * 1e-3), path="auto (tcgen05)")
        print(json.dumps(out), flush=True)
        del eng, prob
        torch.cuda.empty_cache()
if "c2" in which:
    # C2: Ionosphere 34-50-2, 245 train / 109 test rows, 10 replicas, Langevin-gradient PT (reference-exact per-row SGD)
    # with the surrogate interval enabled and the PyTorch surrogate trainer in the loop, through the class surface
    from sapt_b200.parallel_tempering import ParallelTempering
    D, H, O, R, S = 34, 50, 2, 10, 400
    rs = np.random.RandomState(7)
    mk = lambda n: np.concatenate([rs.standard_normal((n, D)), rs.randint(0, O, n)[:, None].astype(np.float64)], axis=1)
    tr, te = mk(245), mk(109)
    # one short untimed run first: the first torch matmul of a process loads cuBLAS (~3 s), which is not the workload
    warm = ParallelTempering(True, True, 0.01, False, tr, te, [D, H, O], R, 4.0, R * 120, 100000, 50, 0.5, "", "", 1, variant="lg", seed=3)
    warm.initialize_chains(0.25); warm.run_chains(); torch.cuda.synchronize()
    for use_sur in (False, True):
        pt = ParallelTempering(use_sur, True, 0.01, False, tr, te, [D, H, O], R, 4.0, R * S, 100000, 50, 0.5, "", "", 1, variant="lg", seed=3)
        pt.initialize_chains(0.25)
        torch.cuda.synchronize(); t0 = time.time()
        pt.run_chains()
        torch.cuda.synchronize(); dt = time.time() - t0
        out[f"C2 ionosphere-shape 34-50-2, 245+109 rows, 10 chains x {S} samples, Langevin (exact SGD) p=0.5, surrogate {'on (interval 50, prob 0.5, torch trainer)' if use_sur else 'off'}"] = dict(
            wall_s=dt, steps_per_s=(S - 1) / dt, samples_per_s=R * (S - 1) / dt)
        print(json.dumps(out), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
