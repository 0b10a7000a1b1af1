"""C1 / C2 end-to-end runs through the reference-facing ParallelTempering class (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
from sapt_b200 import ParallelTempering

def run(name, problem, chains, samples, use_sur, use_lg, sprob, swap_interval, sur_interval, maxtemp, lr, stopo, variant):
    g = gu.load(f"unit_{problem}")
    topo = [int(v) for v in g["topology"]]
    t0 = time.time()
    pt = ParallelTempering(use_sur, use_lg, lr, False, g["train"], g["test"], topo, chains, maxtemp, samples * chains, swap_interval,
                           sur_interval, sprob, "/tmp/x", "/tmp/x", stopo, variant=variant, seed=3)
    pt.initialize_chains(0.5)
    t1 = time.time()
    out = pt.run_chains()
    t2 = time.time()
    (pos_w, fx_tr, fx_te, rmse_tr, rmse_te, acc_tr, acc_te, accept_list, swap_perc, lik, rmse_surr, surr_list, accept) = out
    f = lambda a: a[np.isfinite(a)]
    print(f"{name}: init {t1-t0:.2f}s run {t2-t1:.2f}s  steps/s {samples/(t2-t1):.0f}  chain-samples/s {samples*chains/(t2-t1):.0f}")
    print(f"   acc_train {f(acc_tr).mean():.2f} +- {f(acc_tr).std():.2f} max {f(acc_tr).max():.2f} | acc_test {f(acc_te).mean():.2f} +- {f(acc_te).std():.2f} max {f(acc_te).max():.2f}"
          f" | accept {accept:.1f}% swap {swap_perc:.1f}% rmse_surr {rmse_surr}")

run("C1 cancer 9-12-2 RW 8 chains x 5000 (no surrogate)", "cancer", 8, 5000, False, False, 0.0, 50, 50, 4.0, 0.5, 1, "rw")
run("C1 cancer RW + surrogate 0.5", "cancer", 8, 2500, True, False, 0.5, 100, 100, 4.0, 0.5, 1, "rw")
run("C2 ions 34-50-2 LG langevin + surrogate, 10 chains x 1000", "ions", 10, 1000, True, True, 0.5, 100000, 100, 2.0, 0.01, 1, "lg")
