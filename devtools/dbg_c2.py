import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
from sapt_b200 import ParallelTempering
g = gu.load("unit_ions"); topo = [int(v) for v in g["topology"]]
for use_sur in (False, True):
    pt = ParallelTempering(use_sur, True, 0.01, False, g["train"], g["test"], topo, 10, 2.0, 300 * 10, 100000, 100, 0.5, "/tmp/x", "/tmp/x", 1, variant="lg", seed=3)
    pt.initialize_chains(0.5)
    eng = pt.engine
    for i in range(299):
        eng.step(i); eng.rendezvous(i)
        if i in (0, 1, 2, 50, 99, 100, 101, 102, 150, 298):
            torch.cuda.synchronize()
            print(use_sur, i, "theta nan", bool(torch.isnan(eng.theta).any()), "lik_cur", eng.lik_cur[:3].cpu().numpy(), "Ltrain", eng.L_train[:3].cpu().numpy(),
                  "n_true/sur/lang", eng.counters[:3].cpu().numpy(), "acc", eng.hist["acc_train"][i + 1, :3].cpu().numpy(), "sur ready", eng.surrogate.ready, eng.surrogate.min_y)
