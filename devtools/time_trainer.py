"""Times the surrogate trainer (N1) on a C2-sized interval batch, graph vs eager (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200.surrogate import SurrogateTrainer
rs = np.random.RandomState(1)
P, n = 1852, 10 * 51
X = rs.standard_normal((n, P)) * 0.3
Y = -300.0 - 40.0 * np.tanh(X[:, 0]) - 10.0 * X[:, 2] ** 2
batch = torch.as_tensor(np.concatenate([X, Y[:, None]], axis=1)).cuda()
for graph in (True, False):
    tr = SurrogateTrainer(P, 1, "cuda", seed=9, graph=graph)
    for sig in (1, 2, 3):
        torch.cuda.synchronize(); t0 = time.time()
        tr(batch, sig)
        torch.cuda.synchronize()
        print("graph" if graph else "eager", "signature", sig, "%.1f ms" % ((time.time() - t0) * 1e3), "mse %.3e" % tr.train_mse[-1], flush=True)
