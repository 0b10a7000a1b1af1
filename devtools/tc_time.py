"""Times the tcgen05 likelihood path at the C3 shape (run on the GPU box). SAPT_LIB selects the library."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.problem import Problem

rs = np.random.RandomState(0)
n, D, H, O, R = 1_000_000, 16, 30, 10, 1024
x = rs.standard_normal((n, D)).astype(np.float32); y = rs.randint(0, O, n).astype(np.int32)
P = D * H + H * O + H + O
th = torch.as_tensor(rs.standard_normal((R, P)).astype(np.float32)).cuda()
prob = Problem.from_f32([D, H, O], x, y)
out = prob.loglik(th, 0, path=_lib.PATH_TC); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = prob.loglik(th, 0, path=_lib.PATH_TC)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 3)
print(os.environ.get("SAPT_LIB", "default"), "tc ms %.3f" % best, "evals/s %.3e" % (R * n / best * 1e3), "L[0] %.6f" % out[0][0].item(), flush=True)
