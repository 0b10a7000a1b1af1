"""Times the tcgen05 likelihood kernel at the C3 shape (dev tool; SAPT_LIB selects an A/B build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.problem import Problem
rs = np.random.RandomState(0)
n, D, H, O, R = 1_000_000, 16, 30, 10, int(os.environ.get("R", "1024"))
x = rs.standard_normal((n, D)).astype(np.float32); y = rs.randint(0, O, n).astype(np.int32)
P = D*H+H*O+H+O
th = torch.as_tensor(rs.standard_normal((R, P)).astype(np.float32)).cuda()
prob = Problem.from_f32([D, H, O], x, y)
out = prob.loglik(th, 0, path=_lib.PATH_TC); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = prob.loglik(th, 0, path=_lib.PATH_TC)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/3
print(os.environ.get("SAPT_LIB", "default"), "ms %.2f evals/s %.3e" % (ms, R*n/ms*1e3), flush=True)
