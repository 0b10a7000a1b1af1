"""Development check of the tcgen05 path against the fp32 path and the oracle (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
from oracle import sapt_oracle as so
from sapt_b200 import _lib
from sapt_b200.problem import Problem

for name in ("cancer", "pendigit"):
    g = gu.load(f"unit_{name}")
    topo = [int(v) for v in g["topology"]]
    prob = Problem(topo, g["train"], g["test"])
    print(name, topo, "tc_supported", prob.tc_supported, flush=True)
    W = g["W"].astype(np.float32)
    th = torch.as_tensor(W).cuda()
    L1, n1, s1 = prob.loglik(th, 0, path=_lib.PATH_F32); torch.cuda.synchronize()
    L2, n2, s2 = prob.loglik(th, 0, path=_lib.PATH_TC); torch.cuda.synchronize()
    ref = np.array([so.loglik_stats(g["train"], W[k].astype(np.float64), topo)[0] for k in range(W.shape[0])])
    print(" f32", L1.cpu().numpy()); print(" tc ", L2.cpu().numpy()); print(" ref", ref)
    print(" rel err f32 %.3e  tc %.3e" % (np.max(np.abs(L1.cpu().numpy()-ref)/np.abs(ref)), np.max(np.abs(L2.cpu().numpy()-ref)/np.abs(ref))))
    print(" nc", n1.cpu().numpy(), n2.cpu().numpy(), "sq", s1.cpu().numpy(), s2.cpu().numpy(), flush=True)

# C3 shape timing
rs = np.random.RandomState(0)
n, D, H, O, R = 1_000_000, 16, 30, 10, 1024
x = rs.standard_normal((n, D)).astype(np.float32); y = rs.randint(0, O, n).astype(np.int32)
P = D*H+H*O+H+O
W = (rs.standard_normal((R, P))).astype(np.float32)
th = torch.as_tensor(W).cuda()
prob = Problem.from_f32([D, H, O], x, y)
for path, nm in ((_lib.PATH_F32, "f32"), (_lib.PATH_TC, "tc")):
    out = prob.loglik(th, 0, path=path); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = prob.loglik(th, 0, path=path)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    print(nm, "ms", ms, "evals/s %.3e" % (R*n/ms*1e3), "L[:3]", out[0][:3].cpu().numpy(), flush=True)
    if nm == "f32": Lf = out[0].cpu().numpy(); nf = out[1].cpu().numpy()
    else:
        Lt = out[0].cpu().numpy(); nt = out[1].cpu().numpy()
        print("tc vs f32 max rel", np.max(np.abs(Lt-Lf)/np.abs(Lf)), "max abs", np.max(np.abs(Lt-Lf)), "nc diff max", np.max(np.abs(nt-nf)))
