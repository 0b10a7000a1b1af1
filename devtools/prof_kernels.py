"""Small single-kernel workloads for ncu. usage: python devtools/prof_kernels.py c3|c4ll|c4lt|c5ll|c5sp [chains]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.problem import Problem

def synth(D, O, n, seed):
    rs = np.random.RandomState(seed)
    return rs.standard_normal((n, D)).astype(np.float32), rs.randint(0, O, n).astype(np.int32)

which = sys.argv[1]
if which == "c3":
    D, H, O, N, R = 16, 30, 10, 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    x, y = synth(D, O, N, 1234)
    prob = Problem.from_f32([D, H, O], x, y)
    th = torch.as_tensor((np.random.RandomState(1).randn(R, prob.P) * 0.5).astype(np.float32)).cuda()
    for _ in range(2):
        prob.loglik(th, 0, path=_lib.PATH_TC)
elif which in ("c4ll", "c4lt"):
    D, H, O, N, R = 54, 64, 7, 581_012, int(sys.argv[2]) if len(sys.argv) > 2 else 256
    x, y = synth(D, O, N, 1)
    prob = Problem.from_f32([D, H, O], x, y)
    th = torch.as_tensor((np.random.RandomState(1).randn(R, prob.P) * 0.5).astype(np.float32)).cuda()
    for _ in range(2):
        if which == "c4ll":
            prob.loglik(th, 0, path=_lib.PATH_TC)
        else:
            prob.langevin(th, 0.5, mode=1, batch_scale=1.0 / N)
elif which == "c5ll":
    D, H, O, N, R = 6, 25, 18, 16_833, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    x, y = synth(D, O, N, 5)
    prob = Problem.from_f32([D, H, O], x, y)
    th = torch.as_tensor((np.random.RandomState(1).randn(R, prob.P) * 0.5).astype(np.float32)).cuda()
    for _ in range(2):
        prob.loglik(th, 0, path=_lib.PATH_TC)
elif which == "c5sp":
    import ctypes as C, time
    P, R, h1, h2 = 643, int(sys.argv[2]) if len(sys.argv) > 2 else 32768, 200, 50
    rs = np.random.RandomState(3)
    u = lambda *s: torch.as_tensor(rs.uniform(-0.05, 0.05, s).astype(np.float32)).cuda()
    W1, b1, W2, b2, W3, b3 = u(P, h1), u(h1), u(h1, h2), u(h2), u(h2, 1), u(1)
    th = torch.as_tensor(rs.standard_normal((R, P + 1)).astype(np.float32)).cuda()
    s = torch.zeros(R, dtype=torch.float32).cuda()
    vp = lambda t: C.c_void_p(t.data_ptr())
    L = _lib.lib()
    def run():
        _lib.check(L.sapt_surrogate_predict(vp(th), P + 1, R, P, None, None, vp(W1), vp(b1), h1, vp(W2), vp(b2), h2, vp(W3), vp(b3), vp(s), None))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print("surrogate_predict", "tc" if os.environ.get("SAPT_SURROGATE_TC", "-1") != "0" else "cuda-core", R, "rows:", e0.elapsed_time(e1) / 5, "ms")
torch.cuda.synchronize()
