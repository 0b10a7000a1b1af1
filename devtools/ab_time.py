"""A/B timing of likelihood / Langevin kernels for one libsapt build (select with SAPT_LIB). Prints one JSON line.
usage: SAPT_LIB=... python devtools/ab_time.py [c3] [c4] [acc]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.problem import Problem

def synth(D, O, n, seed):
    rs = np.random.RandomState(seed)
    return rs.standard_normal((n, D)).astype(np.float32), rs.randint(0, O, n).astype(np.int32)

def t_alone(fn, reps=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        flush.zero_(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.min(ms)), float(np.mean(ms))

which = sys.argv[1:] or ["c3", "c4"]
out = dict(lib=os.path.basename(_lib.LIB_PATH))
if "c3" in which:
    D, H, O, N, R = 16, 30, 10, 1_000_000, 1024
    x, y = synth(D, O, N, 1234)
    prob = Problem.from_f32([D, H, O], x, y)
    th = torch.as_tensor((np.random.RandomState(1).randn(R, prob.P) * 0.5).astype(np.float32)).cuda()
    out["c3_loglik_ms"] = t_alone(lambda: prob.loglik(th, 0, path=_lib.PATH_TC))
    out["c3_evals_per_s"] = R * N / (out["c3_loglik_ms"][1] * 1e-3)
    if "acc" in which:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        from oracle import sapt_oracle as so
        rs = np.random.RandomState(7)
        Rk = 6
        W = np.zeros((2 * Rk, prob.P), dtype=np.float32)
        W[:Rk] = (rs.standard_normal((Rk, prob.P)) * np.linspace(0.3, 1.5, Rk)[:, None]).astype(np.float32)
        W[Rk:] = (W[:Rk].astype(np.float64) + 0.025 * rs.standard_normal((Rk, prob.P))).astype(np.float32)
        def ora(w):
            tot = 0.0
            for a in range(0, N, 125_000):
                data = np.concatenate([x[a:a + 125_000].astype(np.float64), y[a:a + 125_000, None].astype(np.float64)], axis=1)
                tot += so.loglik_stats(data, w, [D, H, O])[0]
            return tot
        ref = np.array([ora(W[k].astype(np.float64)) for k in range(2 * Rk)])
        for name, path in (("tc", _lib.PATH_TC), ("f32", _lib.PATH_F32)):
            L = prob.loglik(torch.as_tensor(W).cuda(), 0, path=path)[0].cpu().numpy()
            out[f"c3_{name}_rel_L"] = float(np.max(np.abs(L - ref) / np.abs(ref)))
            out[f"c3_{name}_abs_L"] = float(np.max(np.abs(L - ref)))
            out[f"c3_{name}_dL_err"] = float(np.max(np.abs((L[Rk:] - L[:Rk]) - (ref[Rk:] - ref[:Rk]))))
    del prob
    torch.cuda.empty_cache()
if "c4" in which:
    D, H, O, N, R = 54, 64, 7, 581_012, 256
    x, y = synth(D, O, N, 1)
    prob = Problem.from_f32([D, H, O], x, y)
    th = torch.as_tensor((np.random.RandomState(1).randn(R, prob.P) * 0.5).astype(np.float32)).cuda()
    out["c4_loglik_ms"] = t_alone(lambda: prob.loglik(th, 0, path=_lib.PATH_TC), 3)
    out["c4_evals_per_s"] = R * N / (out["c4_loglik_ms"][1] * 1e-3)
    out["c4_langevin_ms"] = t_alone(lambda: prob.langevin(th, 0.5, mode=1, batch_scale=1.0 / N), 3)
    if "acc" in which:
        from oracle import sapt_oracle as so
        n = 20000
        data = np.concatenate([x[:n].astype(np.float64), y[:n, None].astype(np.float64)], axis=1)
        sub = Problem.from_f32([D, H, O], x[:n], y[:n])
        Ls = sub.loglik(th[:4], 0, path=_lib.PATH_TC)[0].cpu().numpy()
        ref = np.array([so.loglik_stats(data, th[k].cpu().numpy().astype(np.float64), [D, H, O])[0] for k in range(4)])
        out["c4_tc_rel_L_20k"] = float(np.max(np.abs(Ls - ref) / np.abs(ref)))
        og = sub.langevin(th[:3], 0.01, mode=1, batch_scale=1.0 / n).cpu().numpy().astype(np.float64)
        err = 0.0
        for k in range(3):
            gk = so.langevin_batch_gradient(data, th[k].cpu().numpy().astype(np.float64), [D, H, O]) / n
            err = max(err, float(np.max(np.abs((og[k] - th[k].cpu().numpy().astype(np.float64)) / 0.01 - gk)) / np.max(np.abs(gk))))
        out["c4_langevin_grad_rel_err"] = err
if "c5" in which:
    D, H, O, N, R = 6, 25, 18, 16_833, 8192
    x, y = synth(D, O, N, 5)
    prob = Problem.from_f32([D, H, O], x, y)
    th = torch.as_tensor((np.random.RandomState(1).randn(R, prob.P) * 0.5).astype(np.float32)).cuda()
    out["c5_loglik_ms"] = t_alone(lambda: prob.loglik(th, 0, path=_lib.PATH_TC))
    out["c5_evals_per_s"] = R * N / (out["c5_loglik_ms"][1] * 1e-3)
print(json.dumps(out), flush=True)
