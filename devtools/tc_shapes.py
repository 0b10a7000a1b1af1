"""Times the fp32 and tcgen05 likelihood paths at the config topologies (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.problem import Problem

rs = np.random.RandomState(0)
for (D, H, O) in ((9, 12, 2), (16, 30, 10), (34, 50, 2), (54, 64, 7), (64, 64, 18), (36, 25, 18), (4, 12, 3), (6, 25, 18), (8, 16, 4), (12, 16, 2), (10, 20, 3), (11, 50, 10), (20, 50, 2), (34, 15, 2), (9, 8, 2), (4, 8, 3)):
    n, R = 262144, 512
    x = rs.standard_normal((n, D)).astype(np.float32); y = rs.randint(0, O, n).astype(np.int32)
    P = D * H + H * O + H + O
    th = torch.as_tensor((rs.standard_normal((R, P)) * 0.5).astype(np.float32)).cuda()
    prob = Problem.from_f32([D, H, O], x, y)
    res = {}
    for path, nm in ((_lib.PATH_F32, "f32"), (_lib.PATH_TC, "tc")):
        if nm == "tc" and not prob.tc_supported:
            continue
        out = prob.loglik(th, 0, path=path); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = prob.loglik(th, 0, path=path)
        e1.record(); torch.cuda.synchronize()
        res[nm] = (e0.elapsed_time(e1) / 3, out[0].cpu().numpy(), out[1].cpu().numpy())
    line = f"{D}-{H}-{O}: " + "  ".join(f"{k} {v[0]:.2f} ms ({R*n/v[0]*1e3:.3e}/s)" for k, v in res.items())
    if "tc" in res:
        a, b = res["f32"][1], res["tc"][1]
        line += f"  max rel {np.max(np.abs(a-b)/np.abs(a)):.2e}  nc diff {np.max(np.abs(res['f32'][2]-res['tc'][2]))}"
    print(line, flush=True)
