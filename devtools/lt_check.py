"""Development check of the tcgen05 batch-Langevin kernel (sapt_langevin mode 1) against the oracle and the
CUDA-core kernel (mode 2). Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sapt_oracle as so
from sapt_b200.problem import Problem

def synth(D, O, n, seed):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, D)); y = rs.randint(0, O, n)
    return np.concatenate([x, y[:, None].astype(np.float64)], axis=1)

topos = [tuple(int(v) for v in a.split("-")) for a in sys.argv[1:]] or [(9, 12, 2), (16, 30, 10), (34, 50, 2), (54, 64, 7), (20, 50, 2), (6, 25, 7)]
for (D, H, O) in topos:
    topo = [D, H, O]
    n = 700
    data = synth(D, O, n, 11)
    P = so.num_param(topo)
    R = 5
    W = (np.random.RandomState(3).standard_normal((R, P)) * 0.4).astype(np.float32)
    prob = Problem(topo, data, None)
    th = torch.as_tensor(W).cuda()
    lr = 0.001
    o1 = prob.langevin(th, lr, mode=1, batch_scale=1.0).cpu().numpy().astype(np.float64); torch.cuda.synchronize()
    o2 = prob.langevin(th, lr, mode=2, batch_scale=1.0).cpu().numpy().astype(np.float64); torch.cuda.synchronize()
    worst1 = worst2 = 0.0
    for k in range(R):
        g = so.langevin_batch_gradient(data, W[k].astype(np.float64), topo)
        g1 = (o1[k] - W[k]) / lr; g2 = (o2[k] - W[k]) / lr
        worst1 = max(worst1, np.max(np.abs(g1 - g)) / np.max(np.abs(g))); worst2 = max(worst2, np.max(np.abs(g2 - g)) / np.max(np.abs(g)))
        if k == 0:
            # per block
            a, b, c = D * H, D * H + H * O, D * H + H * O + H
            for nm, sl in (("W1", slice(0, a)), ("W2", slice(a, b)), ("B1", slice(b, c)), ("B2", slice(c, P))):
                print(f"   {nm}: tc err {np.max(np.abs(g1[sl]-g[sl])):.3e} f32 err {np.max(np.abs(g2[sl]-g[sl])):.3e} |g| {np.max(np.abs(g[sl])):.3e}")
    print(f"{D}-{H}-{O}: max err / max|g|: tc {worst1:.3e}  cuda-core {worst2:.3e}", flush=True)
