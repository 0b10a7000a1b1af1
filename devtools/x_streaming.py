"""X-streaming regime (few chains per GPU: SURVEY 8d, R_gpu * H < ~430): achieved HBM GB/s of the likelihood kernels
at the C4 shape (54-64-7, 581 012 rows) with 1..8 chains. L2 is flushed between launches (the images fit in L2 otherwise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200 import _lib
from sapt_b200.problem import Problem
rs = np.random.RandomState(0)
D, H, O, N = 54, 64, 7, 581012
x = rs.standard_normal((N, D)).astype(np.float32); y = rs.randint(0, O, N).astype(np.int32)
prob = Problem.from_f32([D, H, O], x, y)
P = prob.P
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
n_pad = ((N + 255) // 256) * 256
bytes_tc = n_pad * 2 * 128 + n_pad * 4       # fp16 two-piece image [x_hi | x_lo] (bias in the spare x_hi column): 2 blocks x 128 B per row, + labels
bytes_f32 = n_pad * 64 * 4 + n_pad * 4       # fp32 X padded to 64 columns, + labels
for R in (1, 2, 4, 8, 16):
    th = torch.as_tensor((rs.standard_normal((R, P)) * 0.3).astype(np.float32)).cuda()
    for path, nm, nbytes in ((_lib.PATH_TC, "tcgen05", bytes_tc), (_lib.PATH_F32, "fp32", bytes_f32)):
        prob.loglik(th, 0, path=path); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); prob.loglik(th, 0, path=path); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        print(f"R={R:2d} {nm:8s}: {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s of operand image  ({R*N/ms/1e6:.2f}e9 evals/s)", flush=True)
