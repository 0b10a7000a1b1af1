"""Numerical emulation (numpy, fp64 accumulation) of the split-operand products of the likelihood kernel's first layer:
three bf16 pieces / six products (round 1) vs two fp16 pieces / three products (round 2) vs a plain fp32 dot product.
Result: the fp16 scheme is unbiased and as accurate as an fp32 dot product; the bf16 scheme is ~10x tighter."""
import numpy as np

rs = np.random.RandomState(0)


def bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000
    return u.astype(np.uint32).view(np.float32)


def split_bf3(x):
    p1 = bf16(x); r = x - p1; p2 = bf16(r); r2 = r - p2
    return p1, p2, bf16(r2)


def split_h2(x):
    hi = x.astype(np.float16).astype(np.float32)
    return hi, (x - hi).astype(np.float16).astype(np.float32)


f = lambda a, b: a.astype(np.float64) @ b.astype(np.float64)
for D, scale_w in ((16, 1.0), (54, 1.0), (54, 5.0)):
    N = 200000
    x = rs.standard_normal((N, D)).astype(np.float32)
    w = (rs.standard_normal(D) * scale_w * 1.4427).astype(np.float32)
    ref = f(x, w)
    x1, x2, x3 = split_bf3(x); w1, w2, w3 = split_bf3(w)
    xh, xl = split_h2(x); wh, wl = split_h2(w)
    for name, z in (("bf16 x3, 6 products", f(x1, w1) + f(x1, w2) + f(x2, w1) + f(x1, w3) + f(x2, w2) + f(x3, w1)),
                    ("fp16 x2, 3 products", f(xh, wh) + f(xh, wl) + f(xl, wh)), ("plain fp32 dot", (x @ w).astype(np.float64))):
        e = z - ref
        print(f"D={D} |w|~{scale_w}: {name:20s} rms err {e.std():.2e}  mean err {e.mean():+.2e}  max {np.abs(e).max():.2e}  (rms |z| {ref.std():.2f})")
