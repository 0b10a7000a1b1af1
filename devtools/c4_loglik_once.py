import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sapt_b200.problem import Problem
rs = np.random.RandomState(0)
D,H,O,N,R = 54,64,7,581012,256
x = rs.standard_normal((N, D)).astype(np.float32); y = rs.randint(0, O, N).astype(np.int32)
prob = Problem.from_f32([D,H,O], x, y)
th = torch.as_tensor((rs.standard_normal((R, prob.P))*0.3).astype(np.float32)).cuda()
for _ in range(2):
    out = prob.loglik(th, 0)
torch.cuda.synchronize()
