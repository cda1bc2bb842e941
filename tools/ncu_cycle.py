"""three eager 4097^2 V-cycles (no graph, no control block) for `ncu -k regex:mg_stream2 -s 5 -c 5` (second cycle)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import torch

from bench import model_u0
from FEANet.drivers import Multigrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = n + 1
prob = Multigrid(n)
eng = prob._engine(1, 1, 0, B=1)
eng.set_u(torch.from_numpy(model_u0(n)).reshape(1, 1, N, N))
eng.set_f(torch.zeros(1, 1, N, N))
for _ in range(3):
    eng.cycle()
torch.cuda.synchronize()
print("ok", float(eng.sumsq.sum().item()))
