"""Run a few eager V-cycles at 4097^2 (no CUDA graph) -- target for `ncu` launch lists / full captures.
usage: cycle_profile.py [ncycles] [n] [mode: iso|hjac_iface]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

ncyc = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mode = sys.argv[3] if len(sys.argv) > 3 else "iso"
from FEANet.drivers import Multigrid

np.random.seed(123)
prob = Multigrid(n)
eng = prob._engine(1, 1, 0, B=1)
eng.set_u(prob.initial_v.reshape(1, 1, n + 1, n + 1))
eng.set_f(torch.zeros(1, 1, n + 1, n + 1))
torch.cuda.synchronize()
for _ in range(ncyc):
    eng.cycle()
torch.cuda.synchronize()
print("res", float(torch.sqrt(eng.sumsq.sum()).item()))
