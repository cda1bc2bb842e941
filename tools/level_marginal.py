"""Marginal cost of every level of the iso V(1,1) cycle inside the replayed CUDA graph: time the cycle of the
hierarchies truncated to L = 1 .. log2(n) levels (Multigrid(n, final_level=L)).  usage: level_marginal.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from FEANet.drivers import Multigrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Lmax = int(np.log2(n))
prev = 0.0
for L in range(1, Lmax + 1):
    np.random.seed(123)
    prob = Multigrid(n, final_level=L)
    eng = prob._engine(1, 1, 0, B=1)
    eng.set_u(prob.initial_v.reshape(1, 1, n + 1, n + 1))
    eng.refresh()
    eng._ctl_reset(0, -1.0, eng.max_cycles)
    eng._ensure_graph()
    z = eng.ctl.clone()
    c0 = mgfea.launch_count()
    eng.cycle(use_ctl=True)
    nl = mgfea.launch_count() - c0
    eng.ctl.copy_(z)
    for _ in range(10):
        eng._graph.replay()
    eng.ctl.copy_(z)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        eng._graph.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 10
    print(f"L={L:2d} coarsest N={n // 2 ** (L - 1) + 1:5d} launches={nl:2d} cycle {us:8.2f} us  (+{us - prev:7.2f})", flush=True)
    prev = us
    del eng, prob
