"""Timeline of one replayed row-slab V-cycle on every rank (mgfea_trace stamps around each slab kernel, exchange step
and coarse-cycle launch).  Run under torchrun: slab_trace.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch
import torch.distributed as dist

import mgfea
from FEANet.distributed import SlabMultigrid

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else {2: 8192, 4: 8192, 8: 16384}[world]
mg = SlabMultigrid(n)
N = n + 1


def u_rows(row0, nrows, NN):
    g = torch.Generator(device="cuda").manual_seed(123 + rank)
    return 1.2e5 * torch.rand((nrows, NN), generator=g, device="cuda") + 1.3e5


mg.fill_local(u_rows)
mg.exchange_initial()
for _ in range(3):
    mg._cycle_eager()
torch.cuda.synchronize()
dist.barrier()
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
mgfea.lib().mgfea_trace(buf.data_ptr(), 128)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    mg._cycle_eager()
mgfea.lib().mgfea_trace(None, 0)
acc, reps = None, 20
for r in range(reps + 3):
    dist.barrier()
    g.replay()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().astype(np.int64)[:128]
    k = int((t > 0).sum())
    t = t[:k]
    if r >= 3:
        acc = (t - t[0]) if acc is None else acc + (t - t[0])
acc = acc / reps / 1e3
durs = [acc[i + 1] - acc[i] for i in range(0, k, 2)]
out = [None] * world
dist.all_gather_object(out, (float(acc[k - 1]), [round(float(d), 2) for d in durs]))
if rank == 0:
    ld = mg.part.ld
    print(f"n={n} world={world} first replicated level {ld} exchange={'peer' if mg.peer is not None else 'nccl'}")
    for q, (tot, d) in enumerate(out):
        print(f"rank {q}: total {tot:8.2f} us  launches: {d}")
torch.cuda.synchronize()
os._exit(0)
