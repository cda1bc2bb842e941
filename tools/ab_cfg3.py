"""A/B timing of BASELINE config 3 (two-phase circle, HNet smoother, 16-channel table R/P) at 4097^2: whole cycle and
the per-kernel trace.  MGFEA_HSTREAM_MIN_N=0 selects the tile programs, the default the streaming HNet kernels."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
out = bench.secondary_configs(which.split(","))
print(json.dumps({"hstream_min_n": os.environ.get("MGFEA_HSTREAM_MIN_N", "default"),
                  **{k: {kk: v[kk] for kk in v if kk in ("ms_per_cycle", "ms_runs", "cycle_roofline_frac", "frac")}
                     for k, v in out.items()}}))
