#!/bin/bash
# knob sweep: level-0 kernel times + cycle time at 4097^2 (prints the bench roofline fields)
for cfg in "32 2 0" "32 1 0" "32 1 3" "16 2 0" "16 2 3" "24 2 0" "48 1 0" "64 1 0" "48 2 0"; do
  set -- $cfg
  MGFEA_TH=$1 MGFEA_STAGES=$2 MGFEA_CTAS=$3 timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    r=d['roofline']; print('TH=$1 stages=$2 ctas=$3', 'cycle_ms=%.4f'%d['ms_per_step'], 'down=%.4f up=%.4f'%(r['other_kernel_ms']['down_leg'], r['other_kernel_ms']['up_leg']), 'cyc/s=%.0f'%d['value'], 'frac=%.3f'%r['cycle']['frac'])
"
done
