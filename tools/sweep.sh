#!/bin/bash
# knob sweep: level-0 kernel times + cycle time at 4097^2. args: "TH STAGES CTAS LIBVARIANT"
L=multigrid-feanet_b200/mgfea
cp $L/libmgfea.so /tmp/libmgfea_keep.so
for cfg in "${@}"; do
  set -- $cfg
  if [ -n "$4" ] && [ -f $L/libmgfea_$4.so ]; then cp $L/libmgfea_$4.so $L/libmgfea.so; fi
  MGFEA_TH=$1 MGFEA_STAGES=$2 MGFEA_CTAS=$3 timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    r=d['roofline']; print('TH=$1 stages=$2 ctas=$3 lib=$4', 'cycle_ms=%.4f'%d['ms_per_step'], 'down=%.4f up=%.4f'%(r['other_kernel_ms']['down_leg'], r['other_kernel_ms']['up_leg']), 'cyc/s=%.0f'%d['value'], 'frac=%.3f'%r['cycle']['frac'])
"
done
cp /tmp/libmgfea_keep.so $L/libmgfea.so
