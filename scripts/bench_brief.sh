#!/bin/bash
# brief bench summary: scripts/bench_brief.sh [bench.py args]
python bench.py --steps ${STEPS:-10} --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('ms/step',round(d['ms_per_step'],3),'fps',round(d['frames_per_s']),'iters/s',round(d['value']),'roof(irls it)',r['frac'],'whole step',r['whole_step']['frac'],'e2e fps',round(d['e2e']['frames_per_s']),'e2e ok',d['e2e']['matches_device_run_bitwise'],'clk',d['clocks']['sm_mhz'], d['clocks']['reasons'])
for n,e in r['kernels'].items():
    print('  %-90s %8.4f ms  %7.1f GB/s  frac %.3f' % (n[:90], e['ms_per_step'], e['achieved'], e['frac']))
print('  sum of per-kernel ms (one stream):', round(sum(e['ms_per_step'] for n,e in r['kernels'].items() if 'all ' not in n and 'irls_pass' not in n),3))
"
