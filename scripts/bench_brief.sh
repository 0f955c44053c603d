#!/bin/bash
# brief bench summary: python bench.py ... | this
python bench.py --steps ${STEPS:-10} --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',round(d['ms_per_step'],3),'fps',round(d['frames_per_s']),'iters/s',round(d['value']),'roof',round(d['roofline']['frac'],3),'e2e fps',round(d['e2e']['frames_per_s']),'clk',d['clocks']['sm_mhz'])
k=d['kernel_ms_per_step']
agg={}
for n,v in k.items():
    agg[n.rsplit('_L',1)[0]]=agg.get(n.rsplit('_L',1)[0],0)+v
print({a:round(b,3) for a,b in agg.items()}, 'sum', round(sum(agg.values()),3))
print({n:v for n,v in k.items() if n.endswith('L0')})
"
