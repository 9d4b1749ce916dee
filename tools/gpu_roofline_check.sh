#!/bin/bash
# tests + the GraphLayer roofline leg at C5/C4 (one short gpurun call)
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for w in ${WORKLOADS:-C5 C4}; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],4), 'fwd', round(r['fwd_ms'],4), 'bwd', round(r['bwd_ms'],4))
print('   ', {k:round(v,3) for k,v in r['kernels_ms'].items()})
print('   ', {k:round(v,3) for k,v in list(d['kernels_ms_per_step'].items())[:12]})"
done
