#!/bin/bash
O=gpurun_out
for o in 1 0 1 0; do
  python bench.py --no-cpu --no-extra --steps 20 --warmup 3 --e2e-first $o > $O/r03t_bench_$o.json 2> $O/r03t_bench_$o.err
  python - $o <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r03t_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('e2e-first',sys.argv[1], 'value %.0f e2e %.0f step %.3f nets %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['clocks']['sm_mhz']))
PY
done
