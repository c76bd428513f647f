#!/bin/bash
# run r: two against three lanes with the asynchronous tail, interleaved on one box
O=gpurun_out
for l in 2 3 2 3; do
  python bench.py --no-cpu --no-extra --steps 30 --warmup 6 --lanes $l > $O/r03r_bench_l$l.json 2> $O/r03r_bench_l$l.err
  python - $l <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r03r_bench_l%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('lanes',sys.argv[1], 'value %.0f e2e %.0f step %.3f nets %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['clocks']['sm_mhz']))
PY
done
