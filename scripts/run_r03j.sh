#!/bin/bash
O=gpurun_out
python -m pytest tests -q -m gpu > $O/r03j_pytest.log 2>&1; tail -n 4 $O/r03j_pytest.log
for cfg in "128 1" "128 2" "64 2"; do
  set -- $cfg
  python bench.py --no-cpu --no-extra --steps 20 --warmup 3 --batch $1 --lanes $2 > $O/r03j_bench_b$1_l$2.json 2> $O/r03j_bench_b$1_l$2.err
  python - $1 $2 <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r03j_bench_b%s_l%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
print('batch',sys.argv[1],'lanes',sys.argv[2], 'value %.0f e2e %.0f step %.3f nets %.3f one-lane %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['nets_ms_one_lane'], d['clocks']['sm_mhz']))
PY
done
