#!/bin/bash
# run o: conv kernels capped at 128 registers (PnP CTAs fit beside a conv CTA again) against the previous build (139), interleaved on one box
O=gpurun_out
for v in prev cap prev cap; do
  if [ $v = cap ]; then unset BP_LIB_PATH; else export BP_LIB_PATH=$PWD/betapose_b200/libbetapose_b200_$v.so; fi
  python bench.py --no-cpu --no-extra --steps 30 --warmup 4 > $O/r03o_bench_$v.json 2> $O/r03o_bench_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r03o_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.0f e2e %.0f step %.3f nets %.3f one-lane %.3f frac %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['nets_ms_one_lane'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
done
