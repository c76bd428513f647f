#!/bin/bash
# run h: asynchronous tail on / off, interleaved on one box (8 epilogue warps + store warp build), 40 timed steps each
O=gpurun_out
for rep in 1 2 3; do
for m in 1 0; do
  BP_ASYNC_TAIL=$m python bench.py --no-cpu --no-extra --steps 40 --warmup 4 > $O/r03h_bench_$m$rep.json 2> $O/r03h_bench_$m$rep.err
  python - $m$rep <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r03h_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('async' if sys.argv[1][0]=='1' else 'sync ', 'value %.0f e2e %.0f step %.3f nets %.3f one-lane %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['nets_ms_one_lane'], d['clocks']['sm_mhz']))
PY
done
done
