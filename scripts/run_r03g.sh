#!/bin/bash
# run g: same box, three builds of the library (base = committed epilogue, w16 / w8 = store warp + pipelined TMEM loads with 16 / 8 epilogue warps)
O=gpurun_out
for v in base w8 w16 base w8 w16; do
  if [ $v = w16 ]; then unset BP_LIB_PATH; else export BP_LIB_PATH=$PWD/betapose_b200/libbetapose_b200_$v.so; fi
  python bench.py --no-cpu --no-extra --steps 20 --warmup 3 > $O/r03g_bench_$v.json 2> $O/r03g_bench_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r03g_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.0f e2e %.0f step %.3f nets %.3f one-lane %.3f frac %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['nets_ms_one_lane'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
done
