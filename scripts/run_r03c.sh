#!/bin/bash
# run c: 16 epilogue warps + single-barrier epilogue: harness correctness matrix, per-tile trace, network parity, bench (async tail on / off)
O=gpurun_out
timeout 600 betapose_b200/csrc/build/conv_harness quick > $O/r03c_harness.log 2>&1
grep -c " ok" $O/r03c_harness.log; grep -i "fail" $O/r03c_harness.log | head
timeout 300 betapose_b200/csrc/build/conv_harness trace > $O/r03c_trace.log 2>&1
grep -A1 "^\[" $O/r03c_trace.log | cut -c1-400
python -m pytest tests/test_nets_gpu.py -x -q > $O/r03c_pytest.log 2>&1
tail -n 3 $O/r03c_pytest.log
python bench.py --no-cpu --no-extra --steps 20 --warmup 3 --dump-ops $O/r03c_ops.json > $O/r03c_bench.json 2> $O/r03c_bench.err
BP_ASYNC_TAIL=0 python bench.py --no-cpu --no-extra --steps 20 --warmup 3 > $O/r03c_bench_sync.json 2> $O/r03c_bench_sync.err
python - <<'PY'
import json
for f in ('r03c_bench','r03c_bench_sync'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['frac'], d['roofline']['nets_ms_one_lane'])
PY
