#!/bin/bash
# run b: per-tile clock-stamp trace of the production shapes (harness), network parity tests with the staged upsample
# store + fp32 head epilogue, async-tail tests, quick bench
O=gpurun_out
timeout 300 betapose_b200/csrc/build/conv_harness trace > $O/r03b_trace.log 2>&1
tail -n 40 $O/r03b_trace.log
python -m pytest tests/test_nets_gpu.py tests/test_engine_gpu.py -x -q -k "yolov3 or fastpose or async_tail or pipelined or stagewise" > $O/r03b_pytest.log 2>&1
tail -n 5 $O/r03b_pytest.log
python bench.py --no-cpu --no-extra --steps 20 --warmup 3 --dump-ops $O/r03b_ops.json > $O/r03b_bench.json 2> $O/r03b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03b_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['frac'], d.get('stage_ms'))
PY
