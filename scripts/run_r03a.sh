#!/bin/bash
# round-2 session-2 run a: async-tail tests, quick bench, ncu full captures of the detector heads / fused-upsample layers
O=gpurun_out
python -m pytest tests/test_engine_gpu.py -x -q -k "async_tail or pipelined or run_stream_matches or graph_replay" > $O/r03a_pytest.log 2>&1
tail -n 3 $O/r03a_pytest.log
python bench.py --no-cpu --no-extra --steps 20 --warmup 3 > $O/r03a_bench.json 2> $O/r03a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03a_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['nets_ms'], d['roofline']['frac'], d.get('stage_ms'))
PY
NCU_OPS=yolo:58:60 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/r03a_head13 python scripts/ncu_step.py 64 > $O/r03a_ncu1.log 2>&1
NCU_OPS=yolo:74:75 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/r03a_head52 python scripts/ncu_step.py 64 > $O/r03a_ncu2.log 2>&1
for f in r03a_head13 r03a_head52; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
  ncu -i $O/$f.ncu-rep --page source --csv > $O/$f.source.csv 2>/dev/null
done
ls -la $O/r03a_*
