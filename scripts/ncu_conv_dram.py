"""Turn an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_umma) of ONE
bench step into profiles/conv_dram_traffic.json: mean DRAM bytes per conv launch (bench.py reports it as roofline.traffic).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_umma \
      -s <3 warm-up steps x 190> -c 190 --csv --log-file gpurun_out/conv_dram.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu
  python scripts/ncu_conv_dram.py gpurun_out/conv_dram.csv 64
"""
import csv
import json
import os
import sys

path, batch = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    per.setdefault(int(r[0]), {})[r[-3]] = (float(r[-1].replace(",", "")), r[-2])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
n = len(per)
rd = sum(v["dram__bytes_read.sum"][0] * scale[v["dram__bytes_read.sum"][1]] for v in per.values())
wr = sum(v["dram__bytes_write.sum"][0] * scale[v["dram__bytes_write.sum"][1]] for v in per.values())
tm = sum(v["gpu__time_duration.sum"][0] * scale[v["gpu__time_duration.sum"][1]] for v in per.values())
out = {"batch": batch, "launches": n, "dram_bytes_read": rd, "dram_bytes_written": wr, "dram_bytes_per_launch": (rd + wr) / n,
       "kernel_time_s_under_ncu": tm, "source": os.path.basename(path)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "conv_dram_traffic.json"), "w"), indent=1)
print(json.dumps(out))
