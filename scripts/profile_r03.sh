#!/bin/bash
# ncu evidence for round 2, second session (build with the store warp) (run under gpurun on ONE GPU): launch list, DRAM counters of every kernel of one step,
# --set full captures of the stage kernels, the PnP kernels and three conv variants.
set -x
O=gpurun_out
M1=gpu__time_duration.sum
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --profile-from-start off --metrics $M1 --clock-control none --csv --log-file $O/r03_launches_b64.csv python scripts/ncu_step.py 64 > $O/r03_ncu1.log 2>&1
ncu --profile-from-start off --metrics $M2 --clock-control none --csv --log-file $O/r03_counters_b64.csv python scripts/ncu_step.py 64 > $O/r03_ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'resize|yolo_decode|crop_resize|heatmap_decode|pnp_|pack_records' -o $O/r03_stage_kernels python scripts/ncu_step.py 64 > $O/r03_ncu3.log 2>&1
# conv: the grouped stem (launch 1 of the detector), a CTA-pair 3x3 layer, a 20x16 1x1 layer of the key-point net
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -c 1 -o $O/r03_conv_stem python scripts/ncu_step.py 64 > $O/r03_ncu4.log 2>&1
# one FastPose 20x16 bottleneck: CTA-pair 3x3 256->256, 1x1 256->1024 + residual, 1x1 1024->256 (ops 60..62 of the key-point net)
NCU_OPS=kpd:60:63 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/r03_kpd_bottleneck python scripts/ncu_step.py 64 > $O/r03_ncu5.log 2>&1
for f in r03_stage_kernels r03_conv_stem r03_kpd_bottleneck; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
done
ls -la $O/r03_*
# detector heads (fp32, per-thread stores) and a fused-upsample layer (staged, row-wise coalesced stores): ops 58..59 and 74 of the detector
NCU_OPS=yolo:58:60 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/r03_head_up2 python scripts/ncu_step.py 64 > $O/r03_ncu6.log 2>&1
ncu -i $O/r03_head_up2.ncu-rep --page raw --csv > $O/r03_head_up2.raw.csv 2>/dev/null
betapose_b200/csrc/build/conv_harness trace > $O/r03_harness_trace.log 2>&1
ls -la $O/r03_*
