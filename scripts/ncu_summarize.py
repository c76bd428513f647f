"""profiles/r02_summary.txt from the ncu CSVs of scripts/profile_r02.sh (one eager step at batch 64, every kernel once):
per kernel family: launches, share of the step, DRAM bytes read / written (dram__bytes_*.sum), achieved DRAM GB/s, DRAM %,
tensor-pipe %, fp64 instructions; then the stage kernels against their algorithmic bytes, and selected metrics of the
--set full captures.
    python scripts/ncu_summarize.py gpurun_out/r02_counters_b64.csv [gpurun_out/r02_stage_kernels.raw.csv ...] > profiles/r02_summary.txt"""
import collections, csv, json, os, sys

csv.field_size_limit(10 ** 9)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r[0]), {"name": r[4]})
    d[r[-3]] = (r[-1], r[-2])

def val(d, k):
    v, u = d[k]
    return float(v.replace(",", "")) * scale.get(u, 1)

agg, tot = collections.OrderedDict(), 0.0
for d in per.values():
    nm = d["name"].split("(")[0].replace("void ", "").replace("bp::", "").replace("<unnamed>::", "")
    t = val(d, "gpu__time_duration.sum")
    tot += t
    a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += val(d, "dram__bytes_read.sum"); a[3] += val(d, "dram__bytes_write.sum")
    a[4] += float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]) * t
    a[5] += float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]) * t
    a[6] += float(d["sm__inst_executed_pipe_fp64.sum"][0].replace(",", ""))
print(f"ncu counters of ONE eager step, batch 64, one lane ({len(per)} launches; serialised, cold cache: compare shares).  "
      f"Sum of kernel durations {tot / 1e3:.3f} ms.  HBM peak of record {HBM:.0f} GB/s (MEASURED_PEAKS.json).")
print("conv_umma_kernel<BLOCK_N, BLOCK_K, STAGES, CG (2 = CTA pairs), NBUF, MT (2 / 4 = 256- / 512-pixel tiles)>\n")
print(f"{'kernel':56s} {'n':>3s} {'us':>8s} {'share':>6s} {'DRAM rd MB':>10s} {'DRAM wr MB':>10s} {'GB/s':>7s} {'of HBM':>6s} {'dram%':>5s} {'tensor%':>7s} {'fp64 inst':>9s}")
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / a[1] / 1e3
    print(f"{nm[:56]:56s} {a[0]:3d} {a[1]:8.1f} {100 * a[1] / tot:5.1f}% {a[2] / 1e6:10.1f} {a[3] / 1e6:10.1f} {gbs:7.0f} {100 * gbs / HBM:5.1f}% {a[4] / a[1]:5.1f} {a[5] / a[1]:7.1f} {a[6]:9.3g}")

# the stage kernels of the path against SURVEY 8(d)'s algorithmic bytes per frame
B = 64
alg = {"resize_fused_kernel": ("a1", 921600 + 1038336), "yolo_decode_argmax_kernel": ("a3-a5", 255528 + 36), "crop_resize_kernel": ("a6", 491520),
       "heatmap_decode_kernel": ("a8", 1024000 + 600)}
print("\nstage kernels: measured DRAM traffic (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch) against the algorithmic bytes of SURVEY 8(d)")
print(f"{'kernel':28s} {'row':>5s} {'algorithmic MB':>14s} {'measured MB':>11s} {'ratio':>6s} {'us':>7s} {'GB/s on algorithmic':>19s} {'GB/s measured':>13s} {'of HBM peak':>11s}")
for nm, (row, per_frame) in alg.items():
    if nm not in agg:
        continue
    a = agg[nm]
    meas = a[2] + a[3]
    print(f"{nm:28s} {row:>5s} {per_frame * B / 1e6:14.1f} {meas / 1e6:11.1f} {meas / (per_frame * B):6.2f} {a[1]:7.1f} {per_frame * B / a[1] / 1e3:19.0f} {meas / a[1] / 1e3:13.0f} {100 * meas / a[1] / 1e3 / HBM:10.1f}%")
print("(crop: + the crop window's uint8 reads, which depend on the box; decode reads only the objectness column of the fp32 heads)")

want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
for path in sys.argv[2:]:
    rr = list(csv.reader(open(path)))
    if len(rr) < 3:
        continue
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"\n==== ncu --set full: {os.path.basename(path)}")
    for r in rr[2:]:
        print("## " + r[idx["Kernel Name"]].split("(CUtensorMap")[0].split("(const")[0].replace("void ", "").replace("<unnamed>::", "")[:110])
        for w in want:
            if w in idx and r[idx[w]] not in ("", "n/a"):
                print(f"   {w} [{units[idx[w]]}] = {r[idx[w]]}")
