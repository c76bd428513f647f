"""Where the end-to-end stream's time goes beyond K x the device step (bench.py: timed_e2e): host->device rate of a pinned batch,
GPU-side completion time of every batch of a K-batch stream from an empty pipeline, and the host's total.
    python scripts/diag_e2e.py [K]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betapose_b200 import synth
from betapose_b200.engine import PipelinedEngine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = 64
pipe = PipelinedEngine(2, B, synth.cached_yolo_weights(1000), synth.cached_kpd_state_dict(2000), synth.synth_kp_model(1, 50))
host = [torch.from_numpy(synth.synth_frames(16, seed=100 + s)).repeat(4, 1, 1, 1)[:B].contiguous().pin_memory() for s in range(4)]
dev = host[0].cuda()
torch.cuda.synchronize()
# pinned host -> device rate
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    e0.record(); dev.copy_(host[1], non_blocking=True); e1.record(); torch.cuda.synchronize()
print("H2D of one batch (%.1f MB): %.3f ms = %.1f GB/s" % (host[1].numel() / 1e6, e0.elapsed_time(e1), host[1].numel() / e0.elapsed_time(e1) / 1e6))
for _ in pipe.run_stream((host[i % 4] for i in range(6))):
    pass
torch.cuda.synchronize()
for rep in range(2):
    stamps = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rec in pipe.run_stream((host[i % 4] for i in range(K))):
        stamps.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    d = np.diff([0.0] + stamps) * 1e3
    print("K=%d total %.2f ms (%.0f images/s); first record after %.2f ms; periods ms: %s; after last record %.2f ms"
          % (K, total * 1e3, K * B / total, stamps[0] * 1e3, " ".join("%.2f" % x for x in d[1:]), (total - stamps[-1]) * 1e3))
    # steady-state period from the middle of the stream
    mid = d[4:-2]
    print("   steady period %.3f ms per batch -> %.0f images/s; fill + drain overhead %.2f ms" % (mid.mean(), B / mid.mean() * 1e3, total * 1e3 - K * mid.mean()))

# the same stream with the frames already on the device (no PCIe upload; every host synchronisation and the read-back stay)
devs = [h.cuda() for h in host]
torch.cuda.synchronize()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rec in pipe.run_stream((devs[i % 4] for i in range(K))):
        pass
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print("device-resident inputs through run_stream: K=%d total %.2f ms (%.0f images/s)" % (K, total * 1e3, K * B / total))
# bench.py's `value` loop: submit_device K times, no host synchronisation in between
for rep in range(2):
    torch.cuda.synchronize()
    e0.record(); pipe.fork()
    for i in range(K):
        pipe.submit_device(i, devs[i % 4])
    pipe.join(); e1.record(); torch.cuda.synchronize()
    print("submit_device loop: K=%d total %.2f ms (%.0f images/s)" % (K, e0.elapsed_time(e1), K * B / e0.elapsed_time(e1) * 1e3))
