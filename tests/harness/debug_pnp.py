import sys, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import test_pnp_host as T
from betapose_b200 import synth, stages
from oracle import pnp as opnp, restate as R
kp = synth.synth_kp_model(1, 50)
rng = np.random.default_rng(100)
n = 8
preds = np.zeros((n, 50, 2), np.float32); truth = []
for i in range(n):
    Rt, tt, uv = T.make_case(rng, kp, 0.0)
    preds[i] = uv + np.float32(0.3); truth.append((Rt, tt))
mv = np.full((n, 50), 0.5, np.float32); det = np.ones(n, np.float32)
def run(tag, mode):
    out = stages.pose_pnp(torch.from_numpy(preds).cuda(), torch.from_numpy(mv).cuda(), torch.from_numpy(det).cuda(), torch.from_numpy(kp).cuda(), mode=mode, n_hyp=64, seed=11)
    torch.cuda.synchronize()
    st = out['status'].cpu().numpy(); Rg = out['R'].cpu().numpy().reshape(n, 3, 3)
    errs = [float(np.abs(Rg[i] - truth[i][0]).max()) for i in range(n)]
    print(tag, 'mode', mode, 'status', st.tolist(), 'inl', out['inlier'].sum(1).cpu().tolist(), 'err', ['%.1e' % e for e in errs])
run('default', 0); run('default', 1)
rt = C.CDLL('libcudart.so.12')
sz = C.c_size_t()
rt.cudaDeviceGetLimit(C.byref(sz), 0); print('stack limit', sz.value)
print('set', rt.cudaDeviceSetLimit(0, C.c_size_t(32768)))
rt.cudaDeviceGetLimit(C.byref(sz), 0); print('stack limit', sz.value)
run('big-stack', 0); run('big-stack', 1)
