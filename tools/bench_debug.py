import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepglobalregistration_b200 import _abi, synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration
state = syn.make_checkpoint(0)
dgr = DeepGlobalRegistration(types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False))
dgr.use_icp = False   # profile the benchmarked unit (through the refinement)
pairs = [syn.room_pair(i, n_raw=250000) for i in range(3)]
pdev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b, _ in pairs]
def run(tag, host, profile, n=30):
    torch.cuda.synchronize()
    if profile: _abi.CONV_PROFILE = []
    ts = []
    for s in range(n):
        t = time.perf_counter()
        a, b = (pairs[s % 3][:2] if host else pdev[s % 3])
        dgr.register(a, b)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    _abi.CONV_PROFILE = None
    ts = np.array(ts); print(tag, 'median %.1f mean %.1f max %.1f n>40ms %d' % (np.median(ts), ts.mean(), ts.max(), (ts > 40).sum()), 'threads', torch.get_num_threads())
run('warm', False, False, 5)
for _ in range(2):
    run('dev ', False, False)
    run('host', True, False)
sys.exit(0)
# per-seed stage timing
for s in range(3):
    a, b = pdev[s]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        p0, c0, f0 = dgr.preprocess(a, 0); p1, c1, f1 = dgr.preprocess(b, 1); torch.cuda.synchronize(); t1 = time.perf_counter()
        F0 = dgr.fcgf_feature_extraction(f0, c0); F1 = dgr.fcgf_feature_extraction(f1, c1); torch.cuda.synchronize(); t2 = time.perf_counter()
        idx1 = _abi.knn_top1(F0, F1); torch.cuda.synchronize(); t3 = time.perf_counter()
        c6 = _abi.inlier_coords(c0, c1, idx1)
        logit = dgr.inlier_prediction(torch.ones(len(idx1), 1, device='cuda'), c6); torch.cuda.synchronize(); t4 = time.perf_counter()
        man = c6._dgr_manager if hasattr(c6, '_dgr_manager') else None
    print('seed', s, 'N', len(c0), len(c1), 'pre %.2f fcgf %.2f knn %.2f inlier %.2f' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
